set -x
python -m pytest tests -m gpu -q 2>&1 | tail -6
python bench.py --chunks 296 --chunk-bytes 4096 --steps 1 --warmup 3 --gen-bytes 64 > gpurun_out/bench_small_r2c.json 2> gpurun_out/bench_small_r2c.err; tail -3 gpurun_out/bench_small_r2c.err; cut -c1-3000 gpurun_out/bench_small_r2c.json
python scripts/gpu_sweep.py 8192 8 > gpurun_out/sweep_r2c.log 2>&1; cat gpurun_out/sweep_r2c.log
python scripts/gpu_profile.py 8192 1184 > gpurun_out/phase_r2c.log 2>&1; cat gpurun_out/phase_r2c.log
python bench.py > gpurun_out/bench_r2c.json 2> gpurun_out/bench_r2c.err; tail -3 gpurun_out/bench_r2c.err; cat gpurun_out/bench_r2c.json
