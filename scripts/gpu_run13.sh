set -x
python -m pytest tests -m gpu -x -q --durations=8 2>&1 | tail -16
python scripts/gpu_diag.py 65536 1184
python bench.py > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err; tail -3 gpurun_out/bench_v6.err; cat gpurun_out/bench_v6.json
