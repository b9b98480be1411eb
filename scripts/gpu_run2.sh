set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --chunks 128 --steps 2 --warmup 3 > gpurun_out/bench_r1_v1.json 2> gpurun_out/bench_r1_v1.err; tail -3 gpurun_out/bench_r1_v1.err; cat gpurun_out/bench_r1_v1.json
python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_r1_ref.json 2>&1; cat gpurun_out/bench_r1_ref.json
python bench.py --chunks 128 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --verify 0 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_v1.csv python bench.py --chunks 128 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --verify 0 > gpurun_out/ncu_launch.log 2>&1
tail -3 gpurun_out/ncu_launch.log
gmix_b200/lib/ncu_case 8 1024 > gpurun_out/ncu_case_plain.log 2>&1 && cat gpurun_out/ncu_case_plain.log &&
ncu --set full --clock-control none --import-source on --replay-mode application -k regex:StreamKernel -o gpurun_out/prof_v1 gmix_b200/lib/ncu_case 8 1024 > gpurun_out/ncu_full.log 2>&1
tail -5 gpurun_out/ncu_full.log
ls -la gpurun_out
