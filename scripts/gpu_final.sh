set -x
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err; cat gpurun_out/bench_final.json
python bench.py --impl reference > gpurun_out/bench_final_ref.json 2>&1; cat gpurun_out/bench_final_ref.json
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --verify 0 > gpurun_out/plain_final.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --verify 0 > gpurun_out/ncu_launch_final.log 2>&1
tail -2 gpurun_out/ncu_launch_final.log | cut -c1-300
gmix_b200/lib/ncu_case 1184 2048 > gpurun_out/ncu_case_final.log 2>&1 && cat gpurun_out/ncu_case_final.log &&
ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section SourceCounters --section LaunchStats --section Occupancy --section InstructionStats --section ComputeWorkloadAnalysis --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --import-source on --replay-mode application -k regex:StreamKernel -o /tmp/prof_final gmix_b200/lib/ncu_case 1184 2048 > gpurun_out/ncu_final.log 2>&1
tail -2 gpurun_out/ncu_final.log
ncu -i /tmp/prof_final.ncu-rep --page raw --csv > gpurun_out/prof_final_raw.csv 2>/dev/null
ncu -i /tmp/prof_final.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/prof_final_sass.csv.gz
ncu -i /tmp/prof_final.ncu-rep --page details 2>/dev/null > gpurun_out/prof_final_details.txt
ls -la gpurun_out | tail -12
