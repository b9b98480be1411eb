set -x
python -m pytest tests -m gpu -q 2>&1 | tail -6
python -c "import __graft_entry__ as g; g.smoke()"
python bench.py > gpurun_out/bench_s3f.json 2> gpurun_out/bench_s3f.err; tail -2 gpurun_out/bench_s3f.err; cat gpurun_out/bench_s3f.json
python bench.py --impl reference > gpurun_out/bench_s3f_ref.json 2>&1; cat gpurun_out/bench_s3f_ref.json
python scripts/gpu_profile.py 8192 1184 > gpurun_out/phase_s3f.log 2>&1; cat gpurun_out/phase_s3f.log
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-decompress --no-generate --verify 0 > gpurun_out/plain_s3f.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_s3f.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-decompress --no-generate --verify 0 > gpurun_out/ncu_launch_s3f.log 2>&1
tail -2 gpurun_out/ncu_launch_s3f.log | cut -c1-300
gmix_b200/lib/ncu_case 1184 8192 > gpurun_out/ncu_case_s3f.log 2>&1 && cat gpurun_out/ncu_case_s3f.log &&
ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section SourceCounters --section LaunchStats --section Occupancy --section InstructionStats --section ComputeWorkloadAnalysis --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --import-source on --replay-mode application -k regex:StreamKernel -o /tmp/prof_s3f gmix_b200/lib/ncu_case 1184 8192 > gpurun_out/ncu_s3f.log 2>&1
tail -2 gpurun_out/ncu_s3f.log
ncu -i /tmp/prof_s3f.ncu-rep --page raw --csv > gpurun_out/prof_s3f_raw.csv 2>/dev/null
ncu -i /tmp/prof_s3f.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/prof_s3f_sass.csv.gz
ncu -i /tmp/prof_s3f.ncu-rep --page details 2>/dev/null > gpurun_out/prof_s3f_details.txt
ls -la gpurun_out | tail -8
