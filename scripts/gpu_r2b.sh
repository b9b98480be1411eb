set -x
python -m pytest tests -m gpu -q 2>&1 | tail -8
python scripts/gpu_sweep.py 8192 1 8 > gpurun_out/sweep_r2b.log 2>&1; cat gpurun_out/sweep_r2b.log
python scripts/gpu_profile.py 8192 1184 > gpurun_out/phase_r2b.log 2>&1; cat gpurun_out/phase_r2b.log
gmix_b200/lib/ncu_case 1184 2048 > gpurun_out/ncu_case_r2b.log 2>&1 && cat gpurun_out/ncu_case_r2b.log &&
ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section SourceCounters --section LaunchStats --section Occupancy --section InstructionStats --section ComputeWorkloadAnalysis --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --import-source on --replay-mode application -k regex:StreamKernel -o /tmp/prof_r2b gmix_b200/lib/ncu_case 1184 2048 > gpurun_out/ncu_r2b.log 2>&1
tail -2 gpurun_out/ncu_r2b.log
ncu -i /tmp/prof_r2b.ncu-rep --page raw --csv > gpurun_out/prof_r2b_raw.csv 2>/dev/null
ncu -i /tmp/prof_r2b.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/prof_r2b_sass.csv.gz
ncu -i /tmp/prof_r2b.ncu-rep --page source --csv --print-source cuda 2>/dev/null | gzip > gpurun_out/prof_r2b_cuda.csv.gz
ncu -i /tmp/prof_r2b.ncu-rep --page details 2>/dev/null > gpurun_out/prof_r2b_details.txt
ls -la gpurun_out | tail -8
