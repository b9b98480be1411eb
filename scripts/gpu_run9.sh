set -x
gmix_b200/lib/ncu_case 1036 1024 > gpurun_out/ncu_case_plain4.log 2>&1 && cat gpurun_out/ncu_case_plain4.log &&
ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section SourceCounters --section LaunchStats --section Occupancy --section InstructionStats --section ComputeWorkloadAnalysis --clock-control none --import-source on --replay-mode application -k regex:StreamKernel -o /tmp/prof_v4_1036 gmix_b200/lib/ncu_case 1036 1024 > gpurun_out/ncu_v4.log 2>&1
tail -2 gpurun_out/ncu_v4.log
ncu -i /tmp/prof_v4_1036.ncu-rep --page raw --csv > gpurun_out/prof_v4_1036_raw.csv 2>/dev/null
ncu -i /tmp/prof_v4_1036.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/prof_v4_1036_sass.csv.gz
ncu -i /tmp/prof_v4_1036.ncu-rep --page details 2>/dev/null > gpurun_out/prof_v4_1036_details.txt
ls -la gpurun_out
