#!/bin/bash
# One parameterised driver for everything that runs on the GPU box:
#   gpurun --timeout S -- 'bash scripts/gpu.sh TAG step [step ...]'
# Every step writes gpurun_out/<TAG>_<step>.* . Steps:
#   tests            python -m pytest tests -m gpu
#   smoke            __graft_entry__.smoke()
#   bench[:args]     python bench.py <args>           (':'-separated extra arguments, e.g. bench:--steps:2)
#   ref[:args]       python bench.py --impl reference <args>
#   phase:N:LEN      per-phase cycle counters (PROF kernel variant), N streams of LEN bytes
#   sweep:LEN:a,b,c  throughput vs resident streams per SM
#   launches         ncu launch list (gpu__time_duration) of a 1-step bench run, after the same command ran clean
#   traffic:N:LEN    ncu dram bytes + executed instructions of the compress kernel, ncu_case N x LEN (application replay)
#   ncufull:N:LEN    ncu sections + source counters of the compress kernel; raw/details/sass pages exported
#   ab:LEN:v1,v2     A/B of library variants under gmix_b200/lib/variants/ (golden parity check, then one wave of LEN-byte streams)
#   py:script[:args] python <script> <args>
TAG=$1; shift
O=gpurun_out
mkdir -p $O
for step in "$@"; do
  IFS=':' read -r -a A <<< "$step"
  name=${A[0]}
  echo "=== $TAG $step"
  case $name in
    tests) python -m pytest tests -m gpu -q -x 2>&1 | tail -15 | tee $O/${TAG}_tests.log ;;
    smoke) python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tee $O/${TAG}_smoke.log ;;
    bench) python bench.py "${A[@]:1}" > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; tail -3 $O/${TAG}_bench.err; cat $O/${TAG}_bench.json ;;
    ref) python bench.py --impl reference "${A[@]:1}" > $O/${TAG}_ref.json 2> $O/${TAG}_ref.err; tail -3 $O/${TAG}_ref.err; cat $O/${TAG}_ref.json ;;
    phase) python scripts/gpu_profile.py ${A[2]} ${A[1]} 2>&1 | tee $O/${TAG}_phase_${A[1]}x${A[2]}.txt ;;
    sweep) python scripts/gpu_sweep.py ${A[1]} ${A[2]//,/ } 2>&1 | tee $O/${TAG}_sweep_${A[1]}.txt ;;
    launches)
      B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-legs --verify 0"
      $B > $O/${TAG}_plain.log 2>&1 &&
      ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${TAG}_launches.csv $B > $O/${TAG}_ncu_launch.log 2>&1
      tail -2 $O/${TAG}_ncu_launch.log | cut -c1-300 ;;
    traffic)
      gmix_b200/lib/ncu_case ${A[1]} ${A[2]} > $O/${TAG}_case_${A[1]}x${A[2]}.log 2>&1 && cat $O/${TAG}_case_${A[1]}x${A[2]}.log &&
      ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum --clock-control none --csv \
          --replay-mode application -k regex:StreamKernel --log-file $O/${TAG}_traffic_${A[1]}x${A[2]}.csv gmix_b200/lib/ncu_case ${A[1]} ${A[2]} > $O/${TAG}_traffic.log 2>&1
      tail -8 $O/${TAG}_traffic_${A[1]}x${A[2]}.csv | cut -c1-400 ;;
    ncufull)
      gmix_b200/lib/ncu_case ${A[1]} ${A[2]} > $O/${TAG}_case_${A[1]}x${A[2]}.log 2>&1 && cat $O/${TAG}_case_${A[1]}x${A[2]}.log &&
      ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section SourceCounters \
          --section LaunchStats --section Occupancy --section InstructionStats --section ComputeWorkloadAnalysis \
          --metrics dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --import-source on --replay-mode application \
          -k regex:StreamKernel -o /tmp/prof_$TAG gmix_b200/lib/ncu_case ${A[1]} ${A[2]} > $O/${TAG}_ncufull.log 2>&1
      tail -2 $O/${TAG}_ncufull.log
      ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > $O/${TAG}_ncu_raw_${A[1]}x${A[2]}.csv 2>/dev/null
      ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > $O/${TAG}_ncu_sass.csv.gz
      ncu -i /tmp/prof_$TAG.ncu-rep --page details 2>/dev/null > $O/${TAG}_ncu_details_${A[1]}x${A[2]}.txt ;;
    ab)
      for v in default ${A[2]//,/ }; do
        if [ "$v" = default ]; then unset GMIX_B200_LIB; else export GMIX_B200_LIB=$PWD/gmix_b200/lib/variants/$v/libgmix_b200.so; fi
        echo "== $v"
        python scripts/gpu_parity_quick.py && python scripts/gpu_sweep.py ${A[1]} 0 2>&1 | tail -1
      done 2>&1 | tee $O/${TAG}_ab.txt
      unset GMIX_B200_LIB ;;
    py) python "${A[@]:1}" 2>&1 | tee $O/${TAG}_$(basename ${A[1]} .py).log ;;
    *) echo "unknown step $step" ;;
  esac
done
ls -la $O | tail -5
