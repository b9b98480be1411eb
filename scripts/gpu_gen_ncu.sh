set -x
O=gpurun_out
python scripts/gpu_gen_profile.py 2 > $O/r2F_plain.log 2>&1 || exit 1
grep -q "ran 2" $O/r2F_plain.log || { cat $O/r2F_plain.log; exit 1; }
cat $O/r2F_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r2F_lockstep_tensor_launches.csv -k regex:"Gate|GenStep|StreamKernel" python scripts/gpu_gen_profile.py 2 > $O/r2F_ncu1.log 2>&1
for k in GateGemmTc GenStep GateDotsExact; do
  md=2; [ $k = GateDotsExact ] && md=1
  ncu --set full --clock-control none --import-source on -k regex:$k --launch-skip 12 --launch-count 1 -o /tmp/prof_$k python scripts/gpu_gen_profile.py $md > $O/r2F_ncu_$k.log 2>&1
  ncu -i /tmp/prof_$k.ncu-rep --page raw --csv > $O/r2F_${k}_ncu_raw.csv 2>/dev/null
  ncu -i /tmp/prof_$k.ncu-rep --page details > $O/r2F_${k}_ncu_details.txt 2>/dev/null
done
ls -la $O | grep r2F
