set -x
python -m pytest tests/test_gpu_host_layer.py -m gpu -x -q 2>&1 | tail -5
gmix_b200/lib/ncu_case 1184 2048 > gpurun_out/ncu_case_plain5.log 2>&1 && cat gpurun_out/ncu_case_plain5.log &&
ncu --section SpeedOfLight --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis --section MemoryWorkloadAnalysis_Tables --section SourceCounters --section LaunchStats --section Occupancy --section InstructionStats --section ComputeWorkloadAnalysis --metrics dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sectors_op_read.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_atom.sum,lts__t_sectors_op_red.sum,lts__t_sectors_srcunit_tex_lookup_hit.sum,lts__t_sectors_srcunit_tex_lookup_miss.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum --clock-control none --import-source on --replay-mode application -k regex:StreamKernel -o /tmp/prof_v5 gmix_b200/lib/ncu_case 1184 2048 > gpurun_out/ncu_v5.log 2>&1
tail -2 gpurun_out/ncu_v5.log
ncu -i /tmp/prof_v5.ncu-rep --page raw --csv > gpurun_out/prof_v5_raw.csv 2>/dev/null
ncu -i /tmp/prof_v5.ncu-rep --page source --csv --print-source sass 2>/dev/null | gzip > gpurun_out/prof_v5_sass.csv.gz
ncu -i /tmp/prof_v5.ncu-rep --page details 2>/dev/null > gpurun_out/prof_v5_details.txt
cuobjdump -xelf kernel_compress gmix_b200/lib/libgmix_b200.so > /dev/null 2>&1; ls *.cubin 2>/dev/null
ls -la gpurun_out
