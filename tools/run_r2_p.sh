#!/bin/bash
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-files-e2e --no-cpu-baseline"
M="gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__inst_executed_op_local_ld.sum,smsp__inst_executed_op_local_st.sum,l1tex__data_pipe_lsu_wavefronts.sum"
for k in 25 28; do
ncu --metrics $M --clock-control none -k regex:scatter_kernel --launch-skip 3 -c 1 --csv --log-file gpurun_out/r2p_sc_k$k.csv $B --kmer $k > gpurun_out/r2p_k$k.log 2>&1
done
VG_LIB=$PWD/varigraph_b200/libvgb200_evenskip.so ncu --metrics $M --clock-control none -k regex:scatter_kernel --launch-skip 3 -c 1 --csv --log-file gpurun_out/r2p_sc_k28skip.csv $B --kmer 28 > gpurun_out/r2p_k28skip.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scatter_kernel --launch-skip 3 -c 1 -f -o gpurun_out/prof_scatter_k28 $B --kmer 28 > gpurun_out/r2p_full.log 2>&1
ls -la gpurun_out | tail -5
