#!/bin/bash
# refresh of the ncu evidence for the final kernels: launch list of one count pass + --set full of scatter and one probe
cd "${GRAFT_REPO_ROOT:-.}"
mkdir -p gpurun_out
K='regex:scatter_kernel|probe_slice_kernel|rescatter_kernel|sum_cursors_kernel'
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-files-e2e"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" --launch-skip 150 -c 50 \
    --csv --log-file gpurun_out/launches_r2_u.csv $B > gpurun_out/ncu_list_r2_u.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scatter_kernel --launch-skip 3 -c 1 -f \
    -o gpurun_out/prof_scatter_r2_u $B > gpurun_out/ncu_sc_r2_u.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:probe_slice_kernel --launch-skip 168 -c 1 -f \
    -o gpurun_out/prof_probe_slice_r2_u $B > gpurun_out/ncu_ps_r2_u.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scatter_kernel --launch-skip 3 -c 1 -f \
    -o gpurun_out/prof_scatter_k28_r2_u $B --kmer 28 > gpurun_out/ncu_sc28_r2_u.log 2>&1
ls -la gpurun_out | tail -6
