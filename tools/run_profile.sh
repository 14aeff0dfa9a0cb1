#!/bin/bash
# ncu evidence for profiles/ (run on the GPU box: gpurun -- 'bash tools/run_profile.sh r1_f').
# 1. launch list of four device steps (the last one is summarised: warm, one round)
# 2. --set full capture of the timed step's scatter_kernel and of one mid-sweep probe_slice_kernel
cd "$GRAFT_REPO_ROOT"
tag=${1:-r1_x}
K='regex:scatter_kernel|probe_slice_kernel|extract_kernel|clear_counts_kernel'
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" -c 253 \
    --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/ncu_list_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scatter_kernel --launch-skip 4 -c 1 -f \
    -o gpurun_out/prof_scatter_$tag $B > gpurun_out/ncu_sc_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:probe_slice_kernel --launch-skip 220 -c 1 -f \
    -o gpurun_out/prof_probe_slice_$tag $B > gpurun_out/ncu_ps_$tag.log 2>&1
python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_err_$tag.log | tee gpurun_out/bench_$tag.json | cut -c1-400
ls -la gpurun_out | tail -8
