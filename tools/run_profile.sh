#!/bin/bash
# ncu evidence for profiles/ (run on the GPU box: gpurun -- 'bash tools/run_profile.sh r2_a').
# 1. launch list (duration + DRAM bytes) of one warm count pass of the default bench workload
# 2. --set full captures: scatter_kernel, one mid-sweep probe_slice_kernel; at human scale: rescatter_kernel, probe_slice_kernel
# 3. cbf_add_kernel (construct side) on a 250 Mb chromosome
cd "${GRAFT_REPO_ROOT:-.}"
tag=${1:-r2_x}
K='regex:scatter_kernel|probe_slice_kernel|rescatter_kernel|sum_cursors_kernel|gather_counts_kernel|extract_kernel|clear_counts_kernel'
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-files-e2e"
# 5 device steps come first (3 warm-up + 1 timed + 2 with phase events); each is 1 scatter + 48 probes + 1 sum = 50 launches
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k "$K" --launch-skip 150 -c 50 \
    --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/ncu_list_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scatter_kernel --launch-skip 3 -c 1 -f \
    -o gpurun_out/prof_scatter_$tag $B > gpurun_out/ncu_sc_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:probe_slice_kernel --launch-skip 168 -c 1 -f \
    -o gpurun_out/prof_probe_slice_$tag $B > gpurun_out/ncu_ps_$tag.log 2>&1
H="python bench.py --config human --coverage 3.75 --steps 1 --warmup 3 --no-cpu-baseline --no-files-e2e"
ncu --set full --clock-control none --import-source on -k regex:rescatter_kernel --launch-skip 200 -c 1 -f \
    -o gpurun_out/prof_rescatter_human_$tag $H > gpurun_out/ncu_rs_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:probe_slice_kernel --launch-skip 3400 -c 1 -f \
    -o gpurun_out/prof_probe_slice_human_$tag $H > gpurun_out/ncu_psh_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scatter_kernel --launch-skip 3 -c 1 -f \
    -o gpurun_out/prof_scatter_human_$tag $H > gpurun_out/ncu_sch_$tag.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:cbf_add_kernel --launch-skip 1 -c 1 -f \
    -o gpurun_out/prof_cbf_add_$tag python tools/cbf_bench.py > gpurun_out/ncu_cbf_$tag.log 2>&1
python tools/cbf_bench.py > gpurun_out/cbf_bench_$tag.json 2> gpurun_out/cbf_bench_$tag.err
ls -la gpurun_out | tail -12
