#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "even_k or encoder or count_random or partitioned_matches or two_level or k28" > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2q_pytest.log
tail -5 gpurun_out/r2q_pytest.log
B="--steps 5 --warmup 3 --no-files-e2e --no-cpu-baseline"
timeout 600 python bench.py $B --kmer 28 > gpurun_out/r2q_k28_window.json 2> gpurun_out/r2q_k28_window.err
VG_LIB=$PWD/varigraph_b200/libvgb200_evenskip.so timeout 600 python bench.py $B --kmer 28 > gpurun_out/r2q_k28_skip.json 2> gpurun_out/r2q_k28_skip.err
timeout 600 python bench.py $B --kmer 22 > gpurun_out/r2q_k22_window.json 2> gpurun_out/r2q_k22_window.err
python tools/show_bench.py gpurun_out/r2q_k28_window.json gpurun_out/r2q_k28_skip.json gpurun_out/r2q_k22_window.json
M="gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active"
ncu --metrics $M --clock-control none -k regex:scatter_kernel --launch-skip 3 -c 1 --csv --log-file gpurun_out/r2q_sc_k28.csv python bench.py --steps 1 --warmup 3 --no-files-e2e --no-cpu-baseline --kmer 28 > gpurun_out/r2q_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:scatter_kernel --launch-skip 3 -c 1 -f -o gpurun_out/prof_scatter_k28 python bench.py --steps 1 --warmup 3 --no-files-e2e --no-cpu-baseline --kmer 28 > gpurun_out/r2q_full.log 2>&1
grep -E "inst_executed|duration" gpurun_out/r2q_sc_k28.csv | cut -d, -f 12-
