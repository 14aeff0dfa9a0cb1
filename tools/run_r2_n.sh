#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "even_k or encoder or count_random or partitioned_matches or two_level or k28" > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
tail -5 gpurun_out/r2n_pytest.log
B="--steps 5 --warmup 3 --no-files-e2e --no-cpu-baseline"
timeout 600 python bench.py $B --kmer 28 > gpurun_out/r2n_k28_window.json 2> gpurun_out/r2n_k28_window.err
VG_EVEN_WINDOW=0 timeout 600 python bench.py $B --kmer 28 > gpurun_out/r2n_k28_bytewise.json 2> gpurun_out/r2n_k28_bytewise.err
timeout 600 python bench.py $B --kmer 22 > gpurun_out/r2n_k22_window.json 2> gpurun_out/r2n_k22_window.err
timeout 600 python bench.py $B > gpurun_out/r2n_k27.json 2> gpurun_out/r2n_k27.err
python tools/show_bench.py gpurun_out/r2n_k28_window.json gpurun_out/r2n_k28_bytewise.json gpurun_out/r2n_k22_window.json gpurun_out/r2n_k27.json
