#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "partition or two_level or slot_order or encoder or count_random or full_size" > gpurun_out/r2l_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2l_pytest.log
tail -3 gpurun_out/r2l_pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-files-e2e --no-cpu-baseline > gpurun_out/r2l_chr20.json 2> gpurun_out/r2l_chr20.err
timeout 900 python bench.py --config human --coverage 3.75 --steps 2 --warmup 3 --no-files-e2e --no-cpu-baseline > gpurun_out/r2l_human.json 2> gpurun_out/r2l_human.err
python tools/cbf_bench.py > gpurun_out/r2l_cbf.json 2> gpurun_out/r2l_cbf.err
