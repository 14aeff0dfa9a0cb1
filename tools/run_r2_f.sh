#!/bin/bash
# round 2, sixth GPU session (1 GPU): the multi-rank tests on a shared device + the host binary's --gpu list
mkdir -p gpurun_out
timeout 1800 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -x -q -k "two_ranks or two_gpus or sharded or replica or genotype or slot_order" > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -5 gpurun_out/r2f_pytest.log
