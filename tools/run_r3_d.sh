#!/bin/bash
mkdir -p gpurun_out
timeout 280 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -x -k "span8 or sharded" > gpurun_out/r3d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3d_pytest.log
tail -5 gpurun_out/r3d_pytest.log
