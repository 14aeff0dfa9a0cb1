#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "sharded" > gpurun_out/r3g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3g_pytest.log
tail -3 gpurun_out/r3g_pytest.log
