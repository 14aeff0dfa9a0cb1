#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "span8" > gpurun_out/r3e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3e_pytest.log
tail -3 gpurun_out/r3e_pytest.log
