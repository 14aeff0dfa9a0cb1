#!/bin/bash
mkdir -p gpurun_out
timeout 190 python -m pytest tests -m gpu -q -x > gpurun_out/r3h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3h_pytest.log
tail -3 gpurun_out/r3h_pytest.log
