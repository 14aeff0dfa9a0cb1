#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2z_pytest.log
tail -4 gpurun_out/r2z_pytest.log
python tools/cbf_bench.py > gpurun_out/r2z_cbf.json 2> gpurun_out/r2z_cbf.err; tail -c 600 gpurun_out/r2z_cbf.json
