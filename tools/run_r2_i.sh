#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2i_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i_pytest.log
tail -4 gpurun_out/r2i_pytest.log; grep -h "^FAILED\|^ERROR" gpurun_out/r2i_pytest.log | head
VG_FEEDER_DEBUG=1 timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2i_chr20.json 2> gpurun_out/r2i_chr20.err; echo "rc=$?" >> gpurun_out/r2i_chr20.err
timeout 2400 bash tools/run_profile.sh r2_a
