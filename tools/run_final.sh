#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r3f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3f_pytest.log
tail -3 gpurun_out/r3f_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r3f_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r3f_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r3f_bench.json 2> gpurun_out/r3f_bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3f_ref.json 2> gpurun_out/r3f_ref.err; echo "ref rc=$?"
nproc; free -g | head -2
