#!/bin/bash
# round 2, first GPU session: parity suite, smoke, default bench, human-scale baseline (one GPU, 3x coverage)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2a_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2a_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?" >> gpurun_out/r2a_bench.err
timeout 900 python bench.py --config human --coverage 3 --steps 2 --warmup 3 --no-files-e2e > gpurun_out/r2a_human.json 2> gpurun_out/r2a_human.err; echo "human rc=$?" >> gpurun_out/r2a_human.err
VG_PREFILTER_BYTES=1500000000 timeout 900 python bench.py --config human --coverage 3 --steps 2 --warmup 3 --no-files-e2e --no-cpu-baseline > gpurun_out/r2a_human_pf.json 2> gpurun_out/r2a_human_pf.err; echo "human_pf rc=$?" >> gpurun_out/r2a_human_pf.err
tail -3 gpurun_out/r2a_pytest.log; tail -3 gpurun_out/r2a_smoke.log; tail -2 gpurun_out/r2a_bench.err; tail -2 gpurun_out/r2a_human.err; tail -2 gpurun_out/r2a_human_pf.err
