#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
tail -4 gpurun_out/r2j_pytest.log; grep -h "^FAILED\|^ERROR" gpurun_out/r2j_pytest.log | head
VG_FEEDER_DEBUG=1 timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2j_chr20.json 2> gpurun_out/r2j_chr20.err; echo "rc=$?" >> gpurun_out/r2j_chr20.err
timeout 900 python bench.py --config human --coverage 3.75 --steps 2 --warmup 3 --no-files-e2e --no-cpu-baseline > gpurun_out/r2j_human.json 2> gpurun_out/r2j_human.err; echo "rc=$?" >> gpurun_out/r2j_human.err
timeout 900 python bench.py --config human --coverage 0.05 --steps 2 --warmup 3 --no-files-e2e --no-cpu-baseline > gpurun_out/r2j_human_empty.json 2> gpurun_out/r2j_human_empty.err; echo "rc=$?" >> gpurun_out/r2j_human_empty.err
VG_PREFETCH_AHEAD=1 timeout 900 python bench.py --config human --coverage 3.75 --steps 2 --warmup 3 --no-files-e2e --no-cpu-baseline > gpurun_out/r2j_human_ahead.json 2> gpurun_out/r2j_human_ahead.err
timeout 600 python bench.py --kmer 28 --steps 3 --warmup 3 --no-files-e2e --no-cpu-baseline > gpurun_out/r2j_chr20_k28.json 2> gpurun_out/r2j_chr20_k28.err
tail -qn2 gpurun_out/r2j_*.err
