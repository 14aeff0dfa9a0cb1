#!/bin/bash
mkdir -p gpurun_out
VG_ROUND_DEBUG=1 timeout 900 python bench.py --config human --coverage 3.75 --steps 2 --warmup 3 --no-files-e2e --no-cpu-baseline > gpurun_out/r2k_human.json 2> gpurun_out/r2k_human.err; echo "rc=$?" >> gpurun_out/r2k_human.err
python tools/cbf_bench.py > gpurun_out/r2k_cbf.json 2> gpurun_out/r2k_cbf.err
timeout 600 python bench.py --steps 10 --warmup 3 --no-files-e2e --no-cpu-baseline > gpurun_out/r2k_chr20.json 2> gpurun_out/r2k_chr20.err
grep "vg round" gpurun_out/r2k_human.err | head -12; cat gpurun_out/r2k_cbf.json
