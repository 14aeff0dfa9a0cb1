#!/bin/bash
# round 2, seventh GPU session: full parity suite; hybrid FASTQ road at several strip shares
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
for sh in 0.6 0.5 0.7 0.4; do
VG_FEEDER_DEBUG=1 VG_STRIP_SHARE=$sh timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2g_chr20_s$sh.json 2> gpurun_out/r2g_chr20_s$sh.err; echo "rc=$?" >> gpurun_out/r2g_chr20_s$sh.err
done
tail -5 gpurun_out/r2g_pytest.log; tail -qn2 gpurun_out/r2g_chr20*.err
