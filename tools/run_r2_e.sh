#!/bin/bash
# round 2, fifth GPU session: full parity suite; file e2e with the feeder's own timing (strip / device road)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_pytest.log
VG_FEEDER_DEBUG=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_chr20.json 2> gpurun_out/r2e_chr20.err; echo "rc=$?" >> gpurun_out/r2e_chr20.err
VG_FEEDER_DEBUG=1 VG_FASTQ_ROAD=device timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2e_chr20_devroad.json 2> gpurun_out/r2e_chr20_devroad.err
VG_FEEDER_DEBUG=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --buffer-mb 16 > gpurun_out/r2e_chr20_b16.json 2> gpurun_out/r2e_chr20_b16.err
nproc > gpurun_out/r2e_host.txt; free -g >> gpurun_out/r2e_host.txt; lscpu | head -25 >> gpurun_out/r2e_host.txt; df -h /dev/shm >> gpurun_out/r2e_host.txt; mount | grep shm >> gpurun_out/r2e_host.txt
tail -3 gpurun_out/r2e_pytest.log; tail -qn3 gpurun_out/r2e_chr20*.err
