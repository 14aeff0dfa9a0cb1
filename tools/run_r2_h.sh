#!/bin/bash
# round 2, eighth GPU session: window encoder + arithmetic 2-bit encode (parity + speed), feeder with a deeper ring
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2h_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_pytest.log
VG_FEEDER_DEBUG=1 timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2h_chr20.json 2> gpurun_out/r2h_chr20.err; echo "rc=$?" >> gpurun_out/r2h_chr20.err
VG_FEEDER_DEBUG=1 VG_FASTQ_ROAD=strip timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_chr20_strip.json 2> gpurun_out/r2h_chr20_strip.err
VG_FEEDER_DEBUG=1 VG_FASTQ_ROAD=device timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2h_chr20_device.json 2> gpurun_out/r2h_chr20_device.err
VG_LIB=$PWD/varigraph_b200/libvgb200_rolling.so timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-files-e2e > gpurun_out/r2h_chr20_rolling.json 2> gpurun_out/r2h_chr20_rolling.err
VG_LIB=$PWD/varigraph_b200/libvgb200_lut.so timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-files-e2e > gpurun_out/r2h_chr20_lut.json 2> gpurun_out/r2h_chr20_lut.err
timeout 900 python bench.py --config human --coverage 3.75 --steps 2 --warmup 3 --no-files-e2e --no-cpu-baseline > gpurun_out/r2h_human.json 2> gpurun_out/r2h_human.err; echo "rc=$?" >> gpurun_out/r2h_human.err
timeout 600 python bench.py --kmer 21 --steps 5 --warmup 3 --no-files-e2e --no-cpu-baseline > gpurun_out/r2h_chr20_k21.json 2> gpurun_out/r2h_chr20_k21.err
tail -5 gpurun_out/r2h_pytest.log; grep -h "^FAILED\|^ERROR" gpurun_out/r2h_pytest.log | head; tail -qn3 gpurun_out/r2h_chr20.err gpurun_out/r2h_chr20_strip.err gpurun_out/r2h_chr20_device.err
