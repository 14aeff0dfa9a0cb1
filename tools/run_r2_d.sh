#!/bin/bash
# round 2, fourth GPU session: strip road, adaptive rounds, DRAM pre-filter by default, replica tests; ncu launch list (human)
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2d_chr20.json 2> gpurun_out/r2d_chr20.err; echo "rc=$?" >> gpurun_out/r2d_chr20.err
VG_FASTQ_ROAD=device timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2d_chr20_devroad.json 2> gpurun_out/r2d_chr20_devroad.err
H="python bench.py --config human --coverage 3.75 --steps 2 --warmup 3 --no-files-e2e"
timeout 900 $H > gpurun_out/r2d_human.json 2> gpurun_out/r2d_human.err; echo "rc=$?" >> gpurun_out/r2d_human.err
timeout 900 $H --load-factor 0.5 --no-cpu-baseline > gpurun_out/r2d_human_lf5.json 2> gpurun_out/r2d_human_lf5.err; echo "rc=$?" >> gpurun_out/r2d_human_lf5.err
VG_ROUND_KEYS=4294967296 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'scatter_kernel|probe_slice_kernel|rescatter_kernel|sum_cursors' --launch-skip 2074 -c 1037 --csv --log-file gpurun_out/r2d_ncu_human_launches.csv \
  python bench.py --config human --coverage 1 --steps 1 --warmup 3 --no-files-e2e --no-cpu-baseline > gpurun_out/r2d_ncu_human.log 2>&1
tail -3 gpurun_out/r2d_pytest.log; tail -qn1 gpurun_out/r2d_*.err
