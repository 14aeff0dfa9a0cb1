#!/bin/bash
# round 2, third GPU session: retire with batched loads / no barrier, TMA-staged scatter A/B, human-scale variants, ncu launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
H="python bench.py --config human --coverage 3 --steps 2 --warmup 3 --no-files-e2e --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 600 $H > gpurun_out/r2c_$name.json 2> gpurun_out/r2c_$name.err; echo "$name rc=$?" >> gpurun_out/r2c_$name.err; }
run A VG_DUMMY=1
run B VG_PREFETCH_AHEAD=1
run D VG_PREFETCH_AHEAD=1 VG_SLICE_BYTES=16777216 VG_PREFILTER_BYTES=1200000000
run E VG_PREFETCH_AHEAD=1 VG_SLICE_BYTES=16777216 VG_PREFILTER_BYTES=1200000000 VG_LIB=$PWD/varigraph_b200/libvgb200_span8.so
run G VG_PREFETCH_AHEAD=1 VG_PREFILTER_BYTES=1200000000 VG_LIB=$PWD/varigraph_b200/libvgb200_span8.so
run I VG_PREFILTER_BYTES=1200000000 VG_LIB=$PWD/varigraph_b200/libvgb200_span8.so
H="$H --load-factor 0.5"
run L VG_PREFETCH_AHEAD=1 VG_PREFILTER_BYTES=1200000000 VG_LIB=$PWD/varigraph_b200/libvgb200_span8.so
C="python bench.py --steps 10 --warmup 3 --no-files-e2e --no-cpu-baseline"
timeout 600 $C > gpurun_out/r2c_chr20.json 2> gpurun_out/r2c_chr20.err
VG_SCATTER_TMA=1 timeout 600 $C > gpurun_out/r2c_chr20_tma.json 2> gpurun_out/r2c_chr20_tma.err
VG_PREFETCH_AHEAD=1 timeout 600 $C > gpurun_out/r2c_chr20_ahead.json 2> gpurun_out/r2c_chr20_ahead.err
VG_ROUND_KEYS=4294967296 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1100 -c 1080 --csv --log-file gpurun_out/r2c_ncu_human_launches.csv \
  python bench.py --config human --coverage 1 --steps 1 --warmup 3 --no-files-e2e --no-cpu-baseline > gpurun_out/r2c_ncu_human.log 2>&1
tail -3 gpurun_out/r2c_pytest.log; tail -qn1 gpurun_out/r2c_*.err
