#!/bin/bash
# round 2, second GPU session: parity suite with the two-level scatter, then the human-scale index on one GPU (3x coverage):
# two-level scatter / prefetch-ahead / slice size / DRAM-resident pre-filter (span 4 and 8)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
H="python bench.py --config human --coverage 3 --steps 2 --warmup 3 --no-files-e2e --no-cpu-baseline"
run() { name=$1; shift; env "$@" timeout 600 $H > gpurun_out/r2b_$name.json 2> gpurun_out/r2b_$name.err; echo "$name rc=$?" >> gpurun_out/r2b_$name.err; }
run A VG_DUMMY=1
run B VG_PREFETCH_AHEAD=1
run C VG_PREFETCH_AHEAD=1 VG_SLICE_BYTES=16777216
run D VG_PREFETCH_AHEAD=1 VG_SLICE_BYTES=16777216 VG_PREFILTER_BYTES=1200000000
run E VG_PREFETCH_AHEAD=1 VG_SLICE_BYTES=16777216 VG_PREFILTER_BYTES=1200000000 VG_LIB=$PWD/varigraph_b200/libvgb200_span8.so
run F VG_PREFETCH_AHEAD=1 VG_SLICE_BYTES=16777216 VG_PREFILTER_BYTES=600000000 VG_LIB=$PWD/varigraph_b200/libvgb200_span8.so
VG_PREFETCH_AHEAD=1 timeout 600 python bench.py --steps 10 --warmup 3 --no-files-e2e --no-cpu-baseline > gpurun_out/r2b_chr20_ahead.json 2> gpurun_out/r2b_chr20_ahead.err
tail -3 gpurun_out/r2b_pytest.log; tail -n1 gpurun_out/r2b_*.err
