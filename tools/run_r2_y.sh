#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2y_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2y_pytest.log
tail -4 gpurun_out/r2y_pytest.log
python tools/cbf_bench.py > gpurun_out/r2y_cbf.json 2> gpurun_out/r2y_cbf.err; tail -c 600 gpurun_out/r2y_cbf.json
B="--steps 5 --warmup 3 --no-files-e2e --no-cpu-baseline"
timeout 600 python bench.py $B > gpurun_out/r2y_tiles1.json 2>/dev/null
VG_SCATTER_TILES=2 timeout 600 python bench.py $B > gpurun_out/r2y_tiles2.json 2>/dev/null
VG_SCATTER_TILES=2 VG_SCATTER_CAPX=1.5 timeout 600 python bench.py $B > gpurun_out/r2y_tiles2_capx15.json 2>/dev/null
python tools/show_bench.py gpurun_out/r2y_tiles*.json
