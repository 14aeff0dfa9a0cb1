#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "gzip or count_files_golden or kseq_edge" > gpurun_out/r3b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3b_pytest.log
tail -3 gpurun_out/r3b_pytest.log
for pt in 4 2 8; do
VG_GZ_PER_THREAD=$pt VG_FEEDER_DEBUG=1 VG_GZ_DEBUG=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r3b_bench_pt$pt.json 2> gpurun_out/r3b_bench_pt$pt.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r3b_bench_pt$pt.json').read().strip().splitlines()[-1])
g=d['e2e_gz']; print('per_thread $pt: e2e_gz %.3f G/s zlib %.3f text %.2f GB/s equal %s'%(g['value']/1e9,g['zlib_road_value']/1e9,g['inflated_text_gb_per_s'],g['counts_equal_zlib_road']))
PY
grep -E "vg_gzip|gzip file" gpurun_out/r3b_bench_pt$pt.err | tail -4
done
