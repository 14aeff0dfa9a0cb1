#!/bin/bash
cd "$GRAFT_REPO_ROOT"
run() { python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.4g e2e %.4g pass_ms %.2f P=%d'%(d['value'],d['e2e']['value'],d['roofline']['kernel_ms'],d['config']['table_partitions']))"; }
echo "== default"; run
echo "== capx 1.8"; VG_SCATTER_CAPX=1.8 run
echo "== P=94 warp"; VG_SLICE_BYTES=16777216 VG_SCATTER_FLAT_MINP=10000 run
echo "== P=94 flat"; VG_SLICE_BYTES=16777216 VG_SCATTER_FLAT_MINP=1 run
echo "== P=188 warp"; VG_SLICE_BYTES=8388608 VG_SCATTER_FLAT_MINP=10000 run
echo "== P=188 flat"; VG_SLICE_BYTES=8388608 run
echo "== P=376 flat"; VG_SLICE_BYTES=4194304 run
echo "== P=47 flat"; VG_SCATTER_FLAT_MINP=1 run
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
