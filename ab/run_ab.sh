#!/bin/bash
cd "$GRAFT_REPO_ROOT"
run() { python bench.py --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.4g e2e %.4g pass_ms %.2f P=%d'%(d['value'],d['e2e']['value'],d['roofline']['kernel_ms'],d['config']['table_partitions']))"; }
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
echo "== default"; run
echo "== capx 1.8"; VG_SCATTER_CAPX=1.8 run
echo "== capx 3"; VG_SCATTER_CAPX=3 run
echo "== P=188"; VG_SLICE_BYTES=8388608 run
