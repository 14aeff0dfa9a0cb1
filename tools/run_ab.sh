#!/bin/bash
cd "$GRAFT_REPO_ROOT"
run() { python bench.py --no-cpu-baseline --steps 10 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.4g e2e %.4g ms %.2f pass_ms %.2f P=%d eq=%s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['config']['table_partitions'],d['e2e']['counts_equal_device_path']))"; }
echo "== buffer 16"; run --buffer-mb 16
echo "== buffer 32"; run --buffer-mb 32
echo "== buffer 128"; run --buffer-mb 128
