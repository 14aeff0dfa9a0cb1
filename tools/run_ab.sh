#!/bin/bash
cd "$GRAFT_REPO_ROOT"
run() { python bench.py --no-cpu-baseline --steps 10 --warmup 3 "$@" 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.4g e2e %.4g ms %.2f pass_ms %.2f P=%d eq=%s h=%.3f frac=%.3f'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['config']['table_partitions'],d['e2e']['counts_equal_device_path'],d['roofline']['hit_fraction'],d['roofline']['frac']))"; }
echo "== human variant density (1 per 124 bp)"; run --variants 516000
echo "== k=21"; run --kmer 21
