#!/bin/bash
cd "$GRAFT_REPO_ROOT"
N=${1:-8}
shift
run() { timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 "$@" 2> gpurun_out/n${N}_err.log | tee -a gpurun_out/bench_n$N.jsonl | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('N=%d value %.4g e2e %.4g ms %.2f pass_ms %.2f | %s | eq=%s launches=%s | %s'%(d['n_gpus'],d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['config']['parallelism'],d['e2e']['counts_equal_device_path'],d['gpu_launches'],d['config'].get('host_binding')))"
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/n${N}_err.log | tail -5 | cut -c1-300; }
rm -f gpurun_out/bench_n$N.jsonl
lscpu | grep -i "numa\|^CPU(s)" | head -6
run
run --index sharded
