#!/bin/bash
cd "$GRAFT_REPO_ROOT"
timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q 2>&1 | tail -3
run() { timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 "$@" 2> gpurun_out/n2_err.log | tee -a gpurun_out/bench_n2.jsonl | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('value %.4g e2e %.4g ms %.2f pass_ms %.2f | %s | eq=%s launches=%s'%(d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['config']['parallelism'],d['e2e']['counts_equal_device_path'],d['gpu_launches']))"; grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/n2_err.log | tail -3 | cut -c1-300; }
rm -f gpurun_out/bench_n2.jsonl
echo "== p2p"; run
echo "== sharded"; run --index sharded
echo "== sharded 1024"; run --index sharded --round-mb 1024
