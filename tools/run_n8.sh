#!/bin/bash
# one bench.py run on N GPUs (gpurun --gpus N -- 'bash tools/run_n8.sh N [bench flags]')
cd "$GRAFT_REPO_ROOT"
N=${1:-8}
shift
timeout 280 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 "$@" 2> gpurun_out/n${N}_err.log | tee -a gpurun_out/bench_n$N.jsonl | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('N=%d value %.4g e2e %.4g ms %.2f pass_ms %.2f | %s | eq=%s launches=%s'%(d['n_gpus'],d['value'],d['e2e']['value'],d['ms_per_step'],d['roofline']['kernel_ms'],d['config']['parallelism'],d['e2e']['counts_equal_device_path'],d['gpu_launches']))"
grep -v "OMP_NUM\|\*\*\*\*" gpurun_out/n${N}_err.log | tail -5 | cut -c1-300
