#!/bin/bash
cd "$GRAFT_REPO_ROOT"
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
nvidia-smi topo -m 2>&1 | head -8
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/n2_err.log | tee gpurun_out/bench_n2.json | cut -c1-900
tail -5 gpurun_out/n2_err.log
