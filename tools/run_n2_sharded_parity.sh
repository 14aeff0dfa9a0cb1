#!/bin/bash
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $T bench.py --gpus 2 --steps 3 --warmup 3 --index sharded --no-files-e2e > gpurun_out/n2p_chr20_sharded.json 2> gpurun_out/n2p_chr20_sharded.err; echo "rc=$?" >> gpurun_out/n2p_chr20_sharded.err
tail -n3 gpurun_out/n2p_chr20_sharded.err | cut -c1-300
python tools/show_bench.py gpurun_out/n2p_chr20_sharded.json
