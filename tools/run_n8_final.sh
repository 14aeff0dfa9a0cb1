#!/bin/bash
# the whole box, final library: what the driver runs (chr20 weak + the human-scale section)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
timeout 1200 $T bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/n8f_default.json 2> gpurun_out/n8f_default.err; echo "rc=$?" >> gpurun_out/n8f_default.err
tail -n3 gpurun_out/n8f_default.err
python tools/show_bench.py gpurun_out/n8f_default.json
