#!/bin/bash
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $T bench.py --gpus 8 --config human --coverage 30 --scaling strong --index sharded --round-mb 4608 --steps 2 --warmup 3 --no-files-e2e --no-cpu-baseline > gpurun_out/n8s_human_sharded_rm4608.json 2> gpurun_out/n8s_human_sharded_rm4608.err; echo "rc=$?" >> gpurun_out/n8s_human_sharded_rm4608.err
python tools/show_bench.py gpurun_out/n8s_human_sharded_rm4608.json
timeout 900 $T bench.py --gpus 8 --config human --coverage 30 --scaling strong --index sharded --round-mb 9100 --steps 2 --warmup 3 --no-files-e2e > gpurun_out/n8s_human_sharded_rm9100.json 2> gpurun_out/n8s_human_sharded_rm9100.err; echo "rc=$?" >> gpurun_out/n8s_human_sharded_rm9100.err
tail -n3 gpurun_out/n8s_human_sharded_rm9100.err | cut -c1-300
python tools/show_bench.py gpurun_out/n8s_human_sharded_rm9100.json
