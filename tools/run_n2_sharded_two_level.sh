#!/bin/bash
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
VG_TWO_LEVEL_FROM=8 timeout 600 $T bench.py --gpus 2 --steps 2 --warmup 3 --index sharded --no-files-e2e > gpurun_out/n2t_chr20_sharded_two_level.json 2> gpurun_out/n2t_chr20_sharded_two_level.err; echo "rc=$?" >> gpurun_out/n2t_chr20_sharded_two_level.err
tail -n2 gpurun_out/n2t_chr20_sharded_two_level.err | cut -c1-300
python tools/show_bench.py gpurun_out/n2t_chr20_sharded_two_level.json
