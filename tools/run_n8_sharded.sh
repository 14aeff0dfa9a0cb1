#!/bin/bash
# the human-scale index cut over the box, bigger rounds (fewer all-to-all rounds per sample)
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
RM=${1:-4864}
timeout 1200 $T bench.py --gpus 8 --config human --coverage 30 --scaling strong --index sharded --round-mb $RM --steps 2 --warmup 3 --no-files-e2e --no-cpu-baseline > gpurun_out/n8s_human_sharded_rm$RM.json 2> gpurun_out/n8s_human_sharded_rm$RM.err; echo "rc=$?" >> gpurun_out/n8s_human_sharded_rm$RM.err
tail -n4 gpurun_out/n8s_human_sharded_rm$RM.err | cut -c1-300
python tools/show_bench.py gpurun_out/n8s_human_sharded_rm$RM.json
