#!/bin/bash
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
run() {
timeout 900 $T bench.py --gpus 8 --config human --coverage 30 --scaling strong --index sharded --round-mb $1 --steps 2 --warmup 3 --no-files-e2e > gpurun_out/n8p_human_sharded_rm$1.json 2> gpurun_out/n8p_human_sharded_rm$1.err
rc=$?; echo "rc=$rc" >> gpurun_out/n8p_human_sharded_rm$1.err
grep -E "VgError|rc=" gpurun_out/n8p_human_sharded_rm$1.err | tail -3 | cut -c1-300
python tools/show_bench.py gpurun_out/n8p_human_sharded_rm$1.json
return $rc
}
run 11100 || run 9100
