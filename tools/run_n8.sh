#!/bin/bash
# the whole box: gpurun --gpus 8 -- 'bash tools/run_n8.sh'
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
nvidia-smi topo -m > gpurun_out/n8_topo.txt 2>&1; nproc >> gpurun_out/n8_topo.txt; free -g >> gpurun_out/n8_topo.txt
# 1. what the driver runs: the chr20 line (weak scaling) + the human-scale section (one 30x sample over the box)
timeout 900 $T bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/n8_default.json 2> gpurun_out/n8_default.err; echo "rc=$?" >> gpurun_out/n8_default.err
# 2. one chr20 sample cut over the box
timeout 600 $T bench.py --gpus 8 --steps 10 --warmup 3 --scaling strong --no-files-e2e --human off > gpurun_out/n8_chr20_strong.json 2> gpurun_out/n8_chr20_strong.err; echo "rc=$?" >> gpurun_out/n8_chr20_strong.err
# 3. the human-scale index cut over the box (index > HBM layout), one 30x sample
timeout 900 $T bench.py --gpus 8 --config human --coverage 30 --scaling strong --index sharded --steps 2 --warmup 3 --no-files-e2e --no-cpu-baseline > gpurun_out/n8_human_sharded.json 2> gpurun_out/n8_human_sharded.err; echo "rc=$?" >> gpurun_out/n8_human_sharded.err
tail -qn2 gpurun_out/n8_*.err
