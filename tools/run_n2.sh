#!/bin/bash
# two real GPUs: gpurun --gpus 2 -- 'bash tools/run_n2.sh'
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi topo -m > gpurun_out/n2_topo.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_parity.py -m gpu -q -k "two_ranks or two_gpus or sharded_group or independent" > gpurun_out/n2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n2_pytest.log
tail -3 gpurun_out/n2_pytest.log
timeout 900 $T bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2_chr20_weak.json 2> gpurun_out/n2_chr20_weak.err; echo "rc=$?" >> gpurun_out/n2_chr20_weak.err
timeout 900 $T bench.py --gpus 2 --steps 10 --warmup 3 --scaling strong --no-files-e2e > gpurun_out/n2_chr20_strong.json 2> gpurun_out/n2_chr20_strong.err; echo "rc=$?" >> gpurun_out/n2_chr20_strong.err
timeout 900 $T bench.py --gpus 2 --steps 5 --warmup 3 --index sharded --no-files-e2e > gpurun_out/n2_chr20_sharded.json 2> gpurun_out/n2_chr20_sharded.err; echo "rc=$?" >> gpurun_out/n2_chr20_sharded.err
timeout 1200 $T bench.py --gpus 2 --config human --coverage 7.5 --scaling strong --steps 2 --warmup 3 --no-files-e2e > gpurun_out/n2_human_strong.json 2> gpurun_out/n2_human_strong.err; echo "rc=$?" >> gpurun_out/n2_human_strong.err
tail -qn2 gpurun_out/n2_*.err
