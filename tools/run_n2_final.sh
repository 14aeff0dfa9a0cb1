#!/bin/bash
# two real GPUs, final library: multi-GPU tests (incl. the sharded index built from device keys), default bench, sharded bench
mkdir -p gpurun_out
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q > gpurun_out/n2f_pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n2f_pytest_multi.log
tail -3 gpurun_out/n2f_pytest_multi.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "two_gpus" > gpurun_out/n2f_pytest_host.log 2>&1; echo "pytest rc=$?" >> gpurun_out/n2f_pytest_host.log
tail -3 gpurun_out/n2f_pytest_host.log
timeout 900 $T bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/n2f_chr20_weak.json 2> gpurun_out/n2f_chr20_weak.err; echo "rc=$?" >> gpurun_out/n2f_chr20_weak.err
timeout 900 $T bench.py --gpus 2 --steps 5 --warmup 3 --index sharded --no-files-e2e > gpurun_out/n2f_chr20_sharded.json 2> gpurun_out/n2f_chr20_sharded.err; echo "rc=$?" >> gpurun_out/n2f_chr20_sharded.err
timeout 900 $T bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/n2f_ref.json 2> gpurun_out/n2f_ref.err; echo "rc=$?" >> gpurun_out/n2f_ref.err
tail -qn2 gpurun_out/n2f_*.err
python tools/show_bench.py gpurun_out/n2f_chr20_weak.json gpurun_out/n2f_chr20_sharded.json
head -c 400 gpurun_out/n2f_ref.json
