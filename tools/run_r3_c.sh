#!/bin/bash
mkdir -p gpurun_out
B="--steps 5 --warmup 3 --no-cpu-baseline --no-gz-e2e"
for sh in 0.65 0.7 0.75 0.65 0.7 0.75; do
VG_STRIP_SHARE=$sh VG_FEEDER_DEBUG=1 timeout 600 python bench.py $B > gpurun_out/r3c_share$sh.json 2> gpurun_out/r3c_share$sh.err
echo "share=$sh $(grep 'block workers' gpurun_out/r3c_share$sh.err | tail -1 | sed 's/.*workers busy/busy/')"
python tools/show_bench.py gpurun_out/r3c_share$sh.json | sed 's/.*h=0.417//'
done
