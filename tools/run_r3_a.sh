#!/bin/bash
mkdir -p gpurun_out
B="--steps 5 --warmup 3 --no-cpu-baseline --no-gz-e2e"
for fx in 1 0 1 0; do
VG_STRIP_FIXED=$fx VG_FEEDER_DEBUG=1 timeout 600 python bench.py $B > gpurun_out/r3a_fixed$fx.json 2> gpurun_out/r3a_fixed$fx.err
echo "fixed=$fx"; grep "block workers" gpurun_out/r3a_fixed$fx.err | tail -2
python tools/show_bench.py gpurun_out/r3a_fixed$fx.json
done
VG_FASTQ_ROAD=strip VG_STRIP_FIXED=1 VG_FEEDER_DEBUG=1 timeout 600 python bench.py $B > gpurun_out/r3a_strip_fixed1.json 2> gpurun_out/r3a_strip_fixed1.err
grep "block workers" gpurun_out/r3a_strip_fixed1.err | tail -1; python tools/show_bench.py gpurun_out/r3a_strip_fixed1.json
lscpu | grep -E "Model name|Socket|Core|Thread|L3|NUMA" | head -8
