#!/bin/bash
mkdir -p gpurun_out
B="--steps 5 --warmup 3 --no-cpu-baseline --no-gz-e2e"
for pop in 1 0; do
VG_POPULATE=$pop VG_FEEDER_DEBUG=1 timeout 600 python bench.py $B > gpurun_out/r2x_pop$pop.json 2> gpurun_out/r2x_pop$pop.err
grep "block workers" gpurun_out/r2x_pop$pop.err | tail -2
done
for road in strip device; do
VG_FASTQ_ROAD=$road VG_FEEDER_DEBUG=1 timeout 600 python bench.py $B > gpurun_out/r2x_$road.json 2> gpurun_out/r2x_$road.err
grep "block workers" gpurun_out/r2x_$road.err | tail -1
done
python tools/show_bench.py gpurun_out/r2x_*.json
