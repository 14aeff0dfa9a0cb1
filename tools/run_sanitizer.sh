#!/bin/bash
# compute-sanitizer over a small slice of the parity suite: memcheck (out-of-bounds / misaligned), racecheck (shared memory hazards)
mkdir -p gpurun_out
S="tests/test_gpu_parity.py -m gpu -q -x -k"
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest $S "test_two_level_scatter_matches_oracle or test_partitioned_overflow_and_saturation or test_slot_order_result or test_count_golden_tiny or test_encoder_alignment_and_tails or test_cbf_golden" > gpurun_out/san_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/san_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 3 --print-limit 20 python __graft_entry__.py --smoke > gpurun_out/san_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/san_racecheck.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 3 --print-limit 20 python -m pytest tests/test_gpu_multi.py -m gpu -q -x -k "group_of_one" > gpurun_out/san_memcheck_sharded.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/san_memcheck_sharded.log
tail -4 gpurun_out/san_memcheck.log; tail -4 gpurun_out/san_racecheck.log; tail -4 gpurun_out/san_memcheck_sharded.log
