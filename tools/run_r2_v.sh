#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "count_files or gzip" > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2v_pytest.log
tail -3 gpurun_out/r2v_pytest.log
VG_FEEDER_DEBUG=1 VG_GZ_DEBUG=1 timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2v_bench.json 2> gpurun_out/r2v_bench.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2v_bench.json').read().strip().splitlines()[-1])
print('value %.4g e2e %.4g'%(d['value'], d['e2e']['value']))
print(json.dumps(d.get('e2e_gz'), indent=1))
PY
grep -E "vg_gzip|inflate" gpurun_out/r2v_bench.err | tail -6
timeout 1200 python tools/config5_timing.py 8 1000000 2000 30 > gpurun_out/r2v_config5.json 2> gpurun_out/r2v_config5.err; echo "config5 rc=$?"; cat gpurun_out/r2v_config5.json; tail -3 gpurun_out/r2v_config5.err
