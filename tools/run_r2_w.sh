#!/bin/bash
mkdir -p gpurun_out
B="--steps 5 --warmup 3 --no-files-e2e --no-cpu-baseline"
for fb in 0 44000000 59108096 88000000 118216192; do
  if [ $fb = 0 ]; then timeout 600 python bench.py $B > gpurun_out/r2w_pf_default.json 2>/dev/null
  else VG_PREFILTER_BYTES=$fb timeout 600 python bench.py $B > gpurun_out/r2w_pf_$fb.json 2>/dev/null; fi
done
python tools/show_bench.py gpurun_out/r2w_pf_*.json
python - <<'PY'
import json,glob
for f in sorted(glob.glob('gpurun_out/r2w_pf_*.json')):
    d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, d['roofline']['binding']['keys_after_prefilter'], d['roofline']['phases_ms'])
PY
