#!/usr/bin/env python
"""Print the figures of bench.py JSON lines side by side: tools/show_bench.py gpurun_out/r2b_*.json"""
import json, sys
for f in sys.argv[1:]:
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f"{f}: unreadable ({e})")
        continue
    if "config" not in d or "roofline" not in d:  # not a bench.py count line (reference arm, cbf_bench, ...)
        print(f"{f}: " + ", ".join(f"{k}={d[k]}" for k in ("impl", "value", "kmers_per_s", "unavailable") if k in d))
        continue
    c, r = d["config"], d["roofline"]
    ph = r.get("phases_ms") or {}
    print(f"{f}: {d['value'] / 1e9:7.2f} G/s  {d['ms_per_step']:8.2f} ms/step  P={c['table_partitions']} slices={c.get('table_slices')} "
          f"scatter {ph.get('scatter_ms_per_sample', 0):.2f} ms x{ph.get('scatter_launches')}  "
          f"sweep {ph.get('sweep_ms_per_sweep', 0):.2f} ms x{ph.get('sweeps')}  h={r['hit_fraction']:.3f} frac={r['frac']:.3f} "
          f"e2e={((d.get('e2e') or {}).get('value') or 0) / 1e9:.2f} gz={((d.get('e2e_gz') or {}).get('value') or 0) / 1e9:.2f} staged={((d.get('e2e_staged') or {}).get('value') or 0) / 1e9:.2f} parity={d.get('parity')}")
