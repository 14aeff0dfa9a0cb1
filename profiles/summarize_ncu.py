#!/usr/bin/env python
"""Turn ncu outputs (gpurun_out/*.ncu-rep, launch-list CSVs) into the small text summaries kept here.

  python profiles/summarize_ncu.py rep  <file.ncu-rep> <out.txt> "<title>"
  python profiles/summarize_ncu.py list <launches.csv> <out.txt> "<title>" [last_n]   (only the last n launches)
"""
import collections
import csv
import subprocess
import sys

RAW_KEEP = ("gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__sectors_read.sum",
            "dram__sectors_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
            "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum", "lts__t_sectors_op_read.sum",
            "lts__t_sectors_op_write.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
            "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_requests_pipe_lsu_mem_global_op_atom.sum",
            "l1tex__t_requests_pipe_lsu_mem_global_op_red.sum", "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum",
            "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
            "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
            "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
            "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__waves_per_multiprocessor",
            "sm__cycles_elapsed.max", "smsp__cycles_active.avg")


def rep(path, out, title):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    src = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    with open(out, "w") as f:
        f.write(title + "\n(ncu --set full --clock-control none --import-source on; cold-cache, serialised)\n\n")
        for vals in rows[2:]:
            name = vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write(f"== {name}\n")
            for i, h in enumerate(hdr):
                if h in RAW_KEEP:
                    f.write(f"{h:72s} {units[i]:16s} {vals[i]}\n")
        if len(srows) > 2:
            shdr, data = srows[1], srows[2:]
            ix = {h: i for i, h in enumerate(shdr)}
            stalls = [h for h in shdr if h.startswith("stall_") and "Not Issued" not in h]
            tot = collections.Counter()
            for r in data:
                for s in stalls:
                    try:
                        tot[s] += int(r[ix[s]])
                    except (ValueError, IndexError):
                        pass
            T = sum(tot.values()) or 1
            f.write(f"\nwarp stall samples: {T} over {len(data)} SASS instructions\n")
            for s, v in tot.most_common(8):
                f.write(f"  {s:28s} {100 * v / T:5.1f}%\n")
            f.write("\nhottest instructions (samples, share, SASS, dominant stall)\n")
            for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:16]:
                st = {s: int(r[ix[s]] or 0) for s in stalls}
                best = max(st, key=st.get)
                n = int(r[ix["# Samples"]] or 0)
                f.write(f"  {n:7d} {100 * n / T:5.1f}%  {r[ix['Source']].strip()[:70]:70s} {best}\n")


def launches(path, out, title, last_n=0):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    i_name, i_m, i_v, i_id = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    d = collections.OrderedDict()
    for r in rows[1:]:
        d.setdefault(r[i_id], {"name": r[i_name].split("(")[0][-48:]})[r[i_m]] = float(r[i_v].replace(",", ""))
    agg = collections.OrderedDict()
    for v in list(d.values())[-last_n:]:
        a = agg.setdefault(v["name"], collections.Counter())
        a["n"] += 1
        for k, x in v.items():
            if k != "name":
                a[k] += x
    total = sum(a["gpu__time_duration.sum"] for a in agg.values()) or 1
    with open(out, "w") as f:
        f.write(title + "\n(ncu --metrics gpu__time_duration.sum,... --clock-control none; per-launch times are cold-cache "
                        "and serialised: compare SHARES)\n\n")
        f.write(f"{'kernel':50s} {'launches':>8s} {'ms':>10s} {'share':>7s} {'dram rd MB':>11s} {'dram wr MB':>11s}\n")
        for name, a in agg.items():
            f.write(f"{name:50s} {a['n']:8d} {a['gpu__time_duration.sum'] / 1e6:10.3f} "
                    f"{100 * a['gpu__time_duration.sum'] / total:6.1f}% {a.get('dram__bytes_read.sum', 0) / 1e6:11.1f} "
                    f"{a.get('dram__bytes_write.sum', 0) / 1e6:11.1f}\n")


if __name__ == "__main__":
    mode, path, out, title = sys.argv[1:5]
    if mode == "rep":
        rep(path, out, title)
    else:
        launches(path, out, title, int(sys.argv[5]) if len(sys.argv) > 5 else 0)
