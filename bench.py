#!/usr/bin/env python
"""bench.py -- read k-mers/sec counted against the graph index (BASELINE.json's metric).

Workload (config.workload): BASELINE.json configs[1], "human chr20-shaped synthetic graph (64 Mb,
~1.5M variants), 30x PE150 on 1 B200", generated here from fixed seeds with torch (plumbing, not
the product): iid genome, SNV/MNP variants every ~43 bp, index = canonical k-mer hashes of every
reference- and alt-allele window overlapping a variant, reads drawn from two haplotypes with 0.3 %
substitution errors (1/5 of them N).  A *step* is one whole sample: zero the counters, count every
read k-mer of the sample against the index, produce the count vector.

  value     k-mer positions / s, reads already resident in HBM (CUDA events, max over ranks)
  e2e       the same through the host-buffer C-ABI call (vg_count_begin / vg_count_submit from
            pinned host memory / vg_count_end into host memory), H2D + D2H inside the timed region
  roofline  the fused count kernel against the measured HBM copy bandwidth (MEASURED_PEAKS.json)
            with B_alg = 1 + 32 + 32 h bytes per position, plus the measured
            random-32-byte-sector gather rate of the same box as a second denominator
  cpu_baseline  the reference's own FastqKmer::build_fastq_index (oracle/_ref) on a bounded sample
  --impl reference   times that reference CPU path alone, same config/metric/unit

N > 1 (torchrun): reads are sharded over ranks (each rank counts its own 30x sample: weak scaling),
the index is replicated, and the only exchange is one reduce of the count vectors per sample:
min(255, sum over ranks), exact.  By default every rank reads its peers' u8 vectors over
NVLink (vg_count_allreduce: CUDA IPC peer memory, a device-side barrier, no NCCL in the data path);
--reduce nccl uses an NCCL all-reduce of u32 instead (also the fallback if peer mapping fails).
--index sharded cuts the index itself over the ranks (the index > HBM layout): the scatter kernel
stores every k-mer straight into the owning GPU's key list over NVLink, rounds of --round-mb.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

K = 27
READ_LEN = 150
METRIC = "read k-mers/sec counted vs graph index"
UNIT = "kmer_positions/s"


# ------------------------------------------------------------------------------------------------
# synthetic workload (torch; identical for both arms)
# ------------------------------------------------------------------------------------------------
def t_hash64(x: torch.Tensor, mask: int) -> torch.Tensor:
    x = (torch.bitwise_not(x) + (x << 21)) & mask
    x = x ^ (x >> 24)
    x = (x + (x << 3) + (x << 8)) & mask
    x = x ^ (x >> 14)
    x = (x + (x << 2) + (x << 4)) & mask
    x = x ^ (x >> 28)
    x = (x + (x << 31)) & mask
    return x


def t_window_keys(codes: torch.Tensor, k: int) -> torch.Tensor:
    """codes uint8 [L] in 0..3 -> int64 [L-k+1]: hash64(canonical k-mer starting at m) << 8 | k."""
    L = codes.numel()
    c = codes.to(torch.int64)
    n = L - k + 1
    fwd = torch.zeros(n, dtype=torch.int64, device=codes.device)
    rev = torch.zeros(n, dtype=torch.int64, device=codes.device)
    for j in range(k):
        seg = c[k - 1 - j: L - j]
        fwd |= seg << (2 * j)
        rev |= (3 - seg) << (2 * (k - 1 - j))
    canon = torch.minimum(fwd, rev)
    del fwd, rev
    return (t_hash64(canon, (1 << (2 * k)) - 1) << 8) | k


def make_graph(dev, genome_len: int, nvar: int, seed: int):
    """-> (ref codes u8[L], alt codes u8[L], var_pos int64[nvar], var_len int64[nvar], keys int64 unique)"""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    L = genome_len
    ref = torch.randint(0, 4, (L,), generator=g, device=dev, dtype=torch.uint8)
    pos = torch.randint(1000, L - 1000, (int(nvar * 1.03),), generator=g, device=dev)
    pos = torch.unique(pos)[:nvar]
    # keep variants at least 10 bp apart so spans never overlap
    keep = torch.ones_like(pos, dtype=torch.bool)
    keep[1:] = (pos[1:] - pos[:-1]) >= 10
    pos = pos[keep]
    nv = pos.numel()
    kind = torch.rand(nv, generator=g, device=dev)
    vlen = torch.ones(nv, dtype=torch.int64, device=dev)
    mnp = kind < 0.15  # stand-in for indels: 2..8 bp substitutions (alt k-mers are novel either way)
    vlen[mnp] = torch.randint(2, 9, (int(mnp.sum()),), generator=g, device=dev)
    alt = ref.clone()
    for j in range(8):
        sel = vlen > j
        p = pos[sel] + j
        alt[p] = (ref[p] + torch.randint(1, 4, (p.numel(),), generator=g, device=dev, dtype=torch.uint8)) % 4
    n = L - K + 1
    d = torch.zeros(L + 2, dtype=torch.int32, device=dev)
    d.index_add_(0, (pos - K + 1).clamp_(min=0), torch.ones(nv, dtype=torch.int32, device=dev))
    d.index_add_(0, pos + vlen, -torch.ones(nv, dtype=torch.int32, device=dev))
    cover = torch.cumsum(d, 0)[:n] > 0
    del d
    kr = t_window_keys(ref, K)[cover]
    ka = t_window_keys(alt, K)[cover]
    keys = torch.unique(torch.cat([kr, ka]))
    del kr, ka, cover
    return ref, alt, pos, vlen, keys


def make_reads(dev, ref, alt, pos, vlen, coverage: float, seed: int) -> torch.Tensor:
    """PE150 from two haplotypes -> uint8 [nreads * 151] staged chunk ('read\\n' records) on `dev`."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    L = ref.numel()
    haps = []
    for _ in range(2):
        carry = torch.rand(pos.numel(), generator=g, device=dev) < 0.3
        d = torch.zeros(L + 1, dtype=torch.int32, device=dev)
        one = torch.ones(int(carry.sum()), dtype=torch.int32, device=dev)
        d.index_add_(0, pos[carry], one)
        d.index_add_(0, pos[carry] + vlen[carry], -one)
        m = torch.cumsum(d, 0)[:L] > 0
        haps.append(torch.where(m, alt, ref))
        del d, m
    hap2 = torch.stack(haps)  # [2, L]
    npairs = int(round(coverage * L / (2 * READ_LEN)))
    ascii_lut = torch.tensor([65, 67, 71, 84, 78], dtype=torch.uint8, device=dev)  # A C G T N
    out = torch.empty((2 * npairs, READ_LEN + 1), dtype=torch.uint8, device=dev)
    out[:, READ_LEN] = 10
    ar = torch.arange(READ_LEN, device=dev)
    B = 1 << 19
    for b0 in range(0, npairs, B):
        nb = min(B, npairs - b0)
        h = torch.randint(0, 2, (nb,), generator=g, device=dev)
        ins = torch.randint(300, 501, (nb,), generator=g, device=dev)
        start = (torch.rand(nb, generator=g, device=dev, dtype=torch.float64) * (L - 502)).to(torch.int64)
        base = h * L
        flat = hap2.reshape(-1)
        m1 = flat[(base + start)[:, None] + ar[None, :]]
        m2 = 3 - flat[(base + start + ins - 1)[:, None] - ar[None, :]]
        both = torch.cat([m1, m2.to(torch.uint8)])
        err = torch.rand(both.shape, generator=g, device=dev) < 0.003
        sub = torch.randint(0, 5, both.shape, generator=g, device=dev, dtype=torch.uint8)  # 1/5 -> N
        both = torch.where(err, sub, both)
        out[2 * b0: 2 * b0 + 2 * nb, :READ_LEN] = ascii_lut[both.to(torch.int64)]
        del m1, m2, both, err, sub
    return out.reshape(-1)


def write_fastq_sample(path: str, lines: np.ndarray, nreads: int) -> int:
    """First `nreads` reads of a staged chunk -> plain FASTQ; returns bases written."""
    rec = lines[: nreads * (READ_LEN + 1)].reshape(nreads, READ_LEN + 1)
    head = np.frombuffer(b"@r\n", dtype=np.uint8)
    plus = np.frombuffer(b"+\n", dtype=np.uint8)
    out = np.empty((nreads, 3 + READ_LEN + 1 + 2 + READ_LEN + 1), dtype=np.uint8)
    out[:, :3] = head
    out[:, 3:3 + READ_LEN + 1] = rec
    out[:, 3 + READ_LEN + 1: 3 + READ_LEN + 3] = plus
    out[:, 3 + READ_LEN + 3: -1] = ord("I")
    out[:, -1] = 10
    out.tofile(path)
    return nreads * READ_LEN


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(gpu_index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def mark(self):
        return time.time()

    def stop(self, t0: float, t1: float) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for (t, r) in self.rows if t0 <= t <= t1 + 0.2] or [r for (_, r) in self.rows]
        sm, smax, reasons = [], [], set()
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic(partitioned: bool):
    """DRAM bytes of one count pass from the committed ncu capture (profiles/traffic.json), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)
        return t["partitioned_pass_bytes" if partitioned else "direct_pass_bytes"]
    except Exception:
        return None


def dist_info():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def workload_name(a) -> str:
    return (f"chr20-shaped synthetic graph ({a.genome_mb} Mb, ~{a.variants / 1e6:.2f}M variants), "
            f"{a.coverage:g}x PE150, k={K}")


# ------------------------------------------------------------------------------------------------
# the reference arm: FastqKmer::build_fastq_index on the host cores
# ------------------------------------------------------------------------------------------------
def reference_run(a, keys_np: np.ndarray, lines_np: np.ndarray, steps: int, warmup: int, want_counts=False):
    from tests import oracle_binding as ob
    cores = os.cpu_count() or 1
    nreads = min(a.cpu_sample_reads, lines_np.size // (READ_LEN + 1))
    tmpdir = tempfile.mkdtemp(prefix="vgbench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    fq = os.path.join(tmpdir, "sample.fq")
    write_fastq_sample(fq, lines_np, nreads)
    sample = f"first {nreads} reads of the sample as plain FASTQ on tmpfs, full index ({keys_np.size} k-mers)"
    out = {}
    try:
        if ob.Reference.available:
            ref = ob.Reference()
            h = ref.map_create(keys_np, K)
            times, counts = [], None
            for i in range(warmup + steps):
                counts, rb, sec = ref.map_count_files(h, keys_np.size, [fq], cores, want_counts=want_counts)
                if i >= warmup:
                    times.append(sec)
            ref.map_destroy(h)
            kind = "reference"
        else:  # the C port, single thread
            orc = ob.Oracle()
            sub = lines_np[: nreads * (READ_LEN + 1)]
            times, counts = [], None
            for i in range(warmup + steps):
                t0 = time.perf_counter()
                counts, _, _ = orc.count_lines(keys_np, sub, K)
                if i >= warmup:
                    times.append(time.perf_counter() - t0)
            kind, cores = "port", 1
        orc = ob.Oracle()
        pos = oracle_positions(orc, keys_np, lines_np[: nreads * (READ_LEN + 1)])
        sec = float(np.mean(times))
        out = {"value": pos / sec, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
               "seconds_per_step": sec, "positions": pos, "counts": counts, "nreads": nreads}
    finally:
        try:
            os.remove(fq)
            os.rmdir(tmpdir)
        except OSError:
            pass
    return out


def oracle_positions(orc, keys_np, sub) -> int:
    """Emitted k-mer positions of a chunk (N-free reads give reads x 124; errors put N in some).  Plain numpy
    (the odd-k run-length rule of SURVEY appendix A.4); `orc` and `keys_np` are unused."""
    b = np.ascontiguousarray(sub)
    valid = np.isin(b, np.frombuffer(b"ACGTacgtUu", dtype=np.uint8))
    # run length of consecutive valid bytes ending at each byte; positions with run >= K emit (odd K)
    idx = np.arange(b.size, dtype=np.int64)
    last_bad = np.maximum.accumulate(np.where(~valid, idx, -1))
    return int(((idx - last_bad) >= K).sum())


# ------------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--genome-mb", type=int, default=64)
    ap.add_argument("--variants", type=int, default=1_500_000)
    ap.add_argument("--coverage", type=float, default=30.0)
    ap.add_argument("--cpu-sample-reads", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--buffer-mb", type=int, default=64)
    ap.add_argument("--load-factor", type=float, default=0.0)
    ap.add_argument("--kmer", type=int, default=27, help="k (BASELINE configs use the default 27)")
    ap.add_argument("--index", default="replicated", choices=["replicated", "sharded"],
                    help="N > 1: a replica of the index per GPU (default) or one index cut over the GPUs")
    ap.add_argument("--reduce", default="p2p", choices=["p2p", "nccl"], help="N > 1, replicated: how counts are combined")
    ap.add_argument("--round-mb", type=int, default=1024, help="sharded index: bases per rank and round")
    a = ap.parse_args()
    global K
    K = a.kmer
    rank, world, local = dist_info()
    steps, warmup = a.steps, max(a.warmup, 3 if a.impl == "b200" else a.warmup)
    L = a.genome_mb * 1_000_000

    if a.impl == "reference":
        if rank != 0:
            return
        dev = torch.device("cuda", local) if torch.cuda.is_available() else torch.device("cpu")
        ref, alt, pos, vlen, keys = make_graph(dev, L, a.variants, seed=20261017)
        keys_np = keys.cpu().numpy().view(np.uint64)
        nreads = a.cpu_sample_reads
        cov = nreads * READ_LEN / L * 1.02 + 0.01
        lines_np = make_reads(dev, ref, alt, pos, vlen, min(cov, a.coverage), seed=1000).cpu().numpy()
        del ref, alt, pos, vlen, keys
        r = reference_run(a, keys_np, lines_np, steps, a.warmup)
        line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
                "warmup": a.warmup, "ms_per_step": r["seconds_per_step"] * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "impl": "reference",
                "config": {"workload": workload_name(a), "index_kmers": int(keys_np.size)},
                "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"],
                                 "sample": r["sample"]},
                "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line), flush=True)
        return

    # ---- our arm -----------------------------------------------------------------------------
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; varigraph_b200 has no CPU path (use --impl reference)")
    import torch.distributed as dist
    from varigraph_b200 import capi
    from varigraph_b200 import dist as vdist
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa = None
    if world > 1:
        numa = vdist.bind_to_gpu_numa(local)  # before any pinned allocation: staging memory next to the GPU
        dist.init_process_group("nccl", device_id=dev)

    ref, alt, pos, vlen, keys = make_graph(dev, L, a.variants, seed=20261017)
    keys_np = keys.cpu().numpy().view(np.uint64)
    nkeys = int(keys_np.size)
    del keys
    lines_dev = make_reads(dev, ref, alt, pos, vlen, a.coverage, seed=1000 + rank)
    del ref, alt, pos, vlen
    torch.cuda.empty_cache()
    nbytes = lines_dev.numel()
    lines_host = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    lines_host.copy_(lines_dev)
    torch.cuda.synchronize()

    ctx = capi.Context(local, buffer_mb=a.buffer_mb)
    stream = torch.cuda.Stream(device=dev)  # explicit: the legacy default stream's handle is 0 (= "unset")
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    # ---- N > 1: map the peers' memory (handles travel through torch.distributed, data never does) ----
    sharded = world > 1 and a.index == "sharded"
    comm, reduce_how = None, ("nccl" if world > 1 else None)
    if world > 1 and (sharded or a.reduce == "p2p"):
        def exchange(mine: bytes):
            got = [None] * world
            dist.all_gather_object(got, mine)
            return got
        round_bytes = a.round_mb << 20
        table_est = int(nkeys / world / (4 * (a.load_factor or 0.3)) * 32 * 1.1) + (64 << 20)
        arena = (nkeys + (1 << 20)) + ((table_est + round_bytes * 10 + (128 << 20)) if sharded else 0)
        ok = torch.ones(1, device=dev)
        try:
            comm = capi.Comm(ctx, rank, world, arena, exchange)
        except capi.VgError as ex:
            sys.stderr.write(f"rank {rank}: peer mapping failed ({ex})\n")
            ok.zero_()
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        if float(ok.item()) == 0.0:
            if sharded:
                raise SystemExit("bench.py: --index sharded needs CUDA IPC peer mappings")
            comm = None
        else:
            reduce_how = "p2p"
    ix = capi.Index(ctx, keys_np, K, a.load_factor, comm=comm if sharded else None,
                    round_bytes=(a.round_mb << 20) if sharded else 0)
    out32 = torch.empty(max(nkeys, 1), dtype=torch.int32, device=dev)
    out8 = torch.empty(max(nkeys, 1) + 32, dtype=torch.uint8, device=dev)
    counts_host = torch.empty(max(nkeys, 1), dtype=torch.uint8, pin_memory=True)
    # sharded index: the sample goes in rounds of at most --round-mb, cut at read boundaries
    rec = READ_LEN + 1
    per_round = max(1, ((a.round_mb << 20) - (1 << 20)) // rec) * rec
    round_cuts = [(o, min(per_round, nbytes - o)) for o in range(0, nbytes, per_round)] if sharded else []

    def device_step(kev=None, want=False):
        ix.begin()
        if kev:
            kev[0].record(stream)
        if sharded:
            for o, ln in round_cuts:
                ix.submit_device(lines_dev.data_ptr() + o, ln)
                ix.flush()
            if kev:
                kev[1].record(stream)
            return ix.end(want_counts=want)[0]  # collective: counts of all keys on every rank (device; host if asked)
        ix.submit_device(lines_dev.data_ptr(), nbytes)
        ix.flush()
        if kev:
            kev[1].record(stream)
        if comm is not None:  # every rank sums its peers' u8 vectors over NVLink
            comm.allreduce_counts(ix, want_host=False, dev_out=out8.data_ptr())
            return out8[:max(nkeys, 1)]
        if world > 1:  # --reduce nccl: all-reduce of u32, clamp to 255
            ix.extract_device(out32.data_ptr(), 4)
            return vdist.reduce_counts(out32)
        ix.extract_device(out8.data_ptr(), 1)  # the sample's result: u8 counts in key order, on the device
        return out8[:max(nkeys, 1)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        device_step()
    barrier()
    positions, hits = ix.stats()
    launches0 = ix.launches
    clocks = ClockSampler(local)
    kevs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_mark0 = clocks.mark()
    e0.record(stream)
    for i in range(steps):
        device_step(kevs[i])
    e1.record(stream)
    barrier()
    t_mark1 = clocks.mark()
    launches_timed = ix.launches - launches0 + steps  # + the extract kernel of each step
    if comm is not None:
        launches_timed += comm.launches * steps // (steps + warmup)
    total_ms = e0.elapsed_time(e1)
    kernel_ms = float(np.mean([k0.elapsed_time(k1) for k0, k1 in kevs]))
    tm = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    tp = torch.tensor([float(positions)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tm, op=dist.ReduceOp.MAX)
        dist.all_reduce(tp, op=dist.ReduceOp.SUM)
    ms_per_step = float(tm.item()) / steps
    total_positions = float(tp.item())
    value = total_positions / (ms_per_step * 1e-3)
    clk = clocks.stop(t_mark0, t_mark1)
    dc = device_step(want=True)
    device_counts = dc if isinstance(dc, np.ndarray) else dc.cpu().numpy()

    # ---- e2e: host buffers through the C ABI, copies inside the timed region -----------------
    def e2e_step():
        ix.begin()
        if sharded:
            for o, ln in round_cuts:
                ix.submit_ptr(lines_host.data_ptr() + o, ln)
                ix.flush()
            capi._chk(capi.lib.vg_count_end(ix._h, counts_host.data_ptr(), None, None))
            return
        ix.submit_ptr(lines_host.data_ptr(), nbytes)
        if comm is not None:
            capi._chk(capi.lib.vg_count_allreduce(comm._h, ix._h, counts_host.data_ptr(), None))
            capi._chk(capi.lib.vg_count_end(ix._h, None, None, None))
        elif world > 1:
            ix.extract_device(out32.data_ptr(), 4, stream.cuda_stream)  # syncs the context streams first
            counts_host.copy_(vdist.reduce_counts(out32), non_blocking=True)
            torch.cuda.synchronize()
            capi.lib.vg_count_end(ix._h, None, None, None)
        else:
            capi._chk(capi.lib.vg_count_end(ix._h, counts_host.data_ptr(), None, None))

    ctx.set_stream(0)  # the staged path pipelines copy and compute on the context's own streams
    for _ in range(2):
        e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    barrier()
    e2e_s = (time.perf_counter() - t0) / steps
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = total_positions / float(te.item())
    e2e_counts_ok = bool(np.array_equal(counts_host.numpy()[:nkeys], device_counts[:nkeys]))

    # ---- roofline of the fused count kernel --------------------------------------------------
    h = hits / max(positions, 1)
    b_alg = 1.0 + 32.0 + 32.0 * h
    achieved = positions * b_alg / (kernel_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak_gbs()
    rnd_gbs = rnd_sec = None
    if rank == 0:
        try:
            free, _ = torch.cuda.mem_get_info()
            tb = min(8 << 30, int(free * 0.5))
            rnd_gbs, rnd_sec = ctx.probe_random_sectors(tb, 64)
        except Exception as ex:  # diagnostic only
            rnd_gbs = rnd_sec = None
            sys.stderr.write(f"random-sector probe failed: {ex}\n")
    probes_per_s = positions / (kernel_ms * 1e-3)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(ix.partitions > 0), "peak_source": peak_src, "kernel": ("count pass = vg::scatter_kernel + vg::probe_slice_kernel sweep (K1 | K2+K3)" if ix.partitions
                           else "vg::count_kernel (K1+K2+K3 fused)"),
                "kernel_ms": kernel_ms, "bytes_per_position": b_alg, "hit_fraction": h,
                "random_sector_peak_gbs": rnd_gbs,
                "frac_of_random_sector_peak": (probes_per_s * (1 + h) / rnd_sec) if rnd_sec else None}

    # ---- CPU baseline beside it (rank 0, N == 1) ---------------------------------------------
    cpu = None
    parity = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        lines_np = lines_host.numpy()
        r = reference_run(a, keys_np, lines_np, steps=1, warmup=0, want_counts=True)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
        # the same sample through the CUDA path must give the reference's counts exactly
        ix.begin()
        ix.submit(lines_np[: r["nreads"] * (READ_LEN + 1)])
        mine, mp, _ = ix.end()
        parity = {"sample_counts_bit_exact": bool(np.array_equal(mine, r["counts"])),
                  "sample_positions_equal": bool(mp == r["positions"])}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u64", "data": "synthetic",
                "config": {"workload": workload_name(a), "index_kmers": nkeys, "index_table_bytes": ix.table_bytes, "table_partitions": ix.partitions,
                           "reads_per_gpu": nbytes // (READ_LEN + 1), "positions_per_gpu": positions,
                           "parallelism": (f"reads sharded x{world}, index "
                                           + (f"sharded x{world} (k-mer all-to-all fused into the scatter over NVLink, "
                                              f"{len(round_cuts)} rounds)" if sharded else f"replicated, counts reduced by {reduce_how}")
                                           if world > 1 else "1 GPU"),
                           "l2": "inputs (reads + index table) far larger than the 126 MB L2; no explicit flush",
                           "host_binding": numa},
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(nbytes),
                        "d2h_bytes_per_step": int(nkeys + 16), "counts_equal_device_path": e2e_counts_ok},
                "gpu_launches": int(launches_timed) + steps, "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
                "parity": parity}
        print(json.dumps(line), flush=True)
    if world > 1:
        barrier()  # nobody unmaps a peer's arena while it is still in use
    ix.close()
    if comm is not None:
        comm.close()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
