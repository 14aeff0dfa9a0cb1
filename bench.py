#!/usr/bin/env python
"""bench.py -- read k-mers/sec counted against the graph index (BASELINE.json's metric).

Workloads (config.workload), generated here from fixed seeds with torch (plumbing, not the product):
  chr20 (default)  BASELINE.json configs[1]: "human chr20-shaped synthetic graph (64 Mb, ~1.5M variants), 30x PE150
                   on 1 B200" -- iid genome, SNV/MNP variants every ~43 bp
  human            BASELINE.json configs[2]: "human-genome-shaped graph (3.1 Gb, ~25M SNV/indel/SV), 30x PE150 across
                   8xB200" -- the same generator run in 128 Mb windows, variants every ~124 bp, ~1.4e9 index k-mers
                   that never leave the device (vg_index_create_device)
index = canonical k-mer hashes of every reference- and alt-allele window overlapping a variant; reads drawn from two
haplotypes with 0.3 % substitution errors (1/5 of them N).  A *step* is one whole sample: zero the counters, count
every read k-mer of the sample against the index, produce the count vector.

  value     k-mer positions / s, reads already resident in HBM (CUDA events, max over ranks)
  e2e       the same through the reference-facing call: vg_count_files over plain FASTQ files on tmpfs (what the
            reference arm reads) -> counts in host memory; file reads, H2D and D2H inside the timed region
  e2e_staged  vg_count_begin / vg_count_submit from pinned host memory holding parsed "read\\n" records /
            vg_count_end_slots into host memory (round 1's e2e)
  e2e_gz    vg_count_files over gzip-compressed FASTQ (a bounded sample, N=1): the parallel inflater against zlib on one
            thread per file
  roofline  the count pass against the measured HBM copy bandwidth (MEASURED_PEAKS.json) with
            B_alg = 1 + 32 + 32 h bytes per position; per-kernel times from the library's own CUDA
            events, the L1TEX gather rate that actually binds the kernels, and the measured random-sector rate
  cpu_baseline  the reference's own FastqKmer::build_fastq_index (oracle/_ref) on a bounded sample
  --impl reference   times that reference CPU path alone, same config/metric/unit

N > 1 (torchrun): reads are sharded over the ranks and the index is replicated; the only exchange is one reduce of the
count vectors per sample: min(255, sum over ranks), exact, each rank reading its peers' vectors over
NVLink (vg_count_allreduce: CUDA IPC peer memory, a device-side barrier, no NCCL in the data path; --reduce nccl is
the alternative).  --scaling weak (default, the driver's line): every rank counts its own --coverage sample;
--scaling strong: ONE --coverage sample is cut over the ranks.  --index sharded cuts the index itself over the ranks
(the index > HBM layout): the scatter kernel stores every k-mer straight into the owning GPU's key list over NVLink.
At N == 8 the default run adds a "human_scale" object: the human workload, strong scaling, measured in the same job.
"""
from __future__ import annotations

import argparse
import gc
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


def make_variants(dev, L: int, nvar: int, g: torch.Generator, ref: torch.Tensor):
    pos = torch.randint(1000, L - 1000, (int(nvar * 1.03),), generator=g, device=dev)
    pos = torch.unique(pos)[:nvar]
    keep = torch.ones_like(pos, dtype=torch.bool)  # variants at least 10 bp apart so spans never overlap
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
    return alt, pos, vlen


def make_graph(dev, genome_len: int, nvar: int, seed: int, want_keys: bool = True):
    """-> (ref codes u8[L], alt codes u8[L], var_pos int64[nvar], var_len int64[nvar], keys int64 unique)"""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    L = genome_len
    ref = torch.randint(0, 4, (L,), generator=g, device=dev, dtype=torch.uint8)
    alt, pos, vlen = make_variants(dev, L, nvar, g, ref)
    if not want_keys:
        return ref, alt, pos, vlen, None
    nv = pos.numel()
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


def make_graph_windows(dev, genome_len: int, nvar: int, seed: int, window: int = 128_000_000, want_keys: bool = True):
    """The same graph shape for genomes of any length, built in windows so that no temporary exceeds a few GB.
    The keys stay where they are made (device).  Not de-duplicated: a random genome repeats a 27-mer a few dozen times
    in 1.4e9 windows; the device index reports and tolerates duplicates (they share a slot)."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    L = genome_len
    ref = torch.empty(L, dtype=torch.uint8, device=dev)
    for s in range(0, L, 1 << 30):
        e = min(L, s + (1 << 30))
        ref[s:e] = torch.randint(0, 4, (e - s,), generator=g, device=dev, dtype=torch.uint8)
    alt, pos, vlen = make_variants(dev, L, nvar, g, ref)
    if not want_keys:
        return ref, alt, pos, vlen, None
    lo_all = (pos - K + 1).clamp(min=0)          # first window start that overlaps each variant
    hi_all = pos + vlen                          # one past the last
    est = int((hi_all - lo_all).sum().item()) * 2 + 1024
    keys = torch.empty(est, dtype=torch.int64, device=dev)
    nkeys = 0
    n = L - K + 1
    for s in range(0, n, window):
        e = min(n, s + window)
        i0 = int(torch.searchsorted(hi_all, torch.tensor([s], device=dev), right=True).item())
        i1 = int(torch.searchsorted(lo_all, torch.tensor([e], device=dev), right=False).item())
        if i1 <= i0:
            continue
        d = torch.zeros(e - s + 1, dtype=torch.int32, device=dev)
        one = torch.ones(i1 - i0, dtype=torch.int32, device=dev)
        d.index_add_(0, (lo_all[i0:i1] - s).clamp(min=0), one)
        d.index_add_(0, (hi_all[i0:i1] - s).clamp(max=e - s), -one)
        cover = torch.cumsum(d, 0, dtype=torch.int32)[: e - s] > 0
        del d, one
        for codes in (ref, alt):
            kk = t_window_keys(codes[s: e + K - 1], K)[cover]
            keys[nkeys: nkeys + kk.numel()] = kk
            nkeys += kk.numel()
            del kk
        del cover
    return ref, alt, pos, vlen, keys[:nkeys]


def make_reads(dev, ref, alt, pos, vlen, coverage: float, seed: int) -> torch.Tensor:
    """PE150 from two haplotypes -> uint8 [nreads * 151] staged chunk ('read\\n' records) on `dev`."""
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    L = ref.numel()
    flat = torch.empty(2 * L, dtype=torch.uint8, device=dev)
    for hidx in range(2):
        carry = torch.rand(pos.numel(), generator=g, device=dev) < 0.3
        d = torch.zeros(L + 1, dtype=torch.int32, device=dev)
        one = torch.ones(int(carry.sum()), dtype=torch.int32, device=dev)
        d.index_add_(0, pos[carry], one)
        d.index_add_(0, pos[carry] + vlen[carry], -one)
        m = torch.cumsum(d, 0, dtype=torch.int32)[:L] > 0
        del d
        flat[hidx * L: (hidx + 1) * L] = torch.where(m, alt, ref)
        del m
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
        m1 = flat[(base + start)[:, None] + ar[None, :]]
        m2 = 3 - flat[(base + start + ins - 1)[:, None] - ar[None, :]]
        both = torch.cat([m1, m2.to(torch.uint8)])
        err = torch.rand(both.shape, generator=g, device=dev) < 0.003
        sub = torch.randint(0, 5, both.shape, generator=g, device=dev, dtype=torch.uint8)  # 1/5 -> N
        both = torch.where(err, sub, both)
        out[2 * b0: 2 * b0 + 2 * nb, :READ_LEN] = ascii_lut[both.to(torch.int64)]
        del m1, m2, both, err, sub
    del flat
    return out.reshape(-1)


def write_fastq_sample(path: str, lines: np.ndarray, nreads: int, first: int = 0) -> int:
    """Reads [first, first + nreads) of a staged chunk -> plain four-line FASTQ; returns bases written."""
    rec = lines[first * (READ_LEN + 1): (first + nreads) * (READ_LEN + 1)].reshape(nreads, READ_LEN + 1)
    head = np.frombuffer(b"@r\n", dtype=np.uint8)
    plus = np.frombuffer(b"+\n", dtype=np.uint8)
    with open(path, "wb") as f:
        B = 1 << 20
        for s in range(0, nreads, B):
            nb = min(B, nreads - s)
            out = np.empty((nb, 3 + READ_LEN + 1 + 2 + READ_LEN + 1), dtype=np.uint8)
            out[:, :3] = head
            out[:, 3:3 + READ_LEN + 1] = rec[s: s + nb]
            out[:, 3 + READ_LEN + 1: 3 + READ_LEN + 3] = plus
            out[:, 3 + READ_LEN + 3: -1] = ord("I")
            out[:, -1] = 10
            out.tofile(f)
    return nreads * READ_LEN


# ------------------------------------------------------------------------------------------------
# helpers
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md's clocks line), sampled through NVML
    in a thread of this process (nvidia-smi takes longer to start than a short timed region lasts); falls back to one
    nvidia-smi query if NVML is not importable."""

    def __init__(self, gpu_index: int):
        self.rows = []
        self.gpu_index = gpu_index
        self.stop_flag = False
        self.nv = None
        try:
            import pynvml as nv
            nv.nvmlInit()
            self.nv = nv
            self.h = nv.nvmlDeviceGetHandleByIndex(gpu_index)
            self.smax = float(nv.nvmlDeviceGetMaxClockInfo(self.h, nv.NVML_CLOCK_SM))
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.nv = None

    def _pump(self):
        nv = self.nv
        R = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
             "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
             "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
             "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
        while not self.stop_flag:
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
                except Exception:
                    mask = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                self.rows.append((time.time(), sm, [k for k, bit in R.items() if mask & bit]))
            except Exception:
                pass
            time.sleep(0.01)

    def mark(self):
        return time.time()

    def stop(self, t0: float, t1: float) -> dict:
        if self.nv is None:
            try:
                out = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm", "--format=csv,noheader,nounits",
                                      "-i", str(self.gpu_index)], capture_output=True, text=True, timeout=20).stdout
                f = [float(x) for x in out.strip().split(",")]
                return {"sm_mhz": f[0], "sm_max_mhz": f[1], "reasons": [], "samples": 1, "how": "one nvidia-smi query after the timed region"}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"], "samples": 0}
        self.stop_flag = True
        self.th.join(timeout=1.0)
        rows = [r for r in self.rows if t0 <= r[0] <= t1] or self.rows[-3:]
        reasons = sorted({x for r in rows for x in r[2]})
        return {"sm_mhz": float(np.median([r[1] for r in rows])) if rows else None, "sm_max_mhz": self.smax,
                "reasons": reasons, "samples": len(rows), "how": "NVML, every 10 ms during the timed region"}


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


SPECS = {
    "chr20": dict(genome_mb=64, variants=1_500_000, coverage=30.0, label="chr20-shaped"),
    "human": dict(genome_mb=3100, variants=25_000_000, coverage=30.0, label="human-genome-shaped"),
}


def workload_name(w) -> str:
    return (f"{w['label']} synthetic graph ({w['genome_mb']} Mb, ~{w['variants'] / 1e6:.2f}M variants), "
            f"{w['coverage']:g}x PE150, k={K}")


def tmpfs_dir():
    return tempfile.mkdtemp(prefix="vgbench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)


def rm_tree(d):
    try:
        for f in os.listdir(d):
            os.remove(os.path.join(d, f))
        os.rmdir(d)
    except OSError:
        pass


# ------------------------------------------------------------------------------------------------
# the reference arm: FastqKmer::build_fastq_index on the host cores
# ------------------------------------------------------------------------------------------------
def reference_run(cpu_sample_reads, keys_np: np.ndarray, lines_np: np.ndarray, steps: int, warmup: int, want_counts=False):
    from tests import oracle_binding as ob
    cores = os.cpu_count() or 1
    nreads = min(cpu_sample_reads, lines_np.size // (READ_LEN + 1))
    tmpdir = tmpfs_dir()
    fq = os.path.join(tmpdir, "sample.fq")
    write_fastq_sample(fq, lines_np, nreads)
    sample = f"first {nreads} reads of the sample as plain FASTQ on tmpfs, index of {keys_np.size} k-mers"
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
        pos = emitted_positions(lines_np[: nreads * (READ_LEN + 1)])
        sec = float(np.mean(times))
        out = {"value": pos / sec, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample,
               "seconds_per_step": sec, "positions": pos, "counts": counts, "nreads": nreads}
    finally:
        rm_tree(tmpdir)
    return out


def emitted_positions(sub) -> int:
    """Emitted k-mer positions of a chunk (N-free reads give reads x 124; errors put N in some).  Plain numpy
    (the odd-k run-length rule of SURVEY appendix A.4)."""
    b = np.ascontiguousarray(sub)
    valid = np.isin(b, np.frombuffer(b"ACGTacgtUu", dtype=np.uint8))
    idx = np.arange(b.size, dtype=np.int64)
    last_bad = np.maximum.accumulate(np.where(~valid, idx, -1))
    return int(((idx - last_bad) >= K).sum())


# ------------------------------------------------------------------------------------------------
# one workload on our arm -> the JSON line's fields
# ------------------------------------------------------------------------------------------------
def measure(a, w, steps, warmup, rank, world, local, dist, capi, vdist, want_files_e2e=True, want_cpu=True, want_gz_e2e=True,
            staged_cap_bytes=None):
    dev = torch.device("cuda", local)
    L = w["genome_mb"] * 1_000_000
    big = L > 512_000_000
    strong = a.scaling == "strong" or w.get("strong", False)
    cov_rank = w["coverage"] / world if strong else w["coverage"]
    t_gen0 = time.perf_counter()
    sharded = world > 1 and a.index == "sharded"
    # replicated index on several GPUs: rank 0 alone makes the keys and builds the index, the others receive a replica
    builder = rank == 0 or sharded or a.reduce == "nccl"
    if big:
        ref, alt, pos, vlen, keys = make_graph_windows(dev, L, w["variants"], seed=20261017, want_keys=builder)
    else:
        ref, alt, pos, vlen, keys = make_graph(dev, L, w["variants"], seed=20261017, want_keys=builder)
    nk = torch.tensor([int(keys.numel()) if keys is not None else 0], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(nk, op=dist.ReduceOp.MAX)
    nkeys = int(nk.item())
    torch.cuda.synchronize()

    ctx = capi.Context(local, buffer_mb=a.buffer_mb)
    stream = torch.cuda.Stream(device=dev)  # explicit: the legacy default stream's handle is 0 (= "unset")
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)
    # ---- N > 1: map the peers' memory (handles travel through torch.distributed, data never does) ----
    comm, reduce_how = None, ("nccl" if world > 1 else None)
    if world > 1 and (sharded or a.reduce == "p2p"):
        def exchange(mine: bytes):
            got = [None] * world
            dist.all_gather_object(got, mine)
            return got
        round_bytes = a.round_mb << 20
        table_est = int(nkeys / world / (4 * (a.load_factor or 0.3)) * 32 * 1.1) + (64 << 20)
        arena = (2 * nkeys + (2 << 20)) + ((table_est + table_est // 8 + nkeys + round_bytes * 10 + (128 << 20)) if sharded else 0)
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
    replica = comm is not None and not sharded
    if world > 1 and not sharded and not replica and keys is None:
        raise SystemExit("bench.py: peer mapping failed and this rank has no keys; rerun with --reduce nccl")
    # a sample of the index for the in-run parity check against the reference (every key when the index is small)
    keys_np = sample_keys_np = sample_idx = None
    if keys is not None:
        if big:
            sel = (((keys >> 8) * -7046029254386353131) >> 58) == 0  # 1/64 of the keys, by hash (0x9E3779B97F4A7C15 as int64)
            sample_keys_np = keys[sel].cpu().numpy().view(np.uint64)
            sample_idx = torch.nonzero(sel).reshape(-1)
            del sel
        else:
            keys_np = keys.cpu().numpy().view(np.uint64)
            sample_keys_np = keys_np
    t_build0 = time.perf_counter()
    if sharded and big:  # every rank made the keys on its own GPU: they are read where they lie
        ix = capi.Index(ctx, (keys.data_ptr(), nkeys), K, a.load_factor, comm=comm, round_bytes=(a.round_mb << 20))
    elif sharded:
        ix = capi.Index(ctx, keys_np, K, a.load_factor, comm=comm, round_bytes=(a.round_mb << 20))
    elif keys is None:
        ix = None
    elif big:
        ix = capi.Index(ctx, (keys.data_ptr(), nkeys), K, a.load_factor)
    else:
        ix = capi.Index(ctx, keys_np, K, a.load_factor)
    torch.cuda.synchronize()
    t_build = time.perf_counter() - t_build0
    t_repl0 = time.perf_counter()
    if replica:
        ix = comm.replicate(0, ix)  # table, slot order and pre-filter travel over NVLink; one slot order for all ranks
        torch.cuda.synchronize()
    t_repl = time.perf_counter() - t_repl0
    del keys
    torch.cuda.empty_cache()
    lines_dev = make_reads(dev, ref, alt, pos, vlen, cov_rank, seed=1000 + rank)
    del ref, alt, pos, vlen
    torch.cuda.empty_cache()
    torch.cuda.synchronize()
    t_gen = time.perf_counter() - t_gen0
    nbytes = lines_dev.numel()
    rec = READ_LEN + 1
    # host copy of the reads (pinned): all of them, or (staged_cap_bytes) the first few GB for the rate
    host_bytes = nbytes if staged_cap_bytes is None else min(nbytes, staged_cap_bytes // rec * rec)
    lines_host = torch.empty(host_bytes, dtype=torch.uint8, pin_memory=True)
    lines_host.copy_(lines_dev[:host_bytes])
    torch.cuda.synchronize()

    nslots = ix.slots if not sharded else 0
    out32 = torch.empty(max(nkeys, 1), dtype=torch.int32, device=dev) if (world > 1 and not replica and not sharded) else None
    out8 = None
    counts_host = torch.empty(max(nkeys, nslots, 1), dtype=torch.uint8, pin_memory=True)
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
        if replica and a.replicas_only:  # independent samples: the rank's own counts are the result
            return ix.slots_device()
        if replica:  # reduce-scatter + all-gather over peer memory, in place, in slot order
            return comm.allreduce_slots(ix, want_host=False)[1]
        if world > 1:  # --reduce nccl: all-reduce of u32, clamp to 255
            ix.extract_device(out32.data_ptr(), 4)
            return vdist.reduce_counts(out32)
        return ix.slots_device()  # the sample's result: the u8 count vector, on the device, in slot order

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(warmup):
        device_step()
    barrier()
    positions, hits = ix.stats()
    keys_pass = ix.keys_scattered or positions  # k-mers of one sample that passed the pre-filter
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
    launches_timed = ix.launches - launches0
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

    # ---- the same pass once more with the library's per-phase CUDA events (diagnostic, untimed) ----------
    phases = None
    if not sharded:
        ix.set_timing(True)
        for _ in range(2):
            device_step()
        torch.cuda.synchronize()
        phases = ix.timing()
        ix.set_timing(False)

    # ---- counts of the device path, key order, for the checks below ---------------------------------------
    if sharded:
        device_counts = device_step(want=True)
    elif replica:
        ix.begin()
        ix.submit_device(lines_dev.data_ptr(), nbytes)
        if big:
            device_counts = None
            comm.allreduce_slots(ix, want_host=False)
        else:
            device_counts = comm.allreduce_slots(ix)[0][ix.slot_perm()]
        ix.end(want_counts=False)
    elif world > 1:
        device_counts = device_step().cpu().numpy()
    else:
        ix.begin()
        ix.submit_device(lines_dev.data_ptr(), nbytes)
        if big:
            device_counts = None  # checked on the sample below
            ix.end(want_counts=False)
        else:
            device_counts = ix.end()[0]

    # ---- e2e_staged: parsed reads in pinned host memory through the C ABI, copies inside the timed region ----
    ctx.set_stream(0)  # the staged path pipelines copy and compute on the context's own streams
    host_positions = positions if host_bytes == nbytes else None

    def staged_step():
        ix.begin()
        if sharded:
            for o, ln in round_cuts:
                ix.submit_ptr(lines_host.data_ptr() + o, ln)
                ix.flush()
            capi._chk(capi.lib.vg_count_end(ix._h, counts_host.data_ptr(), None, None))
            return
        ix.submit_ptr(lines_host.data_ptr(), host_bytes)
        if replica:
            capi._chk(capi.lib.vg_count_allreduce_slots(comm._h, ix._h, counts_host.data_ptr(), None))
            capi._chk(capi.lib.vg_count_end(ix._h, None, None, None))
        elif world > 1:
            ix.extract_device(out32.data_ptr(), 4, stream.cuda_stream)  # syncs the context streams first
            counts_host.copy_(vdist.reduce_counts(out32), non_blocking=True)
            torch.cuda.synchronize()
            capi.lib.vg_count_end(ix._h, None, None, None)
        else:
            capi._chk(capi.lib.vg_count_end_slots(ix._h, counts_host.data_ptr(), None, None))

    e2e_staged = None
    if host_bytes == nbytes or not sharded:
        for _ in range(2):
            staged_step()
        if host_positions is None:
            host_positions = ix.stats()[0]
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            staged_step()
        barrier()
        st_s = (time.perf_counter() - t0) / steps
        te = torch.tensor([st_s], dtype=torch.float64, device=dev)
        hp = torch.tensor([float(host_positions)], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(hp, op=dist.ReduceOp.SUM)
        staged_ok = None
        if device_counts is not None and host_bytes == nbytes:
            got = counts_host.numpy()
            if world == 1 or replica:
                got = got[:nslots][ix.slot_perm()]
            staged_ok = bool(np.array_equal(got[:nkeys], device_counts[:nkeys]))
        e2e_staged = {"value": float(hp.item()) / float(te.item()), "unit": UNIT, "h2d_bytes_per_step": int(host_bytes),
                      "d2h_bytes_per_step": int(nslots if (world == 1 or replica) else nkeys), "counts_equal_device_path": staged_ok,
                      "input": "parsed 'read\\n' records in pinned host memory"
                               + ("" if host_bytes == nbytes else f" (first {host_bytes} bytes of each rank's reads)")}

    # ---- e2e: the reference-facing call, from the FASTQ files the reference arm reads ----------------------
    e2e = None
    e2e_gz = None
    if want_files_e2e and not sharded:
        tmpdir = tmpfs_dir()
        try:
            lines_np = lines_host.numpy()
            nreads = host_bytes // rec
            half = nreads // 2
            fqs = [os.path.join(tmpdir, f"r{rank}_1.fq"), os.path.join(tmpdir, f"r{rank}_2.fq")]
            write_fastq_sample(fqs[0], lines_np, half, 0)
            write_fastq_sample(fqs[1], lines_np, nreads - half, half)
            file_bytes = sum(os.path.getsize(f) for f in fqs)
            threads = max(1, (os.cpu_count() or 1) // world)

            def files_step():
                ix.begin()
                rb = ix.count_files(fqs, threads=threads)
                if replica:
                    capi._chk(capi.lib.vg_count_allreduce_slots(comm._h, ix._h, counts_host.data_ptr(), None))
                    capi._chk(capi.lib.vg_count_end(ix._h, None, None, None))
                elif world > 1:
                    ix.extract_device(out32.data_ptr(), 4, stream.cuda_stream)
                    counts_host.copy_(vdist.reduce_counts(out32), non_blocking=True)
                    torch.cuda.synchronize()
                    capi.lib.vg_count_end(ix._h, None, None, None)
                else:
                    capi._chk(capi.lib.vg_count_end_slots(ix._h, counts_host.data_ptr(), None, None))
                return rb

            for _ in range(2):
                rb = files_step()
            barrier()
            t0 = time.perf_counter()
            for _ in range(steps):
                files_step()
            barrier()
            f_s = (time.perf_counter() - t0) / steps
            te = torch.tensor([f_s], dtype=torch.float64, device=dev)
            hp = torch.tensor([float(host_positions)], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(te, op=dist.ReduceOp.MAX)
                dist.all_reduce(hp, op=dist.ReduceOp.SUM)
            files_ok = None
            if device_counts is not None and host_bytes == nbytes:
                got = counts_host.numpy()
                if world == 1 or replica:
                    got = got[:nslots][ix.slot_perm()]
                files_ok = bool(np.array_equal(got[:nkeys], device_counts[:nkeys])) and rb == nreads * READ_LEN
            e2e = {"value": float(hp.item()) / float(te.item()), "unit": UNIT,
                   "h2d_bytes_per_step": int(ix.h2d_bytes_last), "d2h_bytes_per_step": int(nslots if (world == 1 or replica) else nkeys),
                   "counts_equal_device_path": files_ok, "file_bytes_per_step": int(file_bytes), "host_threads": threads,
                   "input": "two plain four-line FASTQ files per rank on tmpfs through vg_count_files "
                            "(the same kind of file the reference arm reads)"}
            # ---- e2e_gz: the same call on gzip-compressed FASTQ (what sequencers deliver) ----------------------------------
            # one ordinary single-member .gz per mate (zlib level 6), a bounded sample; the parallel inflater
            # (vg_gzip.cpp) against zlib on one thread per file (VG_GZ_PARALLEL=0: round 1's road, and the reference's)
            if world == 1 and want_gz_e2e:
                try:
                    import threading
                    import zlib
                    gz_reads = min(nreads, 2_000_000)
                    ghalf = gz_reads // 2
                    plain = [os.path.join(tmpdir, "g_1.fq"), os.path.join(tmpdir, "g_2.fq")]
                    write_fastq_sample(plain[0], lines_np, ghalf, 0)
                    write_fastq_sample(plain[1], lines_np, gz_reads - ghalf, ghalf)
                    gzs = [f + ".gz" for f in plain]

                    def deflate(src, dst):
                        co = zlib.compressobj(6, zlib.DEFLATED, 31)
                        with open(src, "rb") as fi, open(dst, "wb") as fo:
                            for blk in iter(lambda: fi.read(8 << 20), b""):
                                fo.write(co.compress(blk))
                            fo.write(co.flush())

                    th = [threading.Thread(target=deflate, args=(a_, b_)) for a_, b_ in zip(plain, gzs)]
                    [t.start() for t in th]
                    [t.join() for t in th]
                    gz_bytes = sum(os.path.getsize(f) for f in gzs)
                    text_bytes = sum(os.path.getsize(f) for f in plain)
                    [os.unlink(f) for f in plain]
                    gz_positions = gz_reads * (READ_LEN - a.kmer + 1)  # upper bound; the exact figure comes from the pass below

                    def gz_step():
                        ix.begin()
                        rb_ = ix.count_files(gzs, threads=threads)
                        c_, pos_, _ = ix.end_slots()
                        return rb_, c_, pos_

                    os.environ["VG_GZ_PARALLEL"] = "0"
                    t0 = time.perf_counter()
                    rb_z, c_z, pos_z = gz_step()
                    zlib_s = time.perf_counter() - t0
                    os.environ["VG_GZ_PARALLEL"] = "1"
                    gz_step()
                    t0 = time.perf_counter()
                    gsteps = 3
                    for _ in range(gsteps):
                        rb_p, c_p, pos_p = gz_step()
                    gz_s = (time.perf_counter() - t0) / gsteps
                    e2e_gz = {"value": pos_p / gz_s, "unit": UNIT, "zlib_road_value": pos_z / zlib_s,
                              "inflated_text_gb_per_s": text_bytes / gz_s / 1e9, "gz_bytes_per_step": int(gz_bytes),
                              "text_bytes_per_step": int(text_bytes), "host_threads": threads,
                              "counts_equal_zlib_road": bool(rb_p == rb_z and pos_p == pos_z and np.array_equal(c_p, c_z)),
                              "input": f"first {gz_reads} reads as two single-member .gz files (zlib level 6) on tmpfs through vg_count_files"}
                except Exception as ex:  # a side measurement: never at the cost of the line itself
                    os.environ.pop("VG_GZ_PARALLEL", None)
                    capi.lib.vg_count_end(ix._h, None, None, None)  # in case it stopped between begin and end
                    e2e_gz = {"error": f"{type(ex).__name__}: {ex}"}
        finally:
            rm_tree(tmpdir)
    ctx.set_stream(stream.cuda_stream)

    # ---- roofline of the count pass ------------------------------------------------------------------------
    h = hits / max(positions, 1)
    b_alg = 1.0 + 32.0 + 32.0 * h
    achieved = positions * b_alg / (kernel_ms * 1e-3) / 1e9
    peak, peak_src = measured_peak_gbs()
    rnd_gbs = rnd_sec = h2d_gbs = None
    if rank == 0:
        try:
            free, _ = torch.cuda.mem_get_info()
            tb = min(8 << 30, int(free * 0.5))
            rnd_gbs, rnd_sec = ctx.probe_random_sectors(tb, 64)
        except Exception as ex:  # diagnostic only
            sys.stderr.write(f"random-sector probe failed: {ex}\n")
        try:  # the H2D ceiling of this box for one GPU: a plain pinned copy
            nb = min(host_bytes, 1 << 30)
            tgt = torch.empty(nb, dtype=torch.uint8, device=dev)
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tgt.copy_(lines_host[:nb], non_blocking=True)
            c0.record(stream)
            for _ in range(3):
                tgt.copy_(lines_host[:nb], non_blocking=True)
            c1.record(stream)
            torch.cuda.synchronize()
            h2d_gbs = 3 * nb / (c0.elapsed_time(c1) * 1e-3) / 1e9
            del tgt
        except Exception as ex:
            sys.stderr.write(f"H2D probe failed: {ex}\n")
    if e2e is not None:
        e2e["h2d_ceiling_gbs_one_gpu"] = h2d_gbs
    if e2e_staged is not None:
        e2e_staged["h2d_ceiling_gbs_one_gpu"] = h2d_gbs
    probes_per_s = positions / (kernel_ms * 1e-3)
    # What binds the two kernels is not bytes but random-access ISSUE: one L1TEX wavefront per clock and SM (289 G/s measured
    # with tools/gather_bench.cu, whatever the access size).  Wavefronts of a pass, from the pass's own counters and the
    # per-access costs measured there: one pre-filter gather per `span` positions, per key that reaches the lists a
    # shared-memory rank atomic (0.68) in the scatter -- and in the re-scatter of a two-level index -- and a bucket load in
    # the sweep (1.2 with the second halves), per hit a reduction into the side counters (1.3).
    span = 8 if nkeys > (128 << 20) else 4
    wavefronts = positions / span + keys_pass * (0.68 + 1.2 + (0.68 if ix.slices > ix.partitions else 0.0)) + hits * 1.3
    props = torch.cuda.get_device_properties(dev)
    sm_clock_hz = (clk.get("sm_mhz") or 1965.0) * 1e6
    l1tex_peak = props.multi_processor_count * sm_clock_hz
    binding = {"unit": "L1TEX wavefront issue (1 per clock and SM): every pre-filter gather, bucket load, side-counter reduction and "
                       "shared-memory rank atomic costs about one",
               "wavefronts_per_pass_model": wavefronts, "peak_wavefronts_per_s": l1tex_peak,
               "frac": wavefronts / l1tex_peak / (kernel_ms * 1e-3), "keys_after_prefilter": int(keys_pass),
               "measured": "profiles/r2_*_ncu_*.txt: l1tex__throughput of scatter_kernel and probe_slice_kernel"}
    # "bound": the contract's choice is hbm | tensor; the bytes side of this path is HBM (no contraction anywhere), but see
    # "binding": the pass is issue-bound on random accesses that the partitioned layout turns into L2 hits
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic(ix.partitions > 0), "peak_source": peak_src,
                "kernel": ("count pass = vg::scatter_kernel + vg::probe_slice_kernel sweep (K1 | K2+K3)" if ix.partitions
                           else "vg::count_kernel (K1+K2+K3 fused)"),
                "kernel_ms": kernel_ms, "bytes_per_position": b_alg, "hit_fraction": h,
                "random_sector_peak_gbs": rnd_gbs,
                "frac_of_random_sector_peak": (probes_per_s * (1 + h) / rnd_sec) if rnd_sec else None,
                "phases_ms": phases, "binding": binding}

    # ---- CPU baseline beside it + the same sample through the CUDA path (rank 0) -----------------------------
    cpu = parity = None
    if rank == 0 and want_cpu and not a.no_cpu_baseline:
        lines_np = lines_host.numpy()
        r = reference_run(a.cpu_sample_reads, sample_keys_np, lines_np, steps=1, warmup=0, want_counts=True)
        cpu = {"value": r["value"], "unit": UNIT, "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}
    if want_cpu and not a.no_cpu_baseline:
        # the same reads through the CUDA path must give the reference's counts exactly (rank 0 checks; with N > 1
        # the other ranks submit nothing and the reduce -- or, for a sharded index, the k-mer exchange -- is part of what
        # is checked)
        nr = min(a.cpu_sample_reads, host_bytes // rec)
        barrier()  # rank 0 comes from the CPU run: the library's own barriers give a peer 20 s
        ix.begin()
        if rank == 0:
            ix.submit_ptr(lines_host.data_ptr(), nr * rec)
        if sharded:
            ix.flush()  # collective: one round
            mine, mp, _ = ix.end()  # collective: the counts of all keys, in key order
        elif replica:
            mine = comm.allreduce_counts(ix, want_host=True)
            mp = ix.end(want_counts=False)[1]
        elif world > 1:
            ix.extract_device(out32.data_ptr(), 4)
            mine = vdist.reduce_counts(out32).cpu().numpy()
            mp = ix.end(want_counts=False)[1]
        else:
            mine, mp, _ = ix.end()
        if rank == 0 and cpu is not None:
            if sample_idx is not None:
                mine = mine[sample_idx.cpu().numpy()]
            parity = {"sample_counts_bit_exact": bool(np.array_equal(mine, r["counts"])),
                      "sample_positions_equal": bool(mp == r["positions"]),
                      "index_entries_checked": int(sample_keys_np.size)}

    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if strong else "weak",
                "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "config": {"workload": workload_name(w), "index_kmers": nkeys, "index_table_bytes": ix.table_bytes,
                           "table_partitions": ix.partitions, "table_slices": ix.slices, "index_duplicate_kmers": ix.duplicates,
                           "reads_per_gpu": nbytes // rec, "positions_per_gpu": positions,
                           "coverage_per_gpu": cov_rank,
                           "parallelism": (f"reads sharded x{world}, index "
                                           + (f"sharded x{world} (k-mer all-to-all fused into the scatter over NVLink, "
                                              f"{len(round_cuts)} rounds)" if sharded else
                                              ("built on rank 0 and replicated over NVLink, " +
                                               ("every rank counts samples of its own: no exchange" if a.replicas_only else
                                                "counts combined in slot order by a reduce-scatter + all-gather over peer memory")
                                               if replica else f"replicated, counts reduced by {reduce_how}"))
                                           if world > 1 else "1 GPU"),
                           "l2": "inputs (reads + index table) far larger than the 126 MB L2; no explicit flush",
                           "generate_s": t_gen, "index_build_s": t_build, "index_replicate_s": t_repl if replica else None},
                "e2e": e2e if e2e is not None else e2e_staged, "e2e_staged": e2e_staged, "e2e_gz": e2e_gz,
                "gpu_launches": int(launches_timed), "clocks": clk, "roofline": roofline, "cpu_baseline": cpu,
                "parity": parity}
    if world > 1:
        barrier()  # nobody unmaps a peer's arena while it is still in use
    ix.close()
    if comm is not None:
        comm.close()
    ctx.close()
    del lines_dev, lines_host, counts_host, out8, out32
    gc.collect()
    torch.cuda.empty_cache()
    return line


# ------------------------------------------------------------------------------------------------
def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="chr20", choices=sorted(SPECS))
    ap.add_argument("--genome-mb", type=int, default=None)
    ap.add_argument("--variants", type=int, default=None)
    ap.add_argument("--coverage", type=float, default=None)
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every rank counts its own --coverage sample; strong: one --coverage sample cut over the ranks")
    ap.add_argument("--replicas-only", action="store_true",
                    help="N > 1: every rank counts samples of its own, nothing is exchanged (BASELINE config 5: samples dealt over the GPUs)")
    ap.add_argument("--cpu-sample-reads", type=int, default=1_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-files-e2e", action="store_true")
    ap.add_argument("--no-gz-e2e", action="store_true")
    ap.add_argument("--human", default="auto", choices=["auto", "on", "off"],
                    help="add the human-scale section to the line (auto: at N == 8)")
    ap.add_argument("--human-coverage", type=float, default=30.0)
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
    w = dict(SPECS[a.config])
    for key, val in (("genome_mb", a.genome_mb), ("variants", a.variants), ("coverage", a.coverage)):
        if val is not None:
            w[key] = val
    L = w["genome_mb"] * 1_000_000

    if a.impl == "reference":
        if rank != 0:
            return
        dev = torch.device("cuda", local) if torch.cuda.is_available() else torch.device("cpu")
        if L > 512_000_000:
            raise SystemExit("bench.py --impl reference: the reference's host map does not hold a human-scale index here; "
                             "use the default workload")
        ref, alt, pos, vlen, keys = make_graph(dev, L, w["variants"], seed=20261017)
        keys_np = keys.cpu().numpy().view(np.uint64)
        nreads = a.cpu_sample_reads
        cov = nreads * READ_LEN / L * 1.02 + 0.01
        lines_np = make_reads(dev, ref, alt, pos, vlen, min(cov, w["coverage"]), seed=1000).cpu().numpy()
        del ref, alt, pos, vlen, keys
        r = reference_run(a.cpu_sample_reads, keys_np, lines_np, steps, a.warmup)
        line = {"metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
                "warmup": a.warmup, "ms_per_step": r["seconds_per_step"] * 1e3, "higher_is_better": True,
                "scaling": a.scaling, "vs_baseline": None, "dtype": "u64", "data": "synthetic",
                "impl": "reference",
                "config": {"workload": workload_name(w), "index_kmers": int(keys_np.size)},
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

    big = L > 512_000_000
    line = measure(a, w, steps, warmup, rank, world, local, dist, capi, vdist,
                   want_files_e2e=not a.no_files_e2e and not big, want_gz_e2e=not a.no_gz_e2e, want_cpu=(world == 1 or big or a.index == "sharded"),
                   staged_cap_bytes=(2 << 30) if big else None)
    if line is not None:
        line["config"]["host_binding"] = numa
    # ---- the north-star configuration beside the driver's line: human-scale graph, one 30x sample over the box ----
    if a.config == "chr20" and (a.human == "on" or (a.human == "auto" and world == 8)):
        hw = dict(SPECS["human"], coverage=a.human_coverage, strong=True)
        human = None
        try:
            hl = measure(a, hw, max(2, min(steps, 3)), 1, rank, world, local, dist, capi, vdist, want_files_e2e=False,
                         want_cpu=True, staged_cap_bytes=2 << 30)
            if hl is not None:
                human = {k: hl[k] for k in ("value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "config",
                                            "e2e_staged", "roofline", "cpu_baseline", "parity", "gpu_launches", "clocks")}
        except Exception as ex:  # the driver's line must survive whatever happens here
            human = {"error": f"{type(ex).__name__}: {ex}"}
        if line is not None:
            line["human_scale"] = human
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
