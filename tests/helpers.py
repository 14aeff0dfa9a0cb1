"""Shared test data: the committed golden fixtures and the seeded reads that go with them."""
from __future__ import annotations

import gzip
import json
import os

import numpy as np

from varigraph_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NOKMER = np.uint64(0xFFFFFFFFFFFFFFFF)


def primitives():
    with open(os.path.join(GOLD, "primitives.json")) as f:
        return json.load(f)


def tiny():
    """-> dict(keys, counts, read_bases, k, m1, m2, lines, genome, variants, graph_bin bytes)"""
    d = os.path.join(GOLD, "tiny")
    with open(os.path.join(d, "params.json")) as f:
        p = json.load(f)
    z = np.load(os.path.join(d, "counts.npz"))
    g = synth.make_genome(p["genome_len"], p["genome_seed"])
    v = synth.make_variants(g, p["nvar"], p["nsamples"], p["ploidy"], p["var_seed"])
    haps = [synth.apply_haplotype(g, v, 0, h) for h in range(p["ploidy"])]
    m1, m2 = synth.make_reads(haps, p["coverage"], len(g), seed=p["read_seed"])
    lines = np.concatenate([synth.reads_to_lines(m1), synth.reads_to_lines(m2)])
    with gzip.open(os.path.join(d, "graph.bin.gz"), "rb") as f:
        graph = f.read()
    with open(os.path.join(d, "S0.varigraph.vcf"), "rb") as f:
        vcf = f.read()
    return dict(keys=z["keys"], counts=z["counts"], read_bases=int(z["read_bases"]), k=int(z["k"]), m1=m1,
                m2=m2, lines=lines, genome=g, variants=v, graph_bin=graph, vcf=vcf, params=p)


def cbf_golden():
    return np.load(os.path.join(GOLD, "cbf.npz"))


def write_tiny_fastqs(tmpdir, t, gz=True):
    ext = ".fq.gz" if gz else ".fq"
    f1, f2 = os.path.join(tmpdir, "S0_1" + ext), os.path.join(tmpdir, "S0_2" + ext)
    synth.write_fastq(f1, t["m1"], "a")
    synth.write_fastq(f2, t["m2"], "b")
    return f1, f2


EDGE_FASTQS = {
    "plain": b"@r1\nACGTACGTACGTACGTACGTACGTACGTACGTA\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n",
    "crlf": b"@r1 desc\r\nACGTACGTACGTACGTACGTACGTACGTACGTA\r\n+\r\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\r\n",
    "multiline": b"@r1\nACGTACGTACGTACGT\nACGTACGTACGTACGTA\n+r1\nIIIIIIIIIIIIIIII\nIIIIIIIIIIIIIIIII\n"
                 b"@r2\nTTTTGGGGCCCCAAAATTTTGGGGCCCCAAAAT\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n",
    "fasta": b">c1 x\nACGTACGTACGTACGTACGTACGTACGTACGTA\nGGGGACGTACGTACGTACGTACGTACGTACGTA\n>c2\nTTTTGGGGCCCCAAAATTTTGGGGCCCCAAAAT\n",
    "qual_at": b"@r1\nACGTACGTACGTACGTACGTACGTACGTACGTA\n+\n@IIIIIIIIIIIIIIIIIIIIIIIIIIIIIII@\n"
               b"@r2\nTTTTGGGGCCCCAAAATTTTGGGGCCCCAAAAT\n+\n+IIIIIIIIIIIIIIIIIIIIIIIIIIIIIII>\n",
    "truncated_qual": b"@r1\nACGTACGTACGTACGTACGTACGTACGTACGTA\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n"
                      b"@r2\nTTTTGGGGCCCCAAAATTTTGGGGCCCCAAAAT\n+\nIIII\n",
    "no_final_newline": b"@r1\nACGTACGTACGTACGTACGTACGTACGTACGTA\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII",
    "lower_n": b"@r1\nacgtacgtacgtacgtNacgtacgtacgtacgtacgtacgtacgtacgtacgu\n+\n"
               b"IIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n",
    "leading_junk": b"junk line\n\n@r1\nACGTACGTACGTACGTACGTACGTACGTACGTA\n+\nIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIIII\n",
}


def free_port() -> int:
    import socket
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def spawn_ranks(fn, make_args, nprocs: int, timeout_s: float = 240.0, attempts: int = 2) -> None:
    """torch.multiprocessing.spawn with a deadline and one retry on a fresh port: a rendezvous that never completes or
    times out (a port grabbed between free_port() and the bind, a straggler from an earlier run) must fail or recover,
    not hang the suite; a real failure fails twice.  make_args(port) -> the argument tuple of fn after the rank."""
    import time

    import torch.multiprocessing as mp
    last = None
    for _ in range(attempts):
        ctx = mp.spawn(fn, args=make_args(free_port()), nprocs=nprocs, join=False)
        deadline = time.time() + timeout_s
        try:
            while not ctx.join(timeout=5.0):
                if time.time() > deadline:
                    raise TimeoutError(f"ranks still running after {timeout_s:.0f} s")
            return
        except Exception as ex:  # a rank failed, or the deadline passed: stop exactly the processes started here
            last = ex
            for pr in ctx.processes:
                if pr.is_alive():
                    pr.terminate()
            for pr in ctx.processes:
                pr.join(10)
    raise last
