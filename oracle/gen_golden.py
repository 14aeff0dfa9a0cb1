#!/usr/bin/env python
"""Generate tests/golden/* by running the UNMODIFIED reference in this container.

TEST INFRASTRUCTURE ONLY.  Needs /root/reference compiled by `make -C oracle ref`
(oracle/_ref/varigraph_ref + oracle/_ref/libvgref.so).  The reference ships no
tests or golden vectors, so every fixture here is the reference's
own output on a seeded synthetic input:

  tests/golden/primitives.json   T0  hash64 / nt4 / Murmur3 / CBF sizing / encoder edge cases
  tests/golden/tiny/             T1-T3 a 40 kb graph: genome, VCF, graph.bin (reference
                                 `construct`; its CBF seeds are random, so the file itself is the
                                 fixture), keys + per-k-mer counts from FastqKmer::build_fastq_index,
                                 and the reference `genotype` VCF
  tests/golden/cbf.npz           construct-side counting Bloom filter built by the reference
                                 with injected seeds

Reads are NOT stored: they are regenerated from the same seeds by varigraph_b200.synth.
"""
import ctypes
import gzip
import json
import os
import random
import shutil
import subprocess
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
from varigraph_b200 import synth  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
REF_BIN = os.path.join(HERE, "_ref", "varigraph_ref")
u64, u32, i64 = ctypes.c_uint64, ctypes.c_uint32, ctypes.c_int64

TINY = dict(genome_len=40_000, nvar=120, nsamples=5, ploidy=2, coverage=30.0,
            genome_seed=101, var_seed=102, read_seed=103)


def load_ref():
    r = ctypes.CDLL(os.path.join(HERE, "_ref", "libvgref.so"))
    r.ref_hash64.restype = u64
    r.ref_hash64.argtypes = [u64, u64]
    r.ref_murmur3_x64_128_sum.restype = u64
    r.ref_murmur3_x64_128_sum.argtypes = [u64, u32]
    r.ref_sketch.restype = i64
    r.ref_sketch.argtypes = [ctypes.c_char_p, i64, u32, ctypes.POINTER(u64), i64]
    r.ref_cbf_create.restype = ctypes.c_void_p
    r.ref_cbf_create.argtypes = [u64, ctypes.c_double, ctypes.POINTER(u64), u32]
    r.ref_cbf_destroy.argtypes = [ctypes.c_void_p]
    r.ref_cbf_size.restype = u64
    r.ref_cbf_size.argtypes = [ctypes.c_void_p]
    r.ref_cbf_num_hashes.restype = u32
    r.ref_cbf_num_hashes.argtypes = [ctypes.c_void_p]
    r.ref_cbf_seeds.argtypes = [ctypes.c_void_p, ctypes.POINTER(u64)]
    r.ref_cbf_add.argtypes = [ctypes.c_void_p, u64]
    r.ref_cbf_find.argtypes = [ctypes.c_void_p, u64]
    r.ref_cbf_count.argtypes = [ctypes.c_void_p, u64]
    r.ref_cbf_filter.restype = ctypes.POINTER(ctypes.c_uint8)
    r.ref_cbf_filter.argtypes = [ctypes.c_void_p]
    r.ref_cbf_fill.argtypes = [ctypes.c_void_p, ctypes.c_char_p, i64, u32]
    r.ref_graph_load.restype = ctypes.c_void_p
    r.ref_graph_load.argtypes = [ctypes.c_char_p, u32]
    r.ref_graph_destroy.argtypes = [ctypes.c_void_p]
    r.ref_graph_num_kmers.restype = u64
    r.ref_graph_num_kmers.argtypes = [ctypes.c_void_p]
    r.ref_graph_kmer_len.restype = u32
    r.ref_graph_kmer_len.argtypes = [ctypes.c_void_p]
    r.ref_graph_keys.argtypes = [ctypes.c_void_p, ctypes.POINTER(u64)]
    r.ref_graph_reset.argtypes = [ctypes.c_void_p]
    r.ref_count_files.restype = ctypes.c_double
    r.ref_count_files.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_char_p), ctypes.c_int, u32,
                                  ctypes.POINTER(ctypes.c_uint8), ctypes.POINTER(u64)]
    return r


EDGE_SEQS = [
    "ACGT", "ACGNTACGTA", "A" * 40, "AT" * 30, "ACGTACGTACGTACGTACGTACGTACGTACGT",
    "acgtnACGTUuNNacgtACGTacgtACGTacgtACGTacgtACGT", "N" * 50, "GATTACA" * 12 + "N" + "TTAGGC" * 9,
    "ACG\rT\x00ACGT" + "C" * 30, "TTTTTTTTTTTTTTTTTTTTTTTTTTTTAAAAAAAAAAAAAAAAAAAAAAAAAAAAA",
    "GCGC" * 16, "A", "AC" * 13 + "A", "ACGTTGCA" * 10 + "N" + "TGCAACGT" * 10,
]


def gen_primitives(r):
    rng = random.Random(20261017)
    out = {"hash64": [], "murmur": [], "nt4": [r.ref_nt4(b) for b in range(256)], "cbf_sizing": [],
           "sketch": []}
    for k in (5, 11, 16, 27, 28):
        m = (1 << (2 * k)) - 1
        xs = [0, 1, m, m - 1] + [rng.getrandbits(2 * k) for _ in range(60)]
        out["hash64"].append({"k": k, "x": [str(x) for x in xs],
                              "y": [str(r.ref_hash64(x, m)) for x in xs]})
    for _ in range(200):
        x, s = rng.getrandbits(64), rng.getrandbits(32)
        out["murmur"].append([str(x), s, str(r.ref_murmur3_x64_128_sum(x, s))])
    for n in (1, 100, 999_974, 63_999_974, 3_099_999_974):
        h = r.ref_cbf_create(n, 0.01, None, 0) if n < 10_000_000 else None
        if h:
            out["cbf_sizing"].append([n, 0.01, str(r.ref_cbf_size(h)), r.ref_cbf_num_hashes(h)])
            r.ref_cbf_destroy(h)
    seqs = list(EDGE_SEQS)
    for t in range(60):
        L = rng.randint(1, 260)
        alpha = ["ACGT", "ACGTNacgtnU", "AT", "ACGTN"][t % 4]
        seqs.append("".join(rng.choice(alpha) for _ in range(L)))
    buf = (u64 * 1024)()
    for s in seqs:
        b = s.encode("latin1")
        for k in (4, 5, 6, 16, 27, 28):
            n = r.ref_sketch(b, len(b), k, buf, 1024)
            out["sketch"].append({"k": k, "seq": s, "keys": [str(buf[i]) for i in range(n)]})
    with open(os.path.join(GOLD, "primitives.json"), "w") as f:
        json.dump(out, f)
    print("primitives:", len(out["sketch"]), "sketch cases")


def tiny_inputs(workdir):
    g = synth.make_genome(TINY["genome_len"], TINY["genome_seed"])
    v = synth.make_variants(g, TINY["nvar"], TINY["nsamples"], TINY["ploidy"], TINY["var_seed"])
    haps = [synth.apply_haplotype(g, v, 0, h) for h in range(TINY["ploidy"])]
    m1, m2 = synth.make_reads(haps, TINY["coverage"], len(g), seed=TINY["read_seed"])
    fa, vcf = os.path.join(workdir, "ref.fa"), os.path.join(workdir, "var.vcf")
    synth.write_fasta(fa, g)
    synth.write_vcf(vcf, v, len(g))
    fq1, fq2 = os.path.join(workdir, "S0_1.fq.gz"), os.path.join(workdir, "S0_2.fq.gz")
    synth.write_fastq(fq1, m1, "a")
    synth.write_fastq(fq2, m2, "b")
    return fa, vcf, fq1, fq2


def gen_tiny(r):
    dst = os.path.join(GOLD, "tiny")
    os.makedirs(dst, exist_ok=True)
    with tempfile.TemporaryDirectory() as wd:
        fa, vcf, fq1, fq2 = tiny_inputs(wd)
        graph = os.path.join(wd, "graph.bin")
        subprocess.run([REF_BIN, "construct", "-r", fa, "-v", vcf, "--save-graph", graph, "-t", "4"],
                       check=True, stderr=subprocess.DEVNULL)
        # T1: per-k-mer counts straight from FastqKmer::build_fastq_index
        h = r.ref_graph_load(graph.encode(), 4)
        n = r.ref_graph_num_kmers(h)
        keys = np.zeros(n, dtype=np.uint64)
        r.ref_graph_keys(h, keys.ctypes.data_as(ctypes.POINTER(u64)))
        counts = np.zeros(n, dtype=np.uint8)
        rb = u64(0)
        files = (ctypes.c_char_p * 2)(fq1.encode(), fq2.encode())
        r.ref_count_files(h, files, 2, 4, counts.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)),
                          ctypes.byref(rb))
        k = r.ref_graph_kmer_len(h)
        r.ref_graph_destroy(h)
        order = np.argsort(keys)
        np.savez_compressed(os.path.join(dst, "counts.npz"), keys=keys[order], counts=counts[order],
                            read_bases=np.uint64(rb.value), k=np.uint32(k))
        # T3: the reference genotype VCF
        cfg = os.path.join(wd, "samples.cfg")
        with open(cfg, "w") as f:
            f.write(f"S0 {fq1} {fq2}\n")
        subprocess.run([REF_BIN, "genotype", "--load-graph", graph, "-s", cfg, "-t", "4"], check=True,
                       cwd=wd, stderr=subprocess.DEVNULL)
        with gzip.open(os.path.join(wd, "S0.varigraph.vcf.gz"), "rb") as f:
            vcf_txt = f.read()
        with open(os.path.join(dst, "S0.varigraph.vcf"), "wb") as f:
            f.write(vcf_txt)
        with open(graph, "rb") as f, gzip.open(os.path.join(dst, "graph.bin.gz"), "wb", 9) as o:
            shutil.copyfileobj(f, o)
        with open(os.path.join(dst, "params.json"), "w") as f:
            json.dump(TINY, f)
        print(f"tiny: {n} k-mers, k={k}, read_bases={rb.value}, nonzero={int((counts > 0).sum())}, "
              f"sum={int(counts.sum())}, vcf={len(vcf_txt)} B")


def gen_cbf(r):
    rng = np.random.default_rng(55)
    g = synth.make_genome(30_000, 77)
    g[5000:5040] = ord("N")
    g[12000:12500] = g[2000:2500]  # a repeat, so some cells exceed 1
    g[20000:20400] = ord("A")      # low complexity: saturating cells
    k = 27
    n = len(g) - k + 1
    seeds = rng.integers(1, 2**63, size=7, dtype=np.uint64)
    h = r.ref_cbf_create(n, 0.01, seeds.ctypes.data_as(ctypes.POINTER(u64)), 7)
    m = r.ref_cbf_size(h)
    r.ref_cbf_fill(h, g.tobytes(), len(g), k)
    filt = np.ctypeslib.as_array(r.ref_cbf_filter(h), shape=(m,)).copy()
    probe = rng.integers(0, 2**54, size=64, dtype=np.uint64) << np.uint64(8) | np.uint64(k)
    cnt = np.array([r.ref_cbf_count(h, int(x)) for x in probe], dtype=np.uint8)
    fnd = np.array([r.ref_cbf_find(h, int(x)) for x in probe], dtype=np.uint8)
    r.ref_cbf_destroy(h)
    np.savez_compressed(os.path.join(GOLD, "cbf.npz"), genome=g, k=np.uint32(k), seeds=seeds,
                        m=np.uint64(m), filter=filt, probe=probe, probe_count=cnt, probe_find=fnd)
    print(f"cbf: m={m}, nonzero={int((filt > 0).sum())}, max={int(filt.max())}")


if __name__ == "__main__":
    os.makedirs(GOLD, exist_ok=True)
    ref = load_ref()
    gen_primitives(ref)
    gen_tiny(ref)
    gen_cbf(ref)
