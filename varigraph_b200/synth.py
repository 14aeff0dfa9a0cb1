"""Seeded synthetic workloads of the shapes BASELINE.json names.

Genome: iid uniform ACGT.  VCF: sorted, non-overlapping SNV / small indel records with
phased GT columns.  Reads: PE150 drawn from sample 0's two haplotypes, mate 2
reverse-complemented, substitution errors (1/5 of them 'N').  Everything is numpy on
the host: this is workload generation for tests and bench.py, not the hot path.
"""
from __future__ import annotations

import gzip
import io
from dataclasses import dataclass, field

import numpy as np

ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.full(256, ord("N"), dtype=np.uint8)
for _a, _b in zip(b"ACGTacgtNn", b"TGCAtgcaNn"):
    _COMP[_a] = _b


def make_genome(length: int, seed: int = 20261017) -> np.ndarray:
    rng = np.random.default_rng(seed)
    return ACGT[rng.integers(0, 4, size=length, dtype=np.uint8)]


def revcomp(a: np.ndarray) -> np.ndarray:
    return _COMP[a[..., ::-1]]


@dataclass
class Variants:
    pos: np.ndarray          # 0-based start of REF allele (for indels: the anchor base)
    ref: list                # bytes
    alt: list                # bytes
    gt: np.ndarray           # [nvar, nsamples, ploidy] uint8 in {0,1}
    chrom: str = "chr1"
    samples: list = field(default_factory=list)


def make_variants(genome: np.ndarray, nvar: int, nsamples: int = 5, ploidy: int = 2, seed: int = 7,
                  indel_frac: float = 0.14, max_indel: int = 12, alt_prob: float = 0.3,
                  margin: int = 200) -> Variants:
    """Sorted, non-overlapping variants: SNVs plus anchored insertions/deletions."""
    rng = np.random.default_rng(seed)
    L = len(genome)
    span = max_indel + 2
    nslots = (L - 2 * margin) // span
    if nvar > nslots:
        raise ValueError("too many variants for this genome")
    slots = np.sort(rng.choice(nslots, size=nvar, replace=False))
    pos = margin + slots * span + rng.integers(0, 2, size=nvar)
    kind = rng.random(nvar)
    ref, alt = [], []
    for i in range(nvar):
        p = int(pos[i])
        b = genome[p]
        if kind[i] >= indel_frac:
            others = ACGT[ACGT != b]
            ref.append(bytes([b]))
            alt.append(bytes([others[rng.integers(0, 3)]]))
        elif kind[i] < indel_frac / 2:  # insertion
            n = int(rng.integers(1, max_indel))
            ins = ACGT[rng.integers(0, 4, size=n)]
            ref.append(bytes([b]))
            alt.append(bytes([b]) + ins.tobytes())
        else:  # deletion
            n = int(rng.integers(1, max_indel))
            ref.append(genome[p:p + 1 + n].tobytes())
            alt.append(bytes([b]))
    gt = (rng.random((nvar, nsamples, ploidy)) < alt_prob).astype(np.uint8)
    # every variant carried by someone, so the graph keeps it
    none = gt.reshape(nvar, -1).sum(axis=1) == 0
    gt[none, 0, 0] = 1
    return Variants(pos=pos.astype(np.int64), ref=ref, alt=alt, gt=gt,
                    samples=[f"S{i}" for i in range(nsamples)])


def write_fasta(path: str, genome: np.ndarray, chrom: str = "chr1", width: int = 60) -> None:
    with open(path, "wb") as f:
        f.write(b">" + chrom.encode() + b"\n")
        n = len(genome)
        full = (n // width) * width
        if full:
            body = genome[:full].reshape(-1, width)
            out = np.empty((body.shape[0], width + 1), dtype=np.uint8)
            out[:, :width] = body
            out[:, width] = ord("\n")
            f.write(out.tobytes())
        if n > full:
            f.write(genome[full:].tobytes() + b"\n")


def write_vcf(path: str, v: Variants, chrom_len: int) -> None:
    with open(path, "w") as f:
        f.write("##fileformat=VCFv4.2\n")
        f.write(f"##contig=<ID={v.chrom},length={chrom_len}>\n")
        f.write('##FORMAT=<ID=GT,Number=1,Type=String,Description="Genotype">\n')
        f.write("#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t" + "\t".join(v.samples) + "\n")
        for i in range(len(v.pos)):
            gts = "\t".join("|".join(str(int(x)) for x in v.gt[i, s]) for s in range(v.gt.shape[1]))
            f.write(f"{v.chrom}\t{int(v.pos[i]) + 1}\tv{i}\t{v.ref[i].decode()}\t{v.alt[i].decode()}"
                    f"\t.\t.\t.\tGT\t{gts}\n")


def apply_haplotype(genome: np.ndarray, v: Variants, sample: int, hap: int) -> np.ndarray:
    """Sequence of one haplotype of one sample (alt alleles where GT == 1)."""
    carried = np.nonzero(v.gt[:, sample, hap])[0]
    parts, cur = [], 0
    for i in carried:
        p = int(v.pos[i])
        parts.append(genome[cur:p])
        parts.append(np.frombuffer(v.alt[i], dtype=np.uint8))
        cur = p + len(v.ref[i])
    parts.append(genome[cur:])
    return np.concatenate(parts)


def make_reads(haps: list, coverage: float, genome_len: int, read_len: int = 150, seed: int = 11,
               err: float = 0.003, insert=(300, 500)) -> tuple:
    """PE reads -> (mate1 [n, read_len] uint8, mate2 [n, read_len] uint8)."""
    rng = np.random.default_rng(seed)
    npairs = int(round(coverage * genome_len / (2 * read_len)))
    which = rng.integers(0, len(haps), size=npairs)
    ins = rng.integers(insert[0], insert[1] + 1, size=npairs)
    m1 = np.empty((npairs, read_len), dtype=np.uint8)
    m2 = np.empty((npairs, read_len), dtype=np.uint8)
    ar = np.arange(read_len)
    for h, hap in enumerate(haps):
        sel = np.nonzero(which == h)[0]
        if len(sel) == 0:
            continue
        start = (rng.random(len(sel)) * (len(hap) - ins[sel] - 1)).astype(np.int64)
        m1[sel] = hap[start[:, None] + ar[None, :]]
        end = start + ins[sel]
        m2[sel] = _COMP[hap[(end[:, None] - 1 - ar[None, :])]]
    for m in (m1, m2):
        e = rng.random(m.shape) < err
        ne = int(e.sum())
        if ne:
            sub = ACGT[rng.integers(0, 4, size=ne)]
            isn = rng.random(ne) < 0.2
            sub[isn] = ord("N")
            m[e] = sub
    return m1, m2


def reads_to_lines(reads: np.ndarray) -> np.ndarray:
    """[n, L] -> flat uint8 buffer 'read\\nread\\n...' (the staged chunk format)."""
    n, L = reads.shape
    out = np.empty((n, L + 1), dtype=np.uint8)
    out[:, :L] = reads
    out[:, L] = ord("\n")
    return out.reshape(-1)


def write_fastq(path: str, reads: np.ndarray, prefix: str = "r", level: int = 1) -> None:
    n, L = reads.shape
    qual = b"I" * L
    buf = io.BytesIO()
    for i in range(n):
        buf.write(b"@" + prefix.encode() + str(i).encode() + b"\n")
        buf.write(reads[i].tobytes())
        buf.write(b"\n+\n" + qual + b"\n")
    data = buf.getvalue()
    if path.endswith(".gz"):
        with gzip.open(path, "wb", compresslevel=level) as f:
            f.write(data)
    else:
        with open(path, "wb") as f:
            f.write(data)


def random_reads_lines(nreads: int, read_len: int, genome: np.ndarray, seed: int = 3,
                       err: float = 0.003) -> np.ndarray:
    """Fast single-end reads straight off one sequence (both strands) as a staged chunk."""
    rng = np.random.default_rng(seed)
    start = rng.integers(0, len(genome) - read_len, size=nreads)
    ar = np.arange(read_len)
    r = genome[start[:, None] + ar[None, :]]
    flip = rng.random(nreads) < 0.5
    r[flip] = _COMP[r[flip][:, ::-1]]
    e = rng.random(r.shape) < err
    ne = int(e.sum())
    if ne:
        sub = ACGT[rng.integers(0, 4, size=ne)]
        sub[rng.random(ne) < 0.2] = ord("N")
        r[e] = sub
    return reads_to_lines(r)
