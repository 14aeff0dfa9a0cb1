#!/usr/bin/env python
"""vg_count_files throughput from FASTQ files on tmpfs (the FastqKmerKernel road): plain FASTQ parsed on
the device, plain through the host kseq reader (VG_RAW_FASTQ=0), gzip (zlib-bound).  One JSON line each."""
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import bench  # noqa: E402


def main():
    genome_mb, variants, cov = 64, 1_500_000, float(os.environ.get("COV", "10"))
    dev = torch.device("cuda", 0)
    ref, alt, pos, vlen, keys = bench.make_graph(dev, genome_mb * 1_000_000, variants, seed=20261017)
    keys_np = keys.cpu().numpy().view(np.uint64)
    lines = bench.make_reads(dev, ref, alt, pos, vlen, cov, seed=1000).cpu().numpy()
    del ref, alt, pos, vlen, keys
    nreads = lines.size // 151
    half = nreads // 2
    tmp = tempfile.mkdtemp(prefix="vgfiles_", dir="/dev/shm")
    f1, f2 = os.path.join(tmp, "S_1.fq"), os.path.join(tmp, "S_2.fq")
    bench.write_fastq_sample(f1, lines, half)
    bench.write_fastq_sample(f2, lines[half * 151:], nreads - half)
    size = os.path.getsize(f1) + os.path.getsize(f2)
    subprocess.run(["gzip", "-1", "-k", f1, f2], check=True)
    positions = bench.oracle_positions(None, keys_np, lines[: (half + (nreads - half)) * 151])  # numpy, run-length rule
    threads = int(os.environ.get("THREADS", str(min(16, os.cpu_count() or 1))))
    res = {}
    for name, paths, env in (("plain_device_parse", [f1, f2], "1"), ("plain_kseq", [f1, f2], "0"),
                             ("gzip_kseq", [f1 + ".gz", f2 + ".gz"], "1")):
        os.environ["VG_RAW_FASTQ"] = env
        from varigraph_b200 import capi
        ctx = capi.Context(0, buffer_mb=64)
        ix = capi.Index(ctx, keys_np, 27)
        best, counts = None, None
        for rep in range(3):
            ix.begin()
            t0 = time.perf_counter()
            rb = ix.count_files(paths, threads=threads)
            counts, pos_, hits = ix.end()
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
            assert pos_ == positions and rb == nreads * 150, (pos_, positions, rb)
        res[name] = counts
        print(json.dumps({"road": name, "kmer_positions_per_s": positions / best, "seconds": best, "threads": threads,
                          "fastq_bytes": size, "reads": int(nreads), "raw_blocks": ix.fastq_blocks}), flush=True)
        ix.close()
        ctx.close()
    assert np.array_equal(res["plain_device_parse"], res["plain_kseq"]) and np.array_equal(res["plain_kseq"], res["gzip_kseq"])
    for f in os.listdir(tmp):
        os.remove(os.path.join(tmp, f))
    os.rmdir(tmp)


if __name__ == "__main__":
    main()
