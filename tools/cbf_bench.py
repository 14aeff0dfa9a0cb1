#!/usr/bin/env python
"""Construct side (K4): counting-Bloom-filter fill of one 250 Mb chromosome through vg_cbf_add_sequence.
Prints one JSON line: genome k-mers/s and the fraction of the random-sector rate the 7 cell updates per k-mer reach
(B_cbf = 1 + 7 x (32 + 32) bytes per genome k-mer)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from varigraph_b200 import capi, synth  # noqa: E402

L, K = 250_000_000, 27
g = synth.make_genome(L, seed=99)
n = L - K + 1
m = int(np.ceil(-n * np.log(0.01) / (np.log(2) ** 2)))  # BloomFilter::_calculate_size (src/counting_bloom_filter.cpp:70-73)
seeds = np.random.default_rng(5).integers(1, 2**63, size=7, dtype=np.uint64)
ctx = capi.Context(0, buffer_mb=64)
times = []
for rep in range(3):
    cbf = capi.CountingBloom(ctx, m, seeds)
    t0 = time.perf_counter()
    added = cbf.add_sequence(g, K)
    times.append(time.perf_counter() - t0)
    cbf.close()
sec = min(times)
rnd_gbs, rnd_sec = ctx.probe_random_sectors(8 << 30, 64)
print(json.dumps({"genome_bases": L, "k": K, "cells": m, "hashes": 7, "kmers_added": added, "seconds": sec,
                  "kmers_per_s": added / sec, "bytes_per_kmer_alg": 1 + 7 * 64,
                  "achieved_gbs_alg": added * (1 + 7 * 64) / sec / 1e9,
                  "cell_updates_per_s": 7 * added / sec, "random_sector_rate_per_s": rnd_sec,
                  "frac_of_random_sector_rate": 7 * added / sec / rnd_sec,
                  "note": "wall clock of vg_cbf_add_sequence incl. the H2D copy of the chromosome (pageable host memory)"}))
