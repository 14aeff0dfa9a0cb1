"""CPU suite, part 1: the oracle (oracle/vg_oracle.c) is pinned against the reference.

Level T0: primitives against tests/golden/primitives.json, which
oracle/gen_golden.py produced by calling the unmodified reference; level T1: per-k-mer counts
against tests/golden/tiny (FastqKmer::build_fastq_index on a reference-built graph.bin).
Where oracle/_ref/libvgref.so exists the same checks also run live against the reference.
"""
import random

import numpy as np
import pytest

from tests import helpers


def test_hash64_golden(oracle):
    for case in helpers.primitives()["hash64"]:
        mask = (1 << (2 * case["k"])) - 1
        for x, y in zip(case["x"], case["y"]):
            assert oracle.hash64(int(x), mask) == int(y)


def test_nt4_golden(oracle):
    table = helpers.primitives()["nt4"]
    assert [int(oracle.lib.vgo_nt4(b)) for b in range(256)] == table


def test_murmur_golden(oracle):
    for x, seed, y in helpers.primitives()["murmur"]:
        assert int(oracle.lib.vgo_murmur3_x64_128_sum(int(x), seed)) == int(y)


def test_cbf_sizing_golden(oracle):
    rows = helpers.primitives()["cbf_sizing"]
    assert rows
    for n, p, m, nh in rows:
        assert int(oracle.lib.vgo_cbf_size(n, p)) == int(m)
        assert int(oracle.lib.vgo_cbf_num_hashes(n, int(m))) == nh
    # human-scale sizing from SURVEY appendix C: m = ceil(9.58506 n), 7 hashes
    n = 3_099_999_974
    m = int(oracle.lib.vgo_cbf_size(n, 0.01))
    assert abs(m / n - 9.58506) < 1e-4 and int(oracle.lib.vgo_cbf_num_hashes(n, m)) == 7


def test_sketch_golden(oracle):
    cases = helpers.primitives()["sketch"]
    assert len(cases) > 400
    for c in cases:
        got = oracle.sketch(c["seq"].encode("latin1"), c["k"])
        assert [str(int(x)) for x in got] == c["keys"], (c["k"], c["seq"])


def test_positions_consistent_with_sketch(oracle):
    rng = random.Random(3)
    for k in (4, 5, 27, 28):
        reads = ["".join(rng.choice("ACGTN" if i % 3 else "AT") for _ in range(rng.randint(1, 200)))
                 for i in range(40)]
        buf = ("\n".join(reads) + "\n").encode()
        pos = oracle.positions(buf, k)
        want = np.concatenate([oracle.sketch(r.encode(), k) for r in reads] + [np.zeros(0, np.uint64)])
        assert np.array_equal(pos[pos != helpers.NOKMER], want)


def test_counts_golden_tiny(oracle):
    t = helpers.tiny()
    counts, positions, hits = oracle.count_lines(t["keys"], t["lines"], t["k"])
    assert np.array_equal(counts, t["counts"])
    assert hits == int(t["counts"].astype(np.int64).sum())  # no counter saturates in this fixture
    assert t["lines"].size - (t["m1"].shape[0] + t["m2"].shape[0]) == t["read_bases"]


def test_cbf_golden(oracle):
    g = helpers.cbf_golden()
    m, k, seeds = int(g["m"]), int(g["k"]), g["seeds"]
    assert m == int(oracle.lib.vgo_cbf_size(len(g["genome"]) - k + 1, 0.01))
    filt, n = oracle.cbf_fill(m, seeds, g["genome"], k)
    assert np.array_equal(filt, g["filter"])
    assert int(filt.max()) == 255  # the fixture exercises saturation
    for key, c, f in zip(g["probe"], g["probe_count"], g["probe_find"]):
        assert oracle.cbf_count(filt, m, seeds, key) == int(c)
        assert oracle.cbf_find(filt, m, seeds, key) == int(f)


def test_fastq_parser_cases(oracle):
    want = {
        "plain": (1, 33, -1), "crlf": (1, 33, -1), "multiline": (2, 66, -1), "fasta": (2, 99, -1),
        "qual_at": (2, 66, -1), "truncated_qual": (1, 33, -2), "no_final_newline": (1, 33, -1),
        "lower_n": (1, 53, -1), "leading_junk": (1, 33, -1),
    }
    for name, text in helpers.EDGE_FASTQS.items():
        lines, nreads, bases, status = oracle.fastq_to_lines(text)
        assert (nreads, bases, status) == want[name], name
        assert lines.count(b"\n") == nreads and b"\r" not in lines, name
    assert oracle.fastq_to_lines(b"") == (b"", 0, 0, -1)


# ---- live against the unmodified reference (only where oracle/_ref was built) -------------------
def test_sketch_vs_reference_live(oracle, reference):
    rng = random.Random(11)
    for k in (4, 6, 8, 15, 16, 27, 28):
        for t in range(150):
            alpha = ["ACGT", "ACGTNacgtnU", "AT", "GC"][t % 4]
            s = "".join(rng.choice(alpha) for _ in range(rng.randint(1, 300))).encode()
            assert np.array_equal(oracle.sketch(s, k), reference.sketch(s, k)), (k, s)


def test_counts_vs_reference_live(oracle, reference, tmp_path):
    t = helpers.tiny()
    graph = tmp_path / "graph.bin"
    graph.write_bytes(t["graph_bin"])
    f1, f2 = helpers.write_tiny_fastqs(str(tmp_path), t)
    h, keys, k = reference.graph_load(str(graph))
    try:
        counts, read_bases, _ = reference.count_files(h, keys.size, [f1, f2], threads=3)
    finally:
        reference.graph_destroy(h)
    assert k == t["k"] and read_bases == t["read_bases"]
    order = np.argsort(keys)
    assert np.array_equal(keys[order], t["keys"]) and np.array_equal(counts[order], t["counts"])
    mine, _, _ = oracle.count_lines(keys, t["lines"], k)
    assert np.array_equal(mine, counts)


def test_fastq_parser_vs_reference_live(oracle, reference, tmp_path):
    """kseq semantics: the reference reads each crafted file; the oracle parser must stage the same reads."""
    t = helpers.tiny()
    graph = tmp_path / "graph.bin"
    graph.write_bytes(t["graph_bin"])
    h, keys, k = reference.graph_load(str(graph))
    try:
        # sequences that certainly hit the index: windows of haplotype reads
        body = t["m1"][:40]
        for name, text in helpers.EDGE_FASTQS.items():
            if name == "lower_n":
                continue
            recs = []
            for i, r in enumerate(body):
                s = r.tobytes()
                if name == "crlf":
                    recs.append(b"@q%d c\r\n%s\r\n+\r\n%s\r\n" % (i, s, b"I" * len(s)))
                elif name == "multiline":
                    recs.append(b"@q%d\n%s\n%s\n+\n%s\n%s\n" % (i, s[:70], s[70:], b"I" * 70, b"I" * (len(s) - 70)))
                elif name == "fasta":
                    recs.append(b">q%d\n%s\n%s\n" % (i, s[:70], s[70:]))
                elif name == "qual_at":
                    recs.append(b"@q%d\n%s\n+\n@%s\n" % (i, s, b"I" * (len(s) - 1)))
                elif name == "truncated_qual" and i == 25:
                    recs.append(b"@q%d\n%s\n+\n%s\n" % (i, s, b"I" * 10))
                elif name == "leading_junk" and i == 0:
                    recs.append(b"garbage\n\n@q%d\n%s\n+\n%s\n" % (i, s, b"I" * len(s)))
                else:
                    recs.append(b"@q%d\n%s\n+\n%s\n" % (i, s, b"I" * len(s)))
            data = b"".join(recs)
            if name == "no_final_newline":
                data = data[:-1]
            path = tmp_path / (name + ".fq")
            path.write_bytes(data)
            ref_counts, ref_bases, _ = reference.count_files(h, keys.size, [str(path)], threads=2)
            lines, nreads, bases, status = oracle.fastq_to_lines(data)
            mine, _, _ = oracle.count_lines(keys, lines, k)
            assert bases == ref_bases, name
            assert np.array_equal(mine, ref_counts), name
    finally:
        reference.graph_destroy(h)


def test_fastq_parser_vs_reference_random_damage(oracle, reference, tmp_path):
    """Differential test of the kseq restatement: well-formed FASTQ with random damage (lost or doubled
    newlines, stray '@' '+' '>' NUL CR bytes, cut tails) -- the reference reads each file, the oracle
    parser must stage exactly the reads the reference counted (same per-k-mer counts, same mReadBase)."""
    import random
    t = helpers.tiny()
    graph = tmp_path / "graph.bin"
    graph.write_bytes(t["graph_bin"])
    h, keys, k = reference.graph_load(str(graph))
    rng = random.Random(2026)
    try:
        body = [r.tobytes() for r in t["m1"][:30]]
        checked = 0
        for trial in range(60):
            recs = [b"@q%d\n%s\n+\n%s\n" % (i, s, bytes(rng.choice(b"I@+>#") for _ in s)) for i, s in enumerate(body)]
            data = bytearray(b"".join(recs))
            for _ in range(rng.randint(1, 4)):
                at = rng.randrange(len(data))
                kind = rng.randrange(6)
                if kind == 0:                       # lose a newline (or any byte)
                    nl = data.find(b"\n", at)
                    if nl >= 0:
                        del data[nl]
                elif kind == 1:                     # an extra newline
                    data.insert(at, 10)
                elif kind == 2:                     # a stray marker or control byte
                    data.insert(at, rng.choice(b"@+>\r\x00 "))
                elif kind == 3:                     # overwrite with one
                    data[at] = rng.choice(b"@+>\rN\x00")
                elif kind == 4:                     # cut the tail
                    del data[max(at, len(data) // 2):]
                else:                               # CRLF line end
                    nl = data.find(b"\n", at)
                    if nl >= 0:
                        data.insert(nl, 13)
            data = bytes(data)
            lines, nreads, bases, status = oracle.fastq_to_lines(data)
            if lines.startswith(b"\n") or b"\n\n" in lines:
                continue  # an empty read: the reference itself aborts there (assert len > 0, src/kmer.cpp:124)
            checked += 1
            path = tmp_path / ("dmg%d.fq" % trial)
            path.write_bytes(data)
            ref_counts, ref_bases, _ = reference.count_files(h, keys.size, [str(path)], threads=2)
            mine, _, _ = oracle.count_lines(keys, lines, k)
            assert bases == ref_bases, (trial, data[:200])
            assert np.array_equal(mine, ref_counts), trial
        assert checked >= 30
    finally:
        reference.graph_destroy(h)


def test_even_k_closed_form_rule(oracle):
    """The rule the even-k window encoder rests on (varigraph_b200/csrc/vg_device.cuh, EvenEncoder), checked against the
    state machine (src/kmer.cpp:126-146) without a GPU: outside the reach of a palindrome of the registers spliced over
    an ambiguous byte (its 16-byte segment and the four behind it) and of a clean palindrome (its segment and the two
    behind it), position i emits iff the k bytes ending at i are valid and do not read the same on both strands."""
    import random
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    rng = random.Random(7)
    for k in (4, 8, 12, 16, 20, 28):
        checked = 0
        for trial in range(60):
            parts = []
            for _ in range(rng.randint(1, 6)):
                n = rng.randint(0, 150)
                mode = rng.random()
                if mode < 0.3:
                    s = [rng.choice("AT") for _ in range(n)]
                elif mode < 0.4:
                    s = list(("AT" * n)[:n])
                else:
                    s = [rng.choice("ACGT") for _ in range(n)]
                for j in range(len(s)):
                    if rng.random() < 0.01:
                        s[j] = rng.choice("NnX")
                parts.append("".join(s))
            text = "\n".join(parts) + "\n"
            n = len(text)
            pos = oracle.positions(text.encode(), k)
            truth = {i for i in range(n) if pos[i] != helpers.NOKMER}
            valid = [c in comp for c in text]
            hard = [c == "\n" for c in text]

            def V(i):
                return 0 <= i < n and valid[i]

            def allk(i):
                return all(V(j) for j in range(i - k + 1, i + 1))

            def ispal(w):
                return all(w[d] == comp[w[k - 1 - d]] for d in range(k // 2))

            nseg = (n + 15) // 16
            stale, clean = [False] * nseg, [False] * nseg
            for sg in range(nseg):
                for t in range(sg * 16, min(sg * 16 + 16, n)):
                    if not V(t):
                        continue
                    if allk(t):
                        clean[sg] |= ispal(text[t - k + 1: t + 1])
                        continue
                    b = t - 1
                    while V(b):
                        b -= 1
                    if b < 0 or hard[b]:
                        continue  # registers zeroed there: still filling up, never equal
                    need = k - (t - b)
                    a = b - 1
                    while a >= b - need and V(a):
                        a -= 1
                    if a >= b - need:  # another invalid byte inside the spliced window
                        if a >= 0 and not hard[a]:
                            stale[sg] = True  # a second ambiguous byte: not worked out, the lane takes the state machine
                        continue
                    stale[sg] |= ispal(text[b - need: b] + text[b + 1: t + 1])
            for sg in range(nseg):
                if any(stale[max(0, sg - 4): sg + 1]) or any(clean[max(0, sg - 2): sg + 1]):
                    continue  # these lanes run the state machine itself
                for i in range(sg * 16, min(sg * 16 + 16, n)):
                    want = allk(i) and not ispal(text[i - k + 1: i + 1])
                    assert want == (i in truth), (k, trial, i)
                    checked += 1
        assert checked > (2000 if k >= 12 else 100), (k, checked)  # (small k: nearly every lane has a palindrome in reach)
