"""BASELINE config 5 in miniature, through the drop-in binary: S samples genotyped against one graph, timed with the
samples counted one after the other on one GPU (--gpu 0), dealt over two GPU contexts while the host genotypes in list
order (--gpu a,b: the pipeline of host/varigraph_b200.hpp), and by the reference CPU binary.  VCFs compared byte for
byte.  Usage (on a GPU box): python tools/config5_timing.py [samples] [genome_len] [nvar] [coverage]"""
import gzip
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from tests import oracle_binding as ob  # noqa: E402  (only to locate oracle/_ref)
from varigraph_b200 import synth  # noqa: E402


def run(cmd, cwd, env=None):
    t0 = time.perf_counter()
    r = subprocess.run(cmd, cwd=cwd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, env={**os.environ, **(env or {})})
    if r.returncode != 0:
        sys.stderr.write(r.stdout.decode(errors="replace")[-3000:])
        raise SystemExit(f"{cmd[0]} failed")
    return time.perf_counter() - t0


def main():
    S = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    L = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    nvar = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
    cov = float(sys.argv[4]) if len(sys.argv) > 4 else 30.0
    ref_bin, b200 = ob.REF_BIN, os.path.join(ob.ORACLE_DIR, "_ref", "varigraph_b200")
    tmp = tempfile.mkdtemp(prefix="vgcfg5_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    g = synth.make_genome(L, 41)
    v = synth.make_variants(g, nvar, 5, 2, 42)
    fa, vcf, graph = os.path.join(tmp, "ref.fa"), os.path.join(tmp, "var.vcf"), os.path.join(tmp, "graph.bin")
    synth.write_fasta(fa, g)
    synth.write_vcf(vcf, v, len(g))
    run([ref_bin, "construct", "-r", fa, "-v", vcf, "--save-graph", graph, "-t", "8"], tmp)
    cfg = []
    for s in range(S):
        haps = [synth.apply_haplotype(g, v, s % 5, h) for h in range(2)]
        m1, m2 = synth.make_reads(haps, cov, len(g), seed=100 + s)
        f1, f2 = os.path.join(tmp, f"S{s}_1.fq"), os.path.join(tmp, f"S{s}_2.fq")
        synth.write_fastq(f1, m1, "a")
        synth.write_fastq(f2, m2, "b")
        cfg.append(f"S{s} {f1} {f2}")
    cfgp = os.path.join(tmp, "samples.cfg")
    open(cfgp, "w").write("\n".join(cfg) + "\n")
    two = ("0,1", {}) if torch.cuda.device_count() >= 2 else ("0,0", {"VG_ALLOW_SAME_DEVICE": "1"})
    arms = [("reference_cpu_8_threads", ref_bin, [], {}), ("b200_one_gpu", b200, ["--gpu", "0", "--buffer", "64"], {}),
            ("b200_samples_dealt_over_" + two[0].replace(",", "_and_"), b200, ["--gpu", two[0], "--buffer", "64"], two[1])]
    out, vcfs = {}, {}
    for name, exe, extra, env in arms:
        d = os.path.join(tmp, name)
        os.mkdir(d)
        out[name + "_s"] = run([exe, "genotype", "--load-graph", graph, "-s", cfgp, "-t", "8", *extra], d, env)
        vcfs[name] = [gzip.open(os.path.join(d, f"S{s}.varigraph.vcf.gz"), "rb").read() for s in range(S)]
    names = [a[0] for a in arms]
    out["vcfs_identical_to_reference"] = all(vcfs[n] == vcfs[names[0]] for n in names[1:])
    out.update(samples=S, genome_len=L, variants=nvar, coverage=cov, gpus_visible=torch.cuda.device_count(),
               note="wall clock of the whole `genotype` run: graph load + index build + per sample (count, genotype, VCF)")
    print(json.dumps(out))
    subprocess.run(["rm", "-rf", tmp])


if __name__ == "__main__":
    main()
