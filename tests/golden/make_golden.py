"""Generates the golden vectors under tests/golden/ by running the UNMODIFIED reference in this container
(oracle/_ref/libref_harness.so and oracle/_ref/rsicnv, built in place from /root/reference by
oracle/Makefile.ref).  The reference ships no fixtures of its own; these files pin the
oracle -- and through it the CUDA path -- on machines where /root/reference does not exist.

    python tests/golden/make_golden.py
"""
import hashlib
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from bind import REF_BAMTOOL, REF_BIN, Lib  # noqa: E402
from common import make_case, oracle_params  # noqa: E402
from rsicnv_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

DEPTH_CASES = [
    dict(L=400_000, seed=1), dict(L=400_001, seed=2), dict(L=400_019, seed=3),
    dict(L=600_007, seed=4, kw=dict(gcadjust=False)), dict(L=600_007, seed=5, kw=dict(cap=-1.0)), dict(L=600_007, seed=6, kw=dict(trans="MED")),
    dict(L=800_003, seed=7, kw=dict(m=51)), dict(L=1_500_003, seed=8, kw=dict(m=501)),
    dict(L=700_003, seed=9, stress=True), dict(L=500_003, seed=10, kw=dict(merge=False), stress=True),
    dict(L=3_000_017, seed=5), dict(L=2_000_003, seed=42),
    dict(L=600_007, seed=6, kw=dict(trans="ALL")), dict(L=700_003, seed=9, kw=dict(trans="ALL"), stress=True), dict(L=500_003, seed=39, stress=True),
]
BAM_CASE = dict(L=10_600_000, seed=3, coverage=12, n_events=6, lens=(3000, 8000, 20000), minq=0, min_baseQ=10)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    r = Lib("ref")
    out = {"depth_cases": [], "l0": {}}
    for c in DEPTH_CASES:
        fa, d, _ = make_case(c["L"], c["seed"], stress=c.get("stress", False))
        kw = c.get("kw", {})
        r.set_params(**oracle_params(kw))
        res = r.depth_path(d, fa, 3)
        chrom = "19"
        rows = [r.format_row(x, chrom, res["stats"][0], res["stats"][1]) for x in res["calls"]]
        calls = [x.as_dict() for x in res["calls"]]
        r.set_params(**oracle_params(kw))
        res1 = r.depth_path(d, fa, 1)
        out["depth_cases"].append(dict(L=c["L"], seed=c["seed"], stress=c.get("stress", False), kw=kw, input_sha=sha(d) + sha(fa),
                                       rdmedian=res["stats"][0], rdsd=res["stats"][1], n_compact=int(len(res1["depth"])),
                                       depth_sha=sha(res1["depth"]), rows=rows, calls=calls))
        print(c, len(rows))
    # L0 known-answer vectors
    rng = np.random.default_rng(77)
    xi = rng.poisson(30, 5001).astype(np.int32); xf = (rng.gamma(9, 3.3, 4000)).astype(np.float32)
    out["l0"] = dict(median_i32=r.median(xi), iqr_i32=r.iqr(xi), median_f32=r.median(xf), iqr_f32=r.iqr(xf), true_median_even=r.true_median(xi[:100]),
                     true_median_odd=r.true_median(xi[:101]), variance_i32=r.variance(xi), variance_f32=r.variance(xf),
                     pnorm=[r.pnorm(v) for v in (-9.5, -3.3, -0.7, -0.2, 0.0, 0.31, 0.49, 0.5, 2.2, 7.1, 11.0)])
    # BAM case through the reference CLI
    b = BAM_CASE
    fa = synth.make_fasta(b["L"], b["seed"])
    reads, ev = synth.make_reads(b["L"], b["seed"], fa, coverage=b["coverage"], n_events=b["n_events"], lens=b["lens"])
    with tempfile.TemporaryDirectory() as td:
        synth.write_fasta(os.path.join(td, "t.fa"), "1", fa)
        synth.write_bam(os.path.join(td, "t.bam"), [("1", b["L"])], {0: reads})
        subprocess.run([REF_BAMTOOL, "index", os.path.join(td, "t.bam")], check=True)
        subprocess.run([REF_BIN, "rsi", "-b", os.path.join(td, "t.bam"), "-f", os.path.join(td, "t.fa"), "-q", str(b["minq"]), "-Q", str(b["min_baseQ"]),
                        "-np", "-s", "-o", os.path.join(td, "out.txt")], check=True, capture_output=True)
        table = [ln for ln in open(os.path.join(td, "out.txt")).read().splitlines() if not ln.startswith("#input")]
        a = np.loadtxt(os.path.join(td, "out.txt.1_rd"), dtype=np.int64)
        rd = np.zeros(b["L"], np.int32); rd[a[:, 0] - 1] = a[:, 1]
    out["bam_case"] = dict(BAM_CASE, lens=list(b["lens"]), table=table, raw_depth_sha=sha(rd), n_reads=int(len(reads["pos"])), reads_sha=sha(reads["pos"]) + sha(reads["qual"]))
    print("bam", len(table))
    json.dump(out, open(os.path.join(HERE, "reference_vectors.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
