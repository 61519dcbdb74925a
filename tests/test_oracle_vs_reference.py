"""Pins the oracle (oracle/rsi_oracle.cpp, a CPU restatement) against the UNMODIFIED reference:
 * always: the committed golden vectors (tests/golden/reference_vectors.json, made by tests/golden/make_golden.py
   from the reference's own functions and CLI in the build container);
 * where oracle/_ref exists (built in place from /root/reference): live, function by function."""
import hashlib
import json
import os

import numpy as np
import pytest

from bind import oracle_bam_path, oracle_pileup
from common import DBL_FIELDS, INT_FIELDS, make_case, oracle_params
from rsicnv_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "reference_vectors.json")))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.mark.parametrize("g", GOLD["depth_cases"], ids=lambda g: f"L{g['L']}-s{g['seed']}")
def test_depth_path_golden(g, oracle):
    fa, d, _ = make_case(g["L"], g["seed"], stress=g["stress"])
    assert sha(d) + sha(fa) == g["input_sha"], "synthetic input generator drifted: regenerate the goldens"
    oracle.set_params(**oracle_params(g["kw"]))
    res1 = oracle.depth_path(d, fa, 1)
    assert len(res1["depth"]) == g["n_compact"] and sha(res1["depth"]) == g["depth_sha"]
    oracle.set_params(**oracle_params(g["kw"]))
    res = oracle.depth_path(d, fa, 3)
    assert res["stats"][0] == g["rdmedian"] and res["stats"][1] == g["rdsd"]
    rows = [oracle.format_row(x, "19", res["stats"][0], res["stats"][1]) for x in res["calls"]]
    assert rows == g["rows"]
    for x, want in zip(res["calls"], g["calls"]):
        for f in INT_FIELDS:
            assert getattr(x, f) == want[f], f
        for f in DBL_FIELDS:
            assert getattr(x, f) == want[f], f     # the oracle reproduces the reference's doubles bit for bit


def test_l0_golden(oracle):
    rng = np.random.default_rng(77)
    xi = rng.poisson(30, 5001).astype(np.int32); xf = (rng.gamma(9, 3.3, 4000)).astype(np.float32)
    g = GOLD["l0"]
    assert oracle.median(xi) == g["median_i32"] and oracle.iqr(xi) == g["iqr_i32"]
    assert oracle.median(xf) == g["median_f32"] and oracle.iqr(xf) == g["iqr_f32"]
    assert oracle.true_median(xi[:100]) == g["true_median_even"] and oracle.true_median(xi[:101]) == g["true_median_odd"]
    assert oracle.variance(xi) == g["variance_i32"] and oracle.variance(xf) == g["variance_f32"]
    for v, want in zip((-9.5, -3.3, -0.7, -0.2, 0.0, 0.31, 0.49, 0.5, 2.2, 7.1, 11.0), g["pnorm"]):
        assert oracle.pnorm(v) == want


def test_bam_path_golden(oracle):
    """pileup depth == the reference's `-s` dump, table == the reference CLI's table (RP / Q0 included)"""
    b = GOLD["bam_case"]
    fa = synth.make_fasta(b["L"], b["seed"])
    reads, _ = synth.make_reads(b["L"], b["seed"], fa, coverage=b["coverage"], n_events=b["n_events"], lens=tuple(b["lens"]))
    assert sha(reads["pos"]) + sha(reads["qual"]) == b["reads_sha"], "synthetic read generator drifted: regenerate the goldens"
    res = oracle_bam_path(oracle, reads, fa, minq=b["minq"], min_baseQ=b["min_baseQ"])
    assert sha(res["raw"]) == b["raw_depth_sha"]
    rows = [oracle.format_row(x, "1", res["stats"][0], res["stats"][1]) for x in res["calls"]]
    want = [ln for ln in b["table"] if not ln.startswith("#")]
    assert rows == want


# ---- live comparison against the reference objects (this container only)
def test_l0_live(oracle, ref):
    rng = np.random.default_rng(3)
    for n in (2, 3, 10, 101, 5000):
        xi = rng.poisson(25, n).astype(np.int32); xf = rng.gamma(8, 3, n).astype(np.float32)
        assert oracle.median(xi) == ref.median(xi) and oracle.iqr(xi) == ref.iqr(xi)
        assert oracle.median(xf) == ref.median(xf) and oracle.iqr(xf) == ref.iqr(xf)
        assert oracle.true_median(xi) == ref.true_median(xi)
        assert oracle.variance(xi) == ref.variance(xi)
    for v in np.linspace(-12, 12, 97):
        assert oracle.pnorm(float(v)) == ref.pnorm(float(v))
    xc = np.full(50, 7, np.int32)
    assert oracle.median(xc) == ref.median(xc)


@pytest.mark.parametrize("L", [100_000, 100_001, 100_003, 100_007, 100_019])
def test_gc_cap_compact_live(L, oracle, ref):
    fa, d, _ = make_case(L, L % 97)
    gc = ((fa == ord("G")) | (fa == ord("C"))).astype(np.uint8)
    for lib in (oracle, ref):
        lib.set_params()
    a, b = oracle.checkgccontent(d, gc), ref.checkgccontent(d, gc)
    assert np.array_equal(a, b)
    nb_o, ne_o = oracle.noseq_regions(fa); nb_r, ne_r = ref.noseq_regions(fa)
    assert np.array_equal(nb_o, nb_r) and np.array_equal(ne_o, ne_r)


@pytest.mark.parametrize("case", [dict(L=500_003, seed=12, stress=True), dict(L=900_001, seed=14, kw=dict(trans="MED")), dict(L=700_001, seed=15, kw=dict(m=51)),
                                  dict(L=700_003, seed=9, stress=True, kw=dict(epsilon=0.5)), dict(L=500_003, seed=39, stress=True, kw=dict(chklen=1.5)),
                                  dict(L=700_003, seed=9, stress=True, kw=dict(maxchkbp=500)), dict(L=600_007, seed=5, kw=dict(epsilon=3.0, chklen=4.0, maxchkbp=2000))],
                         ids=lambda c: f"L{c['L']}-s{c['seed']}-{'-'.join(f'{k}{v}' for k, v in c.get('kw', {}).items()) or 'default'}")
def test_depth_path_live(case, oracle, ref):
    fa, d, _ = make_case(case["L"], case["seed"], stress=case.get("stress", False))
    kw = oracle_params(case.get("kw", {}))
    oracle.set_params(**kw); ref.set_params(**kw)
    a, b = oracle.depth_path(d, fa, 3), ref.depth_path(d, fa, 3)
    assert np.array_equal(a["depth"], b["depth"]) and a["stats"] == b["stats"]
    assert [x.as_dict() for x in a["calls"]] == [x.as_dict() for x in b["calls"]]


def test_bam_decode_matches_reference_samtools(tmp_path):
    """oracle/bam_decode.py (BGZF + bam_read1 restated) against the reference's own samtools on a real-looking file"""
    from bind import BAM_FIELDS, have_ref, oracle_bam_decode, ref_bam_records
    from rsicnv_b200 import synth
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    r0, _ = synth.make_reads(150_000, 71, None, coverage=8, n_events=0, tid=0, frac_indel=0.3)
    r1, _ = synth.make_reads(90_000, 72, None, coverage=6, n_events=0, tid=1)
    path = str(tmp_path / "t.bam")
    synth.write_bam(path, [("1", 150_000), ("2", 90_000)], {0: r0, 1: r1}, level=6, rich=5, unmapped_tail=11)
    d = oracle_bam_decode(np.fromfile(path, np.uint8))
    assert d["order"] == [0, 1, -1]
    for tid, src in ((0, r0), (1, r1), (-1, None)):
        R = ref_bam_records(path, tid)
        for k, dt in BAM_FIELDS:
            assert np.array_equal(R[k], d["reads"][tid][k]), (tid, k)
            if src is not None:
                assert np.array_equal(R[k].astype(np.int64), np.asarray(src[k]).astype(np.int64)), (tid, k)
