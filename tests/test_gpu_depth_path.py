"""Parity tests proper: the real library (hand-written CUDA, sm_100a) through the C ABI against the oracle
on the same seeded inputs, plus size-independent properties at BASELINE.json's full contig size."""
import numpy as np
import pytest

from common import assert_calls_equal, make_case, run_depth_case
from rsicnv_b200 import api, synth

pytestmark = pytest.mark.gpu

CASES = [
    dict(L=400_000, seed=1), dict(L=400_001, seed=2), dict(L=400_019, seed=3),
    dict(L=600_007, seed=4, kw=dict(gcadjust=False)), dict(L=600_007, seed=5, kw=dict(cap=-1.0)), dict(L=600_007, seed=6, kw=dict(trans="MED")),
    dict(L=800_003, seed=7, kw=dict(m=51)), dict(L=1_500_003, seed=8, kw=dict(m=501)),
    dict(L=700_003, seed=9, stress=True), dict(L=500_003, seed=10, kw=dict(merge=False), stress=True),
    dict(L=500_003, seed=39, stress=True), dict(L=500_003, seed=51, stress=True),
    dict(L=600_007, seed=6, kw=dict(trans="MED", threshold=5.0)),   # Lmax = (4*5)^2 = 400 > 256: the large-footprint scan instantiation
    # the undocumented knobs (rsi.cpp:2020-2026): -e (RSI factor), -reflen (neighbourhood length), -maxchkbp (sub-sampling of long neighbourhoods)
    dict(L=700_003, seed=9, kw=dict(epsilon=0.5), stress=True), dict(L=500_003, seed=39, kw=dict(chklen=1.5), stress=True),
    dict(L=700_003, seed=9, kw=dict(maxchkbp=500), stress=True), dict(L=600_007, seed=5, kw=dict(epsilon=3.0, chklen=4.0, maxchkbp=2000)),
    dict(L=600_007, seed=6, kw=dict(trans="ALL")), dict(L=700_003, seed=9, kw=dict(trans="ALL"), stress=True),
    dict(L=3_000_017, seed=5), dict(L=12_000_003, seed=21, stress=True), dict(L=12_000_003, seed=22, kw=dict(trans="MED")),
    dict(L=6_000_011, seed=23, kw=dict(m=51), stress=True), dict(L=6_000_000, seed=24, kw=dict(gcadjust=False, cap=-1.0)),
]


@pytest.mark.parametrize("case", CASES, ids=lambda c: f"L{c['L']}-s{c['seed']}-{'-'.join(f'{k}{v}' for k, v in c.get('kw', {}).items()) or 'default'}")
def test_depth_path_matches_oracle(case, gpu_lib, oracle):
    fa, d, _ = make_case(case["L"], case["seed"], stress=case.get("stress", False))
    calls, launches = run_depth_case(gpu_lib, oracle, fa, d, level0_modes=(2, 1, 0), **case.get("kw", {}))
    assert launches > 0


def test_repeat_runs_are_identical(gpu_lib, oracle):
    """the path is deterministic: two runs on one context and a run on a fresh context give identical bytes"""
    fa, d, _ = make_case(3_000_017, 31, stress=True)
    outs = []
    for fresh in range(2):
        with api.Context(lib=gpu_lib) as ctx:
            ctx.set_reference(fa); ctx.set_depth(d)
            for _ in range(2):
                calls = ctx.run()
                outs.append((bytes(b"".join(bytes(c) for c in calls)), ctx.array(api.ARR_BIN_STATUS).tobytes(), ctx.array(api.ARR_DEPTH).tobytes()))
    assert all(o == outs[0] for o in outs[1:])


def test_chr19_full_size_properties(gpu_lib):
    """59,128,983 bp (BASELINE.json config 1 size): properties that do not need the CPU oracle at full size"""
    L = synth.CHR19_LEN
    fa = synth.make_fasta(L, 19)
    d, ev = synth.make_depth(L, 19, fa, n_events=20)
    with api.Context(lib=gpu_lib, gcadjust=False, cap=-1.0) as ctx:
        # no GC adjust, no cap: the compacted depth must be the input with the N intervals cut out
        ctx.set_reference(fa); ctx.set_depth(d); ctx.load_finish()
        nb, ne = ctx.array(api.ARR_NOSEQ_BEG), ctx.array(api.ARR_NOSEQ_END)
        keep = np.ones(L, bool)
        for b, e in zip(nb, ne):
            keep[b:e + 1] = False
        want = d[keep]
        got = ctx.array(api.ARR_DEPTH)
        assert np.array_equal(got, want)
        st = ctx.chr_stats()
        assert st.compact_len == len(want)
        # hist-median of ints == the rank-n/2 order statistic (1-based), SD from exact integer sums
        assert st.rdmedian == float(np.sort(want)[len(want) // 2 - 1])
        m1 = want.astype(np.float64).sum() / len(want); m2 = (want.astype(np.float64) ** 2).sum() / len(want)
        assert abs(st.rdsd - np.sqrt(m2 - m1 * m1)) < 1e-9
        # bin medians / sums against numpy
        m = 101; nbins = len(want) // m
        bins = want[:nbins * m].reshape(nbins, m)
        assert np.array_equal(ctx.array(api.ARR_BIN_MEDINT), np.sort(bins, axis=1)[:, m // 2])
    with api.Context(lib=gpu_lib) as ctx:
        ctx.set_reference(fa); ctx.set_depth(d)
        calls = ctx.run()
        # level-0 float chain: block-scan form == sequential FADD chain at full size
        s1 = ctx.array(api.ARR_BIN_STATUS1).copy(); s2 = ctx.array(api.ARR_BIN_STATUS).copy()
        ctx.set_level0_mode(0)
        calls0 = ctx.run()
        assert np.array_equal(s1, ctx.array(api.ARR_BIN_STATUS1)) and np.array_equal(s2, ctx.array(api.ARR_BIN_STATUS))
        assert_calls_equal(calls, calls0, "level0 scan vs sequential")
        # every planted event that is not the last marked run must be recovered with >= 50% reciprocal overlap (checkbp.pl rule)
        hit = 0
        for s, e, f in ev[:-1]:
            for c in calls:
                ov = min(e, c.end) - max(s, c.start)
                if ov > 0.5 * (e - s) and ov > 0.5 * (c.end - c.start) and c.type == (0 if f < 1 else 1):
                    hit += 1
                    break
        assert hit >= len(ev) - 3, (hit, len(ev))


def test_high_depth_histogram_windows(gpu_lib, oracle):
    fa, d, _ = make_case(3_000_003, 17, mean=200.0)
    run_depth_case(gpu_lib, oracle, fa, d)


def test_level0_float_chain_modes_agree_at_scale(gpu_lib):
    """MED transform at 60x over chr19: integer-valued bins push the float accumulator past 2^24 (ulp 2, 4), where exact
    ties (round-to-even) occur on most additions -- the three implementations of the sequential float sum must agree bit for bit"""
    L = synth.CHR19_LEN
    fa = synth.make_fasta(L, 21)
    d, _ = synth.make_depth(L, 21, fa, n_events=20, mean=60.0)
    outs = []
    with api.Context(lib=gpu_lib, trans="MED") as ctx:
        ctx.set_reference(fa); ctx.set_depth(d)
        for mode in (0, 1, 2):
            ctx.set_level0_mode(mode)
            calls = ctx.run()
            ds = ctx.debug_state()
            outs.append((ctx.array(api.ARR_BIN_STATUS1).tobytes(), ctx.array(api.ARR_BIN_STATUS).tobytes(), b"".join(bytes(c) for c in calls), ds["tmedian"], ds["tlamda"]))
    assert outs[0] == outs[1] == outs[2]
