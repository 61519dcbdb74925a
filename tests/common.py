"""Shared helpers of the parity tests: seeded inputs and comparisons (ints exact, doubles to 1e-6)."""
import numpy as np

from rsicnv_b200 import api, synth

INT_FIELDS = ("tid", "type", "geno", "status", "start", "end", "length", "sc1", "sc2", "pair", "rp")
DBL_FIELDS = ("score", "p1", "p2", "cnvmed", "cnvsd", "cnviqr", "refmed", "refsd", "refiqr", "q0")
RTOL = 1e-6   # north_star: "within a stated relative tolerance (e.g. 1e-6) for SCORE and the SD/NB-transformed floats"


def make_case(L, seed, stress=False, n_events=8, lens=(2000, 5000, 10000, 30000), mean=30.0):
    fa = synth.make_fasta(L, seed)
    if stress:
        ev = synth.stress_events(L, seed, synth._n_runs(fa))
        d, _ = synth.make_depth(L, seed, fa, events=ev, mean=mean)
    else:
        d, ev = synth.make_depth(L, seed, fa, n_events=n_events, lens=lens, mean=mean)
    return fa, d, ev


def oracle_params(kw):
    out = {}
    for k, v in kw.items():
        out[k] = (1 if v is True else 0 if v is False else v)
    return out


def assert_calls_equal(got, want, what=""):
    assert len(got) == len(want), f"{what}: {len(got)} calls vs {len(want)}"
    for i, (a, b) in enumerate(zip(got, want)):
        for f in INT_FIELDS:
            assert getattr(a, f) == getattr(b, f), f"{what} call {i} field {f}: {getattr(a, f)} != {getattr(b, f)}"
        for f in DBL_FIELDS:
            x, y = getattr(a, f), getattr(b, f)
            assert x == y or abs(x - y) <= RTOL * max(abs(x), abs(y)), f"{what} call {i} field {f}: {x} != {y}"


def run_depth_case(lib_path, oracle, fa, d, check_bins=True, level0_modes=(2,), **kw):
    """the whole depth path through the C ABI vs the oracle; returns the calls"""
    oracle.set_params(**oracle_params(kw))
    ro = oracle.depth_path(d, fa, 3, want_bins=True)
    ro_lists = [oracle.last_list(k) for k in range(4)]
    ro1 = None
    ctx = api.Context(lib=lib_path, **kw)
    try:
        ctx.set_reference(fa)
        ctx.set_depth(d)
        calls = None
        for mode in level0_modes:
            ctx.set_level0_mode(mode)
            calls = ctx.run()
            st = ctx.chr_stats()
            nb = st.nbins
            assert st.rdmedian == ro["stats"][0] and st.rdsd == ro["stats"][1], (st.rdmedian, st.rdsd, ro["stats"])
            if check_bins:
                if ro1 is None:
                    oracle.set_params(**oracle_params(kw))
                    ro1 = oracle.depth_path(d, fa, 1)
                assert np.array_equal(ctx.array(api.ARR_DEPTH), ro1["depth"]), "compacted depth"
                bm, bn, bi, bs = ro["bins"]
                assert np.array_equal(ctx.array(api.ARR_BIN_MED), bm[:nb]), "median_transfer"
                assert np.array_equal(ctx.array(api.ARR_BIN_MEDINT), bi[:nb]), "RDmedint"
                assert np.array_equal(ctx.array(api.ARR_BIN_NBN), bn[:nb]), "negative_binomial_transfer"
                assert np.array_equal(ctx.array(api.ARR_BIN_STATUS), bs[:nb]), "RSI status"
            assert_calls_equal(calls, ro["calls"], f"level0_mode={mode}")
            # the intermediate lists: segments leaving rsicnv*, after areblockscnv, before / after mergesegments
            for k, name in enumerate(("segments", "blocks", "premerge", "merged")):
                assert_calls_equal(ctx.array(api.ARR_SEGMENTS + k), ro_lists[k], f"{name} list, level0_mode={mode}")
        return calls, ctx.launch_count()
    finally:
        ctx.close()


def split_reads(reads, cuts):
    """position-sorted SoA -> consecutive batches with batch-relative offsets (what a BAM decoder would push)"""
    co, qo = reads["cigar_off"], reads["qual_off"]
    out = []
    edges = [0] + list(cuts) + [len(reads["pos"])]
    for lo, hi in zip(edges[:-1], edges[1:]):
        d = {k: reads[k][lo:hi] for k in ("pos", "mpos", "isize", "mtid", "flag", "mapq")}
        d["cigar_off"] = (co[lo:hi + 1] - co[lo]).astype(np.uint32); d["cigar"] = reads["cigar"][co[lo]:co[hi]]
        d["qual_off"] = (qo[lo:hi + 1] - qo[lo]).astype(np.uint64); d["qual"] = reads["qual"][qo[lo]:qo[hi]]
        out.append(d)
    return out


def run_bam_case(lib_path, oracle, fa, reads, n_batches=3, **kw):
    """the whole BAM path through the C ABI (pileup -> load_finish -> detectcnv -> sd_filters -> cnv_stat) vs the oracle"""
    from bind import oracle_bam_path, oracle_isize
    want = oracle_bam_path(oracle, reads, fa, **oracle_params(kw))
    n = len(reads["pos"])
    cuts = [n * (k + 1) // n_batches for k in range(n_batches - 1)]
    ctx = api.Context(lib=lib_path, **kw)
    try:
        ctx.set_reference(fa)
        ctx.pileup_begin()
        for b in split_reads(reads, cuts):
            ctx.pileup_push(b)
        ctx.pileup_end()
        assert np.array_equal(ctx.array(api.ARR_RAW_DEPTH), want["raw"]), "pileup depth"
        calls = ctx.run()
        st = ctx.chr_stats()
        assert st.rdmedian == want["stats"][0] and st.rdsd == want["stats"][1]
        assert_calls_equal(calls, want["calls"], "bam path")
        if calls:
            assert (st.isize_mean, st.isize_sd) == oracle_isize(oracle, reads, len(fa))
        return calls, st
    finally:
        ctx.close()
