"""BAM-input path (pileup, insert-size sample, RP / Q0 counts) of the CUDA sources under the fiber emulator,
against the oracle on the same read batches."""
import numpy as np
import pytest

from common import run_bam_case
from rsicnv_b200 import synth

L = 10_600_000   # every BAM contig must exceed 10 Mbp: bam_rd_pr_stats samples from 10 Mbp on (SURVEY hard part 9)


@pytest.fixture(scope="module")
def case():
    fa = synth.make_fasta(L, 3)
    reads, ev = synth.make_reads(L, 3, fa, coverage=12, n_events=6, lens=(3000, 8000, 20000))
    return fa, reads, ev


def test_bam_path_matches_oracle(case, sim_lib, oracle):
    fa, reads, _ = case
    calls, st = run_bam_case(sim_lib, oracle, fa, reads, n_batches=3, min_baseQ=10, minq=0)
    assert len(calls) >= 3 and any(c.rp > 0 for c in calls)


def test_pileup_filters(sim_lib, oracle):
    """mapq / base-quality thresholds, CIGARs with =/X/N/H/P ops and reads hanging over the contig end"""
    from bind import oracle_pileup
    from rsicnv_b200 import api
    Ls = 300_000
    fa = synth.make_fasta(Ls, 8)
    reads, _ = synth.make_reads(Ls, 8, fa, coverage=8, n_events=0, frac_mapq0=0.3, frac_lowq=0.5)
    rng = np.random.default_rng(5)
    # rewrite some CIGARs to exotic but legal shapes of the same query length
    cig = reads["cigar"].copy(); co = reads["cigar_off"]
    one = np.flatnonzero(np.diff(co.astype(np.int64)) == 3)
    for r in one[:: 3]:
        k = co[r]
        cig[k] = (50 << 4) | 7; cig[k + 1] = (3 << 4) | 3; cig[k + 2] = (50 << 4) | 0      # 50= 3N 50M
    for r in one[1:: 3]:
        k = co[r]
        cig[k] = (5 << 4) | 5; cig[k + 1] = (60 << 4) | 0; cig[k + 2] = (40 << 4) | 8      # 5H 60M 40X
    reads["cigar"] = cig
    # reads that run past the end of the contig, and one at position 0 (dropped by the reference)
    reads["pos"][-50:] = np.sort(rng.integers(Ls - 80, Ls - 1, 50)).astype(np.int32)
    reads["pos"][0] = 0
    for minq, Q in ((0, 13), (20, 0), (1, 31)):
        oracle.set_params(minq=minq, min_baseQ=Q)
        want = oracle_pileup(oracle, reads, Ls)
        with api.Context(lib=sim_lib, minq=minq, min_baseQ=Q) as ctx:
            ctx.set_reference(fa); ctx.pileup_begin(); ctx.pileup_push(reads); ctx.pileup_end()
            assert np.array_equal(ctx.array(api.ARR_RAW_DEPTH), want), (minq, Q)


def test_isize_sample_with_gaps_and_foreign_mates(sim_lib, oracle):
    """bam_rd_pr_stats' reset rule (a gap of more than 1 kb restarts the sample), mates on other contigs, reads
    without CIGAR -- the order of the tests is SURVEY.md A.2"""
    fa = synth.make_fasta(L, 4)
    reads, _ = synth.make_reads(L, 4, fa, coverage=10, n_events=4, lens=(3000, 8000, 20000))
    pos = reads["pos"]
    drop = ((pos >= 10_100_000) & (pos < 10_103_000)) | ((pos >= 10_300_000) & (pos < 10_301_500)) | ((pos >= 10_450_000) & (pos < 10_450_900))
    keep = ~drop
    co = reads["cigar_off"].astype(np.int64); qo = reads["qual_off"].astype(np.int64)
    nc = np.diff(co)[keep]; nq = np.diff(qo)[keep]
    ckeep = np.repeat(keep, np.diff(co)); qkeep = np.repeat(keep, np.diff(qo))
    r2 = {k: reads[k][keep] for k in ("pos", "mpos", "isize", "mtid", "flag", "mapq")}
    r2["cigar"] = reads["cigar"][ckeep]; r2["qual"] = reads["qual"][qkeep]
    r2["cigar_off"] = np.concatenate(([0], np.cumsum(nc))).astype(np.uint32); r2["qual_off"] = np.concatenate(([0], np.cumsum(nq))).astype(np.uint64)
    rng = np.random.default_rng(9)
    n = len(r2["pos"])
    r2["mtid"] = r2["mtid"].copy(); r2["mtid"][rng.random(n) < 0.01] = 3      # mate elsewhere: skipped
    r2["mtid"][rng.random(n) < 0.005] = -1                                      # mate unmapped: kept, not proper
    calls, st = run_bam_case(sim_lib, oracle, fa, r2, n_batches=2, min_baseQ=10, minq=0)
    assert len(calls) >= 1 and st.isize_mean > 0


def test_pileup_quality_thresholds_full_byte_range(sim_lib, oracle):
    """qualities over the whole 0..255 byte range against thresholds on both sides of 128 (the SWAR byte test)"""
    from bind import oracle_pileup
    from rsicnv_b200 import api
    Ls = 120_000
    fa = synth.make_fasta(Ls, 9)
    reads, _ = synth.make_reads(Ls, 9, fa, coverage=6, n_events=0)
    rng = np.random.default_rng(11)
    reads["qual"] = rng.integers(0, 256, len(reads["qual"]), dtype=np.uint8)
    for Q in (0, 1, 93, 127, 128, 129, 200, 255, 300, -5):
        oracle.set_params(minq=0, min_baseQ=Q)
        want = oracle_pileup(oracle, reads, Ls)
        with api.Context(lib=sim_lib, minq=0, min_baseQ=Q) as ctx:
            ctx.set_reference(fa); ctx.pileup_begin(); ctx.pileup_push(reads); ctx.pileup_end()
            assert np.array_equal(ctx.array(api.ARR_RAW_DEPTH), want), Q


def test_stat_calls_matches_oracle(case, sim_lib, oracle):
    """`rsicnv stat` (rsi.cpp:2235-2249): RP / Q0 for calls that come from a file -- any coordinates, several contigs in one list
    (entries of other contigs are left alone, but the search distance DIS is carried across them, pairrd.cpp:655-656)"""
    from bind import new_cnv, oracle_cnv_stat
    from rsicnv_b200 import api
    fa, reads, ev = case
    lst = []
    for k, (s, e, f) in enumerate(ev):
        lst.append(new_cnv(start=s, end=e, type=0 if f < 1 else 1, tid=0))
        lst.append(new_cnv(start=s + 700 * k, end=e + 2100, type=1 if f < 1 else 0, tid=0))      # shifted, wrong type
    lst.insert(1, new_cnv(start=5_000_000, end=5_004_000, type=0, tid=1))                           # another contig: DIS grows to 4001 for what follows
    lst.append(new_cnv(start=9_000_000, end=8_990_000, type=2, tid=0))                              # reversed, unknown type
    oracle.set_params(minq=0, min_baseQ=10)
    mine = [c for c in lst if c.tid == 0]
    # the oracle works on one contig's list: hand it the same DIS history by keeping the foreign entry's length in the list
    want_all = oracle_cnv_stat(oracle, reads, L, lst)
    with api.Context(lib=sim_lib, minq=0, min_baseQ=10) as ctx:
        ctx.reads_begin(0, L)
        ctx.pileup_push(reads)
        got = ctx.stat_calls(lst)
    for i, (g, w, src) in enumerate(zip(got, want_all, lst)):
        if src.tid == 0:
            assert g.rp == w.rp and (g.q0 == w.q0 or abs(g.q0 - w.q0) <= 1e-9 * abs(w.q0)), (i, g.rp, w.rp, g.q0, w.q0)
        else:
            assert g.rp == -1 and g.q0 == -1.0
    assert any(g.rp > 0 for g in got) and len(mine) == len(lst) - 1
