"""BAM-input path on the real library: pileup kernel, insert-size sample and RP / Q0 counts through the C ABI
against the oracle on the same read batches, the committed reference table, and properties at chr19 scale."""
import hashlib
import json
import os

import numpy as np
import pytest

from common import run_bam_case
from rsicnv_b200 import api, synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_bam_path_matches_oracle_and_reference_table(gpu_lib, oracle):
    g = json.load(open(os.path.join(HERE, "golden", "reference_vectors.json")))["bam_case"]
    fa = synth.make_fasta(g["L"], g["seed"])
    reads, _ = synth.make_reads(g["L"], g["seed"], fa, coverage=g["coverage"], n_events=g["n_events"], lens=tuple(g["lens"]))
    calls, st = run_bam_case(gpu_lib, oracle, fa, reads, n_batches=4, minq=g["minq"], min_baseQ=g["min_baseQ"])
    lib = api.load_library(gpu_lib)
    rows = [api.format_row(lib, c, "1", st.rdmedian, st.rdsd) for c in calls]
    assert rows == [ln for ln in g["table"] if not ln.startswith("#")]          # byte-identical to the reference CLI's table
    assert api.format_row(lib, None, "", 0, 0) == [ln for ln in g["table"] if ln.startswith("#CHROM")][0]


@pytest.mark.parametrize("cov,seed", [(30, 11), (60, 12)])
def test_bam_path_30x_60x(cov, seed, gpu_lib, oracle):
    L = 12_000_000
    fa = synth.make_fasta(L, seed)
    reads, _ = synth.make_reads(L, seed, fa, coverage=cov, n_events=10, lens=(2000, 5000, 10000, 30000))
    calls, st = run_bam_case(gpu_lib, oracle, fa, reads, n_batches=5, minq=0, min_baseQ=10)
    assert len(calls) >= 5 and sum(c.rp for c in calls) > 0


def test_pileup_filters_gpu(gpu_lib, oracle):
    from bind import oracle_pileup
    Ls = 2_000_000
    fa = synth.make_fasta(Ls, 8)
    reads, _ = synth.make_reads(Ls, 8, fa, coverage=40, n_events=0, frac_mapq0=0.3, frac_lowq=0.5, frac_indel=0.2)
    cig = reads["cigar"].copy(); co = reads["cigar_off"]
    three = np.flatnonzero(np.diff(co.astype(np.int64)) == 3)
    k = co[three[::3]]; cig[k] = (50 << 4) | 7; cig[k + 1] = (3 << 4) | 3; cig[k + 2] = (50 << 4) | 0
    k = co[three[1::3]]; cig[k] = (5 << 4) | 5; cig[k + 1] = (60 << 4) | 0; cig[k + 2] = (40 << 4) | 8
    reads["cigar"] = cig
    rng = np.random.default_rng(5)
    reads["pos"][-500:] = np.sort(rng.integers(Ls - 80, Ls - 1, 500)).astype(np.int32)
    reads["pos"][0] = 0
    for minq, Q in ((0, 13), (20, 0), (1, 31)):
        oracle.set_params(minq=minq, min_baseQ=Q)
        want = oracle_pileup(oracle, reads, Ls)
        with api.Context(lib=gpu_lib, minq=minq, min_baseQ=Q) as ctx:
            ctx.set_reference(fa); ctx.pileup_begin(); ctx.pileup_push(reads); ctx.pileup_end()
            assert np.array_equal(ctx.array(api.ARR_RAW_DEPTH), want), (minq, Q)


def test_unsorted_batch_is_rejected(gpu_lib):
    Ls = 300_000
    fa = synth.make_fasta(Ls, 8)
    reads, _ = synth.make_reads(Ls, 8, fa, coverage=5, n_events=0)
    reads["pos"] = reads["pos"][::-1].copy()
    with api.Context(lib=gpu_lib) as ctx:
        ctx.set_reference(fa); ctx.pileup_begin(); ctx.pileup_push(reads)
        with pytest.raises(api.RsiGpuError):
            ctx.pileup_end()


def test_chr19_pileup_full_size_checksum(gpu_lib):
    """59 Mbp, 30x (17.7 M reads): sum of depths == number of counted bases, computed independently with numpy"""
    L = synth.CHR19_LEN
    fa = synth.make_fasta(L, 19)
    reads, _ = synth.make_reads(L, 19, fa, coverage=30, n_events=20)
    with api.Context(lib=gpu_lib, minq=0, min_baseQ=10) as ctx:
        ctx.set_reference(fa); ctx.pileup_begin(); ctx.pileup_push(reads); ctx.pileup_end()
        raw = ctx.array(api.ARR_RAW_DEPTH)
    keep = (reads["pos"] != 0) & ((reads["flag"] & (256 | 1024)) == 0)
    # counted bases per read: M-op bases with quality >= 10 (synthetic reads use M/I/D/S only and stay inside the contig)
    cop = reads["cigar"] & 15; cl = (reads["cigar"] >> 4).astype(np.int64)
    qadv = np.where((cop == 0) | (cop == 1) | (cop == 4), cl, 0)
    nread = len(reads["pos"])
    rid = np.repeat(np.arange(nread), np.diff(reads["cigar_off"].astype(np.int64)))
    first = reads["cigar_off"][:-1].astype(np.int64)
    qstart = np.cumsum(qadv) - qadv; qstart -= np.repeat(qstart[first], np.diff(reads["cigar_off"].astype(np.int64)))
    good = np.concatenate(([0], np.cumsum(reads["qual"] >= 10, dtype=np.int64)))
    qbase = reads["qual_off"][:-1].astype(np.int64)[rid] + qstart
    per_op = np.where(cop == 0, good[qbase + cl] - good[qbase], 0)
    total = int(per_op[keep[rid]].sum())
    assert int(raw.astype(np.int64).sum()) == total
    assert raw.min() >= 0 and raw[:60_000].max() == 0
