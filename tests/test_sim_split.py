"""One contig split over several part contexts (rsigpu_split_run) under the emulator: per-base work on each part's own base
range, the parts' integer tables added on the lead, halo of the bins that straddle a cut -- the result must be bit-identical to
the unsplit run (depth, bins, statistics, calls incl. RP / Q0).  On a GPU box the parts sit on different devices and the copies
cross NVLink (tests/test_gpu_multi.py)."""
import numpy as np
import pytest

from common import make_case
from rsicnv_b200 import api, synth


def _snapshot(ctx, calls):
    st = ctx.chr_stats()
    return dict(calls=[bytes(c) for c in calls], depth=ctx.array(api.ARR_DEPTH).tobytes(), nbn=ctx.array(api.ARR_BIN_NBN).tobytes(),
                med=ctx.array(api.ARR_BIN_MED).tobytes(), status=ctx.array(api.ARR_BIN_STATUS).tobytes(),
                stats=(st.rdmedian, st.rdsd, st.rdmad, st.isize_mean, st.isize_sd, st.nbins, st.compact_len))


@pytest.mark.parametrize("nparts,kw", [(3, {}), (4, dict(m=51, cap=-1.0)), (2, dict(m=501, trans="MED")), (3, dict(gcadjust=False))],
                         ids=lambda v: str(v).replace(" ", ""))
def test_split_depth_input_identical(nparts, kw, sim_lib):
    L = 1_000_003 if nparts < 4 else 1_200_007
    fa, d, _ = make_case(L, 7, stress=True)
    with api.Context(lib=sim_lib, **kw) as c0:
        c0.set_reference(fa); c0.set_depth(d)
        want = _snapshot(c0, c0.run())
    parts = [api.Context(lib=sim_lib, **kw) for _ in range(nparts)]
    try:
        for p in parts:
            p.set_reference(fa); p.set_depth(d)
        got = _snapshot(parts[0], api.split_run(parts))
    finally:
        for p in parts:
            p.close()
    assert got == want and len(want["calls"]) > 3


def test_split_bam_input_identical(sim_lib, tmp_path):
    """reads decoded from a BAM image and dealt to three parts by position range (rsigpu_bam_take_range) incl. the read halo"""
    L = 10_300_000
    fa = synth.make_fasta(L, 7)
    reads, _ = synth.make_reads(L, 7, fa, coverage=8, n_events=5, lens=(3000, 8000, 20000))
    bam = str(tmp_path / "t.bam")
    synth.write_bam(bam, [("1", L)], {0: reads}, level=1, random_seq=3)
    data = np.fromfile(bam, np.uint8)
    h = api.parse_bam_header(data)
    kw = dict(minq=0, min_baseQ=10)
    with api.Context(lib=sim_lib, **kw) as c0:
        c0.set_reference(fa); c0.pileup_begin(); c0.pileup_push(reads); c0.have_reads()
        want = _snapshot(c0, c0.run())
    parts = [api.Context(lib=sim_lib, **kw) for _ in range(3)]
    dec = api.Context(lib=sim_lib)
    try:
        rng = [api.split_range(dec.lib, L, 3, g) for g in range(3)]
        assert rng[0][0] == 0 and rng[2][1] == L and all(rng[g][1] == rng[g + 1][0] for g in range(2))
        for p in parts:
            p.set_reference(fa); p.pileup_begin()
        dec.bam_begin(1)
        consumed, runs = dec.bam_feed(data[h["coff"]:], skip=h["skip"])
        for i, (tid, n) in enumerate(runs):
            for p, (b, e, halo) in zip(parts, rng):
                dec.bam_take_range(i, p, b - halo, e)
        dec.bam_end()
        for p in parts:
            p.have_reads()
        got = _snapshot(parts[0], api.split_run(parts))
    finally:
        dec.close()
        for p in parts:
            p.close()
    assert got == want and len(want["calls"]) >= 2


def test_split_argument_errors(sim_lib):
    fa, d, _ = make_case(300_000, 3)
    with api.Context(lib=sim_lib) as a, api.Context(lib=sim_lib) as b, api.Context(lib=sim_lib) as c3:
        for p in (a, b, c3):
            p.set_reference(fa); p.set_depth(d)
        with pytest.raises(api.RsiGpuError):      # 300 kb over three parts: shorter than two cut units each
            api.split_run([a, b, c3])
        with pytest.raises(api.RsiGpuError):      # the same context twice
            api.split_run([a, a])
        assert len(api.split_run([a])) == len(api.split_run([a, b]))
