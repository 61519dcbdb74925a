"""BASELINE.json's configurations at their FULL sizes, parity-gated:

  config 1  chr19-shaped depth file (59,128,983 bp) through the CLI        vs the unmodified reference CLI, byte-identical table
  config 2  chr19-shaped 30x BAM (17.7 M reads) through the CLI            vs the unmodified reference CLI, identical table
  config 4  chr1-sized depth (249,250,621 bp) through the C ABI            vs the oracle restatement, sweep -m 51/101/501, -MED, -NOGC, -cap -1

The reference / oracle side of each case costs 20-60 s of one host core; set RSI_SKIP_FULL_SIZE=1 to leave them out of a quick run.
"""
import os
import subprocess

import numpy as np
import pytest

from bind import REF_BAMTOOL, REF_BIN, have_ref
from common import assert_calls_equal, oracle_params
from rsicnv_b200 import api, synth

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(os.environ.get("RSI_SKIP_FULL_SIZE") == "1", reason="RSI_SKIP_FULL_SIZE=1")]

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "rsicnv_b200", "bin", "rsicnv")
CHR1 = synth.B37_LENS["1"]


def _table(path):
    return [ln for ln in open(path).read().splitlines() if not ln.startswith("#input")]


@pytest.fixture(scope="module")
def cli():
    if not os.path.exists(CLI):
        subprocess.run(["make", "-s", "cli"], cwd=ROOT, check=True)
    return CLI


def test_config1_chr19_depth_file_cli_matches_reference(cli, tmp_path):
    """rsicnv rsi -d <chr19 depth file> -c 19 -f <FASTA> -m 101 -np (BASELINE.json configs[0]) at 59,128,983 bp"""
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    L = synth.CHR19_LEN
    fa = synth.make_fasta(L, 19)
    d, ev = synth.make_depth(L, 19, fa, n_events=20)
    fasta = str(tmp_path / "c.fa"); rd = str(tmp_path / "c.rd")
    synth.write_fasta(fasta, "19", fa)
    synth.write_depth_file_fast(rd, d)
    common = ["rsi", "-d", rd, "-c", "19", "-f", fasta, "-m", "101", "-np"]
    subprocess.run([REF_BIN] + common + ["-o", str(tmp_path / "ref.txt")], check=True, capture_output=True)
    out = subprocess.run([cli] + common + ["-o", str(tmp_path / "ours.txt")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    assert open(str(tmp_path / "ours.txt")).read() == open(str(tmp_path / "ref.txt")).read()
    assert len(_table(str(tmp_path / "ref.txt"))) >= 12      # ~19 of the 20 planted events (the last marked run is never reported)
    from test_cli import log_lines
    assert log_lines(str(tmp_path / "ours.txt.log"), str(tmp_path / "ours.txt")) == log_lines(str(tmp_path / "ref.txt.log"), str(tmp_path / "ref.txt"))


def test_config2_chr19_bam_cli_matches_reference(cli, tmp_path):
    """rsicnv rsi -b <chr19 30x BAM> -f <FASTA> -q 0 -Q 10 -np (BASELINE.json configs[1]): full pileup + RP/Q0 at 59,128,983 bp"""
    if not have_ref():
        pytest.skip("oracle/_ref not built")
    L = synth.CHR19_LEN
    fa = synth.make_fasta(L, 19)
    reads, ev = synth.make_reads(L, 19, fa, coverage=30.0, n_events=20)
    bam = str(tmp_path / "t.bam"); fasta = str(tmp_path / "t.fa")
    synth.write_bam(bam, [("19", L)], {0: reads}, level=1, random_seq=7, threads=max(4, min(32, os.cpu_count() or 8)))
    del reads
    synth.write_fasta(fasta, "19", fa)
    subprocess.run([REF_BAMTOOL, "index", bam], check=True)
    common = ["rsi", "-b", bam, "-f", fasta, "-q", "0", "-Q", "10", "-np"]
    subprocess.run([REF_BIN] + common + ["-o", str(tmp_path / "ref.txt")], check=True, capture_output=True)
    out = subprocess.run([cli] + common + ["-o", str(tmp_path / "ours.txt")], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    ours, ref = _table(str(tmp_path / "ours.txt")), _table(str(tmp_path / "ref.txt"))
    assert ours == ref, out.stderr
    assert len(ref) >= 12 and any("RP=" in ln and "RP=0;" not in ln and "RP=-1" not in ln for ln in ref)


@pytest.fixture(scope="module")
def chr1_case():
    fa = synth.make_fasta(CHR1, 1)
    d, ev = synth.make_depth(CHR1, 1, fa, n_events=40, lens=(2000, 5000, 10000, 30000, 100000))
    return fa, d, ev


# config 4: -m 51/101/501 x -NB/-MED x -NOGC x -cap 4/-1, each factor at least once, at 249,250,621 bp (rsi.cpp:2018-2030; Lmax = 196 at m = 51)
CHR1_SWEEP = [dict(m=51), dict(m=101, trans="MED", gcadjust=False), dict(m=501, cap=-1.0), dict(m=101)]


@pytest.mark.parametrize("kw", CHR1_SWEEP, ids=lambda kw: "-".join(f"{k}{v}" for k, v in kw.items()))
def test_config4_chr1_sweep_matches_oracle(kw, chr1_case, gpu_lib, oracle):
    fa, d, ev = chr1_case
    oracle.set_params(**oracle_params(kw))
    want = oracle.depth_path(d, fa, 3, want_bins=True)
    with api.Context(lib=gpu_lib, **kw) as ctx:
        ctx.set_reference(fa); ctx.set_depth(d)
        calls = ctx.run()
        st = ctx.chr_stats()
        assert st.target_len == CHR1 and st.nbins == st.compact_len // ctx.m
        assert st.rdmedian == want["stats"][0] and st.rdsd == want["stats"][1]
        nb = st.nbins
        bm, bn, bi, bs = want["bins"]
        assert np.array_equal(ctx.array(api.ARR_BIN_MED), bm[:nb]), "median_transfer"
        assert np.array_equal(ctx.array(api.ARR_BIN_NBN), bn[:nb]), "negative_binomial_transfer"
        assert np.array_equal(ctx.array(api.ARR_BIN_STATUS), bs[:nb]), "RSI status"
        assert_calls_equal(calls, want["calls"], f"chr1 {kw}")
        for k, name in enumerate(("segments", "blocks", "premerge", "merged")):
            assert_calls_equal(ctx.array(api.ARR_SEGMENTS + k), oracle.last_list(k), f"{name} list")
        assert len(calls) >= 20
