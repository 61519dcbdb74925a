"""BGZF inflate + BAM record decoding on the GPU (k_bam.cuh) through the C ABI: same cases as the emulator suite, plus a
12 Mbp / 30x file and the round trip file -> decode -> pileup -> calls against the host-decoded path."""
import zlib

import numpy as np
import pytest

import test_sim_bam_decode as T
from rsicnv_b200 import api, synth

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=[1, 2], ids=["lane-per-block", "warp-per-block"])
def inflate_mode(request, monkeypatch):
    """every test of this module runs with each of the two inflate kernels (rsigpu_set_inflate_mode through api.Context's env hook)"""
    monkeypatch.setenv("RSIGPU_INFLATE_MODE", str(request.param))


@pytest.mark.parametrize("level,strategy,chunk", [(1, 0, None), (6, 0, 70000), (0, 0, 150000), (9, zlib.Z_FIXED, 200000), (1, zlib.Z_HUFFMAN_ONLY, 90001)])
def test_decode_matches_source_reads(gpu_lib, tmp_path, level, strategy, chunk):
    T.test_decode_matches_source_reads(gpu_lib, tmp_path, level, strategy, chunk)


def test_small_blocks_and_long_header(gpu_lib, tmp_path):
    T.test_small_blocks_and_long_header(gpu_lib, tmp_path)


def test_wrong_guess_is_repaired(gpu_lib, tmp_path):
    T.test_wrong_guess_is_repaired(gpu_lib, tmp_path)


def test_corrupt_and_truncated_input(gpu_lib, tmp_path):
    T.test_corrupt_and_truncated_input(gpu_lib, tmp_path)


def test_decoded_reads_give_the_same_calls(gpu_lib, tmp_path):
    T.test_decoded_reads_give_the_same_calls(gpu_lib, tmp_path)


def test_real_looking_records(gpu_lib, tmp_path):
    T.test_real_looking_records(gpu_lib, tmp_path)


def test_large_file_all_fields(gpu_lib, tmp_path):
    """12 Mbp at 30x (3.6 M records, ~13 k BGZF blocks), fed in 32 MiB pieces"""
    L = 12_000_000
    fa = synth.make_fasta(L, 19)
    reads, _ = synth.make_reads(L, 19, fa, coverage=30, n_events=6)
    path = str(tmp_path / "t.bam")
    synth.write_bam(path, [("19", L), ("20", 1000)], {0: reads}, level=1, random_seq=3)
    got, h = T.decode_file(gpu_lib, path, 32 << 20)
    assert sorted(got) == [0]
    T.assert_same_reads(got[0], reads)


def test_feed_parts_equals_separate_feeds(gpu_lib, tmp_path):
    T.test_feed_parts_equals_separate_feeds(gpu_lib, tmp_path)
