"""Both inflate kernels (one lane / one warp per BGZF block) against zlib on DAMAGED input, in the CPU emulator (out-of-bounds
accesses of a kernel are real segfaults there).  For every file with one or two flipped bits:
  * neither kernel may accept a payload that zlib's raw inflate (what bgzf.c:277-313 calls) rejects or that decodes to another
    size than the block's ISIZE -- and neither may reject, as a deflate error, one that zlib accepts;
  * the two kernels must agree: both refuse, or both decode the same records.
Like the reference's reader, the GPU path does not check CRC32, so damage that leaves the deflate structure valid is only caught
when it breaks a BAM record."""
import hashlib
import struct
import zlib

import numpy as np
import pytest

from rsicnv_b200 import api, synth


def zlib_verdict(b):
    """raw inflate of every BGZF block: True = all fine, False = some block fails, None = a BGZF header itself is damaged"""
    b = bytes(b); off = 0
    while off + 18 <= len(b):
        if b[off:off + 4] != b"\x1f\x8b\x08\x04" or b[off + 12:off + 14] != b"BC":
            return None
        xlen = struct.unpack_from("<H", b, off + 10)[0]
        bsize = struct.unpack_from("<H", b, off + 16)[0] + 1
        if off + bsize > len(b) or bsize < 12 + xlen + 8:
            return None
        isize = struct.unpack_from("<I", b, off + bsize - 4)[0]
        if isize > 65536:
            return None
        try:
            d = zlib.decompressobj(-15)
            out = d.decompress(b[off + 12 + xlen:off + bsize - 8])
            if not d.eof or len(out) != isize:
                return False
        except zlib.error:
            return False
        off += bsize
    return True


def decode(lib, mode, body, skip):
    ctx = api.Context(lib=lib); ctx.set_inflate_mode(mode); ctx.bam_begin(1)
    try:
        consumed, runs = ctx.bam_feed(body, skip=skip)
        sig = [consumed, runs]
        for i in range(len(runs)):
            r = ctx.bam_run_reads(i)
            sig.append(hashlib.sha1(b"".join(np.ascontiguousarray(r[k]).tobytes() for k in sorted(r))).hexdigest())
        return True, sig
    except api.RsiGpuError as e:
        return False, str(e)
    finally:
        ctx.close()


@pytest.mark.parametrize("level", [1, 6, 9])
def test_bit_flips_against_zlib(sim_lib, tmp_path, level):
    fa = synth.make_fasta(30000, 77)
    reads, _ = synth.make_reads(30000, 77, fa, coverage=4, n_events=2)
    path = str(tmp_path / "t.bam")
    synth.write_bam(path, [("1", 30000)], {0: reads}, level=level, block_size=4000)
    data = np.fromfile(path, np.uint8)
    h = api.parse_bam_header(data)
    body = data[h["coff"]:]
    rng = np.random.default_rng(100 + level)
    checked = accepted = 0
    for it in range(14):
        b = body.copy()
        for _ in range(int(rng.integers(1, 3))):
            b[int(rng.integers(18, len(b) - 8))] ^= 1 << int(rng.integers(0, 8))
        z = zlib_verdict(b)
        if z is None:
            continue
        lane_ok, lane = decode(sim_lib, 1, b, h["skip"])
        warp_ok, warp = decode(sim_lib, 2, b, h["skip"])
        assert lane_ok == warp_ok and (not lane_ok or lane == warp), (it, lane, warp)
        if not z:
            assert not lane_ok, (it, "accepted a payload zlib rejects")
        elif not lane_ok:
            assert "deflate" not in lane and "deflate" not in warp, (it, lane, warp)     # refused for a broken RECORD, not for the deflate stream
        checked += 1; accepted += lane_ok
    assert checked >= 10
