"""BGZF inflate + BAM record decoding kernels (k_bam.cuh) in the CPU emulator: the decoded structure-of-arrays must equal
the reads the file was written from, for every deflate block type, chunking and record/block alignment."""
import os
import zlib

import numpy as np
import pytest

from rsicnv_b200 import api, synth

FIELDS = ("pos", "mpos", "isize", "mtid", "flag", "mapq", "cigar_off", "cigar", "qual_off", "qual")

@pytest.fixture(autouse=True, params=[1, 2], ids=["lane-per-block", "warp-per-block"])
def inflate_mode(request, monkeypatch):
    """every test of this module runs with each of the two inflate kernels (rsigpu_set_inflate_mode through api.Context's env hook)"""
    monkeypatch.setenv("RSIGPU_INFLATE_MODE", str(request.param))


def concat_reads(parts):
    out = {k: [] for k in FIELDS}
    co = 0; qo = 0
    for p in parts:
        for k in ("pos", "mpos", "isize", "mtid", "flag", "mapq", "cigar", "qual"):
            out[k].append(p[k])
        out["cigar_off"].append(p["cigar_off"][:-1].astype(np.uint64) + co); out["qual_off"].append(p["qual_off"][:-1].astype(np.uint64) + qo)
        co += int(p["cigar_off"][-1]); qo += int(p["qual_off"][-1])
    res = {k: (np.concatenate(v) if v else np.zeros(0)) for k, v in out.items()}
    res["cigar_off"] = np.concatenate((res["cigar_off"], [co])).astype(np.uint32)
    res["qual_off"] = np.concatenate((res["qual_off"], [qo])).astype(np.uint64)
    return res


def decode_file(lib, path, chunk=None):
    """feed the file in `chunk`-byte pieces (None: all at once); returns {tid: read dict}, decoder statistics"""
    data = np.fromfile(path, np.uint8)
    h = api.parse_bam_header(data)
    ctx = api.Context(lib=lib)
    ctx.bam_begin(len(h["names"]))
    parts = {}
    off = h["coff"]; skip = h["skip"]; pending = np.zeros(0, np.uint8)
    step = chunk or len(data)
    first = True
    while off < len(data) or len(pending):
        piece = data[off:off + step]; off += len(piece)
        buf = np.concatenate((pending, piece))
        consumed, runs = ctx.bam_feed(buf, skip=skip if first else 0)
        if consumed:
            first = False
        for i, (tid, n) in enumerate(runs):
            r = ctx.bam_run_reads(i)
            assert len(r["pos"]) == n
            parts.setdefault(tid, []).append(r)
        pending = buf[consumed:]
        if off >= len(data) and consumed == 0:
            break
    assert len(pending) == 0
    ctx.bam_end()
    ctx.close()
    return {tid: concat_reads(p) for tid, p in parts.items()}, h


def assert_same_reads(got, exp):
    for k in FIELDS:
        e = np.asarray(exp[k])
        assert got[k].shape == e.shape, k
        assert np.array_equal(got[k].astype(np.int64), e.astype(np.int64)), k


def make_reads(L, seed, cov=6):
    fa = synth.make_fasta(L, seed)
    reads, _ = synth.make_reads(L, seed, fa, coverage=cov, n_events=2)
    return fa, reads


@pytest.mark.parametrize("level,strategy,chunk", [(1, 0, None), (6, 0, 70000), (0, 0, 150000), (9, zlib.Z_FIXED, 200000), (1, zlib.Z_HUFFMAN_ONLY, 90001)])
def test_decode_matches_source_reads(sim_lib, tmp_path, level, strategy, chunk):
    contigs = [("1", 60000), ("2", 5000), ("3", 90000), ("MT", 20000)]
    rb = {0: make_reads(60000, 11)[1], 2: make_reads(90000, 12)[1], 3: make_reads(20000, 13, cov=3)[1]}
    path = str(tmp_path / "t.bam")
    synth.write_bam(path, contigs, rb, level=level, strategy=strategy, random_seq=5)
    got, h = decode_file(sim_lib, path, chunk)
    assert h["names"] == ["1", "2", "3", "MT"] and h["lens"] == [60000, 5000, 90000, 20000]
    assert sorted(got) == [0, 2, 3]
    for tid in got:
        assert_same_reads(got[tid], rb[tid])


def test_small_blocks_and_long_header(sim_lib, tmp_path):
    """BGZF blocks far smaller than a record spacing pattern (blocks without any record start) and a header spanning several blocks"""
    contigs = [("c%d" % i, 40000) for i in range(300)]
    rb = {7: make_reads(40000, 21)[1], 299: make_reads(40000, 22, cov=2)[1]}
    path = str(tmp_path / "t.bam")
    synth.write_bam(path, contigs, rb, level=6, block_size=97)
    got, h = decode_file(sim_lib, path, 50000)
    assert len(h["names"]) == 300
    for tid in (7, 299):
        assert_same_reads(got[tid], rb[tid])


def test_wrong_guess_is_repaired(sim_lib, tmp_path):
    """quality bytes that look like a record header right at a BGZF block boundary: the speculative start is wrong, the proof
    step must notice and re-walk that block"""
    import struct
    L = 80000
    fa, reads = make_reads(L, 31, cov=8)
    contigs = [("19", L)]
    # stream offset of each record: header size, then records
    text = "@HD\tVN:1.0\tSO:coordinate\n@SQ\tSN:19\tLN:%d\n" % L
    hdr = 12 + len(text) + 4 + len("19") + 1 + 4
    ncig = np.diff(reads["cigar_off"].astype(np.int64)); lq = np.diff(reads["qual_off"].astype(np.int64))
    size = 4 + 32 + 2 + 4 * ncig + (lq + 1) // 2 + lq
    start = hdr + np.concatenate(([0], np.cumsum(size)))[:-1]
    qstart = start + 36 + 2 + 4 * ncig + (lq + 1) // 2
    BS = 65280
    planted = 0
    qual = reads["qual"].copy()
    for k in range(1, int((start[-1] + size[-1]) // BS) + 1):
        b = k * BS
        r = int(np.searchsorted(qstart, b, side="right") - 1)
        if r < 0 or r + 1 >= len(start) or not (qstart[r] <= b and b + 37 <= qstart[r] + lq[r]):
            continue
        # a 37-byte "record" (no name, CIGAR or bases) whose block_size leads exactly to the next true record: the walk from it
        # is clean, so only the proof step can tell that it is not a record
        bs = int(start[r + 1]) - b - 4
        fake = struct.pack("<IiiBBHHHiiii", bs, 0, 5, 1, 0, 0, 0, 0, 0, 0, 5, 0) + b"\x00"
        o = int(reads["qual_off"][r]) + (b - int(qstart[r]))
        qual[o:o + len(fake)] = np.frombuffer(fake, np.uint8)
        planted += 1
    assert planted >= 1
    reads = dict(reads); reads["qual"] = qual
    path = str(tmp_path / "t.bam")
    synth.write_bam(path, contigs, {0: reads}, level=1)
    data = np.fromfile(path, np.uint8)
    h = api.parse_bam_header(data)
    ctx = api.Context(lib=sim_lib)
    ctx.bam_begin(1)
    consumed, runs = ctx.bam_feed(data[h["coff"]:], skip=h["skip"])
    assert runs == [(0, len(reads["pos"]))]
    assert_same_reads(ctx.bam_run_reads(0), reads)
    assert ctx.debug_state()["bam_rewalked"] >= planted
    ctx.bam_end(); ctx.close()


def test_corrupt_and_truncated_input(sim_lib, tmp_path):
    fa, reads = make_reads(30000, 41, cov=40)
    path = str(tmp_path / "t.bam")
    synth.write_bam(path, [("19", 30000)], {0: reads}, level=6, random_seq=1)
    data = np.fromfile(path, np.uint8)
    h = api.parse_bam_header(data)
    ctx = api.Context(lib=sim_lib)
    # an invalid deflate block type at the start of a payload (like the reference's bgzf.c, the GPU path does not check CRC32:
    # only structural damage is detectable)
    bad = data.copy(); bad[h["coff"] + 18] = 0x07
    ctx.bam_begin(1)
    with pytest.raises(api.RsiGpuError):
        ctx.bam_feed(bad[h["coff"]:], skip=h["skip"])
    # not a block boundary
    ctx.bam_begin(1)
    with pytest.raises(api.RsiGpuError):
        ctx.bam_feed(data[h["coff"] + 1:], skip=h["skip"])
    # file cut inside a record: the first blocks decode, end() reports the cut
    ctx.bam_begin(1)
    assert len(data) > 150000
    consumed, runs = ctx.bam_feed(data[h["coff"]:h["coff"] + 100000], skip=h["skip"])
    assert 0 < consumed <= 100000 and runs and runs[0][1] < len(reads["pos"])
    with pytest.raises(api.RsiGpuError):
        ctx.bam_end()
    ctx.close()


def test_decoded_reads_give_the_same_calls(sim_lib, tmp_path):
    """file bytes -> GPU decode -> take -> run equals host-decoded reads -> pileup_push -> run"""
    L = 300000
    fa = synth.make_fasta(L, 3)
    reads, _ = synth.make_reads(L, 3, fa, coverage=12, n_events=6, lens=(3000, 8000, 20000))
    path = str(tmp_path / "t.bam")
    synth.write_bam(path, [("19", L)], {0: reads}, level=1)
    data = np.fromfile(path, np.uint8)
    h = api.parse_bam_header(data)
    a = api.Context(lib=sim_lib, minq=0, min_baseQ=10)
    a.set_reference(fa); a.pileup_begin(); a.pileup_push(reads); a.have_reads()
    want = a.run()
    b = api.Context(lib=sim_lib, minq=0, min_baseQ=10)
    b.set_reference(fa); b.pileup_begin()
    b.bam_begin(1)
    off = h["coff"]; pending = np.zeros(0, np.uint8); first = True
    while off < len(data) or len(pending):
        buf = np.concatenate((pending, data[off:off + 120000])); off += 120000
        consumed, runs = b.bam_feed(buf, skip=h["skip"] if first else 0)
        first = first and not consumed
        for i, (tid, n) in enumerate(runs):
            assert tid == 0
            b.bam_take(i, b)
        pending = buf[consumed:]
    b.bam_end(); b.have_reads()
    got = b.run()
    assert len(got) == len(want) and len(want) > 0
    for x, y in zip(got, want):
        assert bytes(x) == bytes(y)
    a.close(); b.close()


def _rebgzf(src, dst, compress_block, bs=60000):
    """re-block a BGZF file: same decoded stream, each block's deflate stream produced by compress_block(bytes)"""
    import struct
    data = open(src, "rb").read()
    dec = b""; off = 0
    while off < len(data):
        bsize = struct.unpack_from("<H", data, off + 16)[0] + 1
        dec += zlib.decompress(data[off + 18:off + bsize - 8], -15)
        off += bsize
    with open(dst, "wb") as f:
        for a in range(0, len(dec), bs):
            blk = dec[a:a + bs]
            comp = compress_block(blk)
            f.write(b"\x1f\x8b\x08\x04\x00\x00\x00\x00\x00\xff\x06\x00BC\x02\x00" + struct.pack("<H", len(comp) + 25) + comp +
                    struct.pack("<II", zlib.crc32(blk) & 0xffffffff, len(blk)))
        f.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))


def test_mixed_block_types_inside_one_stream(sim_lib, tmp_path):
    """deflate streams that switch between dynamic, stored (also empty, from a flush) and fixed blocks inside one BGZF block"""
    fa, reads = make_reads(50000, 51, cov=10)
    src = str(tmp_path / "a.bam"); dst = str(tmp_path / "b.bam")
    synth.write_bam(src, [("19", 50000)], {0: reads}, level=6, random_seq=2)

    def comp(blk):
        out = b""
        co = zlib.compressobj(6, zlib.DEFLATED, -15)
        n = len(blk)
        out += co.compress(blk[:n // 5]) + co.flush(zlib.Z_FULL_FLUSH)       # dynamic block + empty stored block
        out += co.compress(blk[n // 5:n // 3]) + co.flush(zlib.Z_SYNC_FLUSH)
        out += co.compress(blk[n // 3:n // 3 + 7]) + co.flush(zlib.Z_FULL_FLUSH)   # a tiny (fixed-Huffman) block
        out += co.compress(blk[n // 3 + 7:]) + co.flush()
        return out
    _rebgzf(src, dst, comp)
    got, h = decode_file(sim_lib, dst, 100000)
    assert_same_reads(got[0], reads)


def test_real_looking_records(sim_lib, tmp_path):
    """read names of varying length, random bases, optional fields after the qualities, unmapped reads (refID -1) at the end"""
    contigs = [("1", 70000), ("2", 50000)]
    rb = {0: make_reads(70000, 61, cov=8)[1], 1: make_reads(50000, 62, cov=5)[1]}
    path = str(tmp_path / "t.bam")
    synth.write_bam(path, contigs, rb, level=6, rich=9, unmapped_tail=25)
    got, h = decode_file(sim_lib, path, 64000)
    assert sorted(got) == [-1, 0, 1]
    assert len(got[-1]["pos"]) == 25 and np.all(got[-1]["pos"] == -1)
    for tid in (0, 1):
        assert_same_reads(got[tid], rb[tid])


def test_header_ending_exactly_at_a_block_boundary(sim_lib, tmp_path):
    """first alignment record at the very start of a BGZF block (skip = 0, the feed starts at the second block)"""
    contigs = [("19", 50000)]
    text = "@HD\tVN:1.0\tSO:coordinate\n@SQ\tSN:19\tLN:50000\n"
    hdr_len = 12 + len(text) + 4 + len("19") + 1 + 4
    fa, reads = make_reads(50000, 81, cov=4)
    path = str(tmp_path / "t.bam")
    synth.write_bam(path, contigs, {0: reads}, level=6, block_size=hdr_len)
    h = api.parse_bam_header(np.fromfile(path, np.uint8))
    assert h["skip"] == 0 and h["coff"] > 0
    got, _ = decode_file(sim_lib, path, 30000)
    assert_same_reads(got[0], reads)


def test_feed_stops_at_its_decoded_size_limit(sim_lib, tmp_path):
    """one feed may only produce so many decoded bytes (5 GiB in the product; lowered here through the test hook): it must take
    a prefix of whole blocks, report what it consumed, and carry the cut record into the next feed"""
    fa, reads = make_reads(120000, 91, cov=10)
    path = str(tmp_path / "t.bam")
    synth.write_bam(path, [("19", 120000)], {0: reads}, level=6, block_size=20000, random_seq=4)
    data = np.fromfile(path, np.uint8)
    h = api.parse_bam_header(data)
    ctx = api.Context(lib=sim_lib)
    ctx.set_feed_limit(70000)          # at most three 20000-byte blocks per feed
    ctx.bam_begin(1)
    parts = []; off = h["coff"]; first = True; feeds = 0
    while off < len(data):
        consumed, runs = ctx.bam_feed(data[off:], skip=h["skip"] if first else 0)
        assert consumed > 0
        for i, (tid, n) in enumerate(runs):
            parts.append(ctx.bam_run_reads(i))
        first = False; off += consumed; feeds += 1
    ctx.bam_end(); ctx.close()
    assert feeds > 5
    assert_same_reads(concat_reads(parts), reads)


def test_separate_decoder_and_two_destinations(sim_lib, tmp_path):
    """the CLI's arrangement: one decoder context, the runs of each contig appended to that contig's own context; the staged
    reads must give the same per-base depth as the same reads pushed from the host (the full path from decoded reads is
    test_decoded_reads_give_the_same_calls)"""
    L0, L1 = 120000, 60000
    fa0 = synth.make_fasta(L0, 3); fa1 = synth.make_fasta(L1, 4)
    r0, _ = synth.make_reads(L0, 3, fa0, coverage=6, n_events=2, tid=0)
    r1, _ = synth.make_reads(L1, 4, fa1, coverage=5, n_events=0, tid=1)
    path = str(tmp_path / "t.bam")
    synth.write_bam(path, [("1", L0), ("2", L1)], {0: r0, 1: r1}, level=1, rich=2)
    want = []
    for fa, r in ((fa0, r0), (fa1, r1)):
        c = api.Context(lib=sim_lib, minq=0, min_baseQ=10)
        c.set_reference(fa); c.pileup_begin(); c.pileup_push(r); c.pileup_end()
        want.append(c.array(api.ARR_RAW_DEPTH)); c.close()
    assert all(int(w.sum()) > 0 for w in want)
    data = np.fromfile(path, np.uint8)
    h = api.parse_bam_header(data)
    dec = api.Context(lib=sim_lib)
    dst = [api.Context(lib=sim_lib, minq=0, min_baseQ=10) for _ in range(2)]
    for c, fa in zip(dst, (fa0, fa1)):
        c.set_reference(fa); c.pileup_begin()
    dec.bam_begin(2)
    off = h["coff"]; pending = np.zeros(0, np.uint8); first = True
    while off < len(data) or len(pending):
        buf = np.concatenate((pending, data[off:off + 150000])); off += 150000
        consumed, runs = dec.bam_feed(buf, skip=h["skip"] if first else 0)
        first = first and not consumed
        for i, (tid, n) in enumerate(runs):
            dec.bam_take(i, dst[tid])
        pending = buf[consumed:]
        if off >= len(data) and consumed == 0:
            break
    dec.bam_end()
    for c, w in zip(dst, want):
        c.pileup_end()
        assert np.array_equal(c.array(api.ARR_RAW_DEPTH), w)
        c.close()
    dec.close()


def test_feed_parts_equals_separate_feeds(sim_lib, tmp_path):
    """rsigpu_bam_feed_parts: the record blocks of three files (same refID 0 in each: only the part boundary separates them) decoded
    as ONE chunk give, per part, exactly what each file gives alone; a part that ends inside a record is refused"""
    files = []
    for k, (L, seed) in enumerate(((70000, 41), (30000, 42), (50000, 43))):
        fa, reads = make_reads(L, seed, cov=5)
        path = str(tmp_path / f"p{k}.bam")
        ix = synth.write_bam_aligned(path, "19", L, reads, level=1, random_seq=3, threads=1)
        data = np.fromfile(path, np.uint8)
        files.append((data[int(ix["rec_off"]):int(ix["blk_end"][-1])], reads))
    ctx = api.Context(lib=sim_lib)
    ctx.bam_begin(1)
    runs = ctx.bam_feed_parts([f[0] for f in files])
    assert [(t, p) for t, n, p in runs] == [(0, 0), (0, 1), (0, 2)]
    for i, (tid, n, part) in enumerate(runs):
        assert n == len(files[part][1]["pos"])
        assert_same_reads(ctx.bam_run_reads(i), files[part][1])
    ctx.bam_end()
    # one part alone goes through the same entry point
    ctx.bam_begin(1)
    runs = ctx.bam_feed_parts([files[1][0]])
    assert len(runs) == 1 and runs[0][1] == len(files[1][1]["pos"])
    ctx.bam_end()
    # a part cut in the middle of a BGZF block, and parts whose last record is incomplete
    ctx.bam_begin(1)
    with pytest.raises(api.RsiGpuError):
        ctx.bam_feed_parts([files[0][0][:-7], files[1][0]])
    ctx.bam_end()
    # all or nothing: parts that do not fit one feed's decoded-size limit are refused, not cut
    ctx.set_feed_limit(1 << 17)
    ctx.bam_begin(1)
    with pytest.raises(api.RsiGpuError):
        ctx.bam_feed_parts([f[0] for f in files])
    ctx.bam_end()
    ctx.close()
