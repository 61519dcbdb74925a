"""TEST INFRASTRUCTURE -- not part of the product (only tests/, smoke() and bench.py's cpu_baseline leg may import it).

CPU restatement of how the reference reads a BAM file on the `rsicnv rsi -b` path, which it does through its vendored
samtools-0.1.18 (file:line relative to the reference's src/samtools-0.1.18):

  BGZF container   bgzf.c:56-70 (header constants), 258-275 check_header (gzip magic 1f 8b, CM 8, FLG.FEXTRA, XLEN, 'B','C'
                   subfield with BSIZE), 277-313 inflate_block (raw deflate, window -15; CRC32 is NOT checked when reading),
                   401-411 / 471-523 read_block + bgzf_read (blocks are concatenated into one byte stream)
  BAM header       bam.c:69-110 bam_header_read ("BAM\\1", l_text, text, n_ref, then l_name/name/l_ref per reference)
  alignment record bam.c:179-210 bam_read1 (block_size, then the 32-byte core of bam.h:131-155: refID, pos, bin_mq_nl,
                   flag_nc, l_seq, next_refID, next_pos, tlen; then read name, CIGAR words, 4-bit bases, qualities, aux)

decode(data) returns the header and, per refID in file order, the structure of arrays the C ABI takes
(rsigpu_read_batch: pos, mpos, isize, mtid, flag, mapq, cigar_off, cigar, qual_off, qual).  Pinned against the reference's own
samtools by tests/test_oracle_vs_reference.py::test_bam_decode_matches_reference_samtools and by the golden fixture
tests/golden/tiny_rich.bam / tiny_rich.json (hashes of what the reference's bam_read1 returns)."""
from __future__ import annotations

import struct
import zlib

import numpy as np

FIELDS = (("pos", np.int32), ("mpos", np.int32), ("isize", np.int32), ("mtid", np.int32), ("flag", np.uint16), ("mapq", np.uint8),
          ("cigar_off", np.uint32), ("cigar", np.uint32), ("qual_off", np.uint64), ("qual", np.uint8))


def bgzf_stream(data) -> bytes:
    """bgzf.c: every block is an independent raw-deflate stream; the file is their concatenation (an empty block ends it)"""
    mv = memoryview(data)
    out = []
    off = 0
    while off < len(mv):
        h = bytes(mv[off:off + 12])
        if len(h) < 12 or h[0] != 0x1f or h[1] != 0x8b or h[2] != 8 or not (h[3] & 4):
            raise ValueError("not a BGZF block header at %d" % off)
        xlen = struct.unpack_from("<H", h, 10)[0]
        extra = bytes(mv[off + 12:off + 12 + xlen])
        bsize = None
        x = 0
        while x + 4 <= xlen:
            sl = struct.unpack_from("<H", extra, x + 2)[0]
            if extra[x] == ord("B") and extra[x + 1] == ord("C") and sl == 2:
                bsize = struct.unpack_from("<H", extra, x + 4)[0] + 1
            x += 4 + sl
        if bsize is None:
            raise ValueError("gzip member without a BGZF size field")
        isize = struct.unpack_from("<I", bytes(mv[off + bsize - 4:off + bsize]))[0]
        block = zlib.decompress(bytes(mv[off + 12 + xlen:off + bsize - 8]), -15)
        if len(block) != isize:
            raise ValueError("ISIZE mismatch")
        out.append(block)
        off += bsize
    return b"".join(out)


def decode(data) -> dict:
    s = bgzf_stream(data)
    if s[:4] != b"BAM\x01":
        raise ValueError("not a BAM file")
    l_text = struct.unpack_from("<i", s, 4)[0]
    n_ref = struct.unpack_from("<i", s, 8 + l_text)[0]
    p = 12 + l_text
    names, lens = [], []
    for _ in range(n_ref):
        ln = struct.unpack_from("<i", s, p)[0]
        names.append(s[p + 4:p + 4 + ln - 1].decode()); lens.append(struct.unpack_from("<i", s, p + 4 + ln)[0])
        p += 8 + ln
    per_tid: dict[int, dict] = {}
    order = []
    buf = np.frombuffer(s, np.uint8)
    while p + 4 <= len(s):
        bs = struct.unpack_from("<i", s, p)[0]
        if bs < 32 or p + 4 + bs > len(s):
            raise ValueError("truncated or corrupt alignment record at %d" % p)
        tid, pos, bmq, fnc, l_seq, mtid, mpos, isize = struct.unpack_from("<iiIIiiii", s, p + 4)
        l_name = bmq & 0xff; n_cig = fnc & 0xffff
        if tid not in per_tid:
            per_tid[tid] = {k: [] for k, _ in FIELDS}
            order.append(tid)
        R = per_tid[tid]
        R["pos"].append(pos); R["mpos"].append(mpos); R["isize"].append(isize); R["mtid"].append(mtid)
        R["flag"].append(fnc >> 16); R["mapq"].append((bmq >> 8) & 0xff)
        c0 = p + 36 + l_name
        R["cigar"].append(buf[c0:c0 + 4 * n_cig].view("<u4"))
        q0 = c0 + 4 * n_cig + (l_seq + 1) // 2
        R["qual"].append(buf[q0:q0 + l_seq])
        R["cigar_off"].append(n_cig); R["qual_off"].append(l_seq)
        p += 4 + bs
    if p != len(s):
        raise ValueError("trailing bytes after the last record")
    reads = {}
    for tid in order:
        R = per_tid[tid]
        out = {}
        for k, dt in FIELDS:
            if k in ("cigar", "qual"):
                out[k] = np.concatenate(R[k]).astype(dt) if R[k] else np.zeros(0, dt)
            elif k in ("cigar_off", "qual_off"):
                out[k] = np.concatenate(([0], np.cumsum(np.asarray(R[k], np.int64)))).astype(dt)
            else:
                out[k] = np.asarray(R[k]).astype(dt)
        reads[tid] = out
    return {"names": names, "lens": lens, "order": order, "reads": reads}
