#include "bam_reader.hpp"

#include <string.h>
#include <zlib.h>

#include <thread>

namespace rsihost {

namespace {
inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }
inline uint16_t rd16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }

struct Block { size_t coff, clen, uoff, ulen; };

bool inflate_block(const uint8_t* src, size_t clen, uint8_t* dst, size_t ulen) {
  // src points at the 18-byte header; payload = clen - 18 - 8 bytes of raw deflate
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  if (inflateInit2(&zs, -15) != Z_OK) return false;
  zs.next_in = const_cast<Bytef*>(src + 18); zs.avail_in = (uInt)(clen - 26);
  zs.next_out = dst; zs.avail_out = (uInt)ulen;
  int rc = inflate(&zs, Z_FINISH);
  inflateEnd(&zs);
  if (rc != Z_STREAM_END || zs.total_out != ulen) return false;
  return (uint32_t)crc32(crc32(0L, Z_NULL, 0), dst, (uInt)ulen) == rd32(src + clen - 8);
}
}  // namespace

bool BamReader::open(const std::string& path, std::string* err) {
  close();
  f_ = fopen(path.c_str(), "rb");
  if (!f_) { *err = "cannot open " + path; return false; }
  eof_ = false; buf_.clear(); off_ = 0;
  if (!fill(12, err)) return false;
  if (buf_.size() - off_ < 12 || memcmp(&buf_[off_], "BAM\1", 4) != 0) { *err = path + " is not a BAM file"; return false; }
  const uint32_t l_text = rd32(&buf_[off_ + 4]);
  if (!fill(12 + (size_t)l_text, err)) return false;
  const uint32_t n_ref = rd32(&buf_[off_ + 8 + l_text]);
  off_ += 12 + l_text;
  hdr_.name.clear(); hdr_.len.clear();
  for (uint32_t i = 0; i < n_ref; ++i) {
    if (!fill(4, err)) return false;
    const uint32_t l_name = rd32(&buf_[off_]);
    if (!fill(8 + (size_t)l_name, err) || buf_.size() - off_ < 8 + (size_t)l_name) { *err = "truncated BAM header"; return false; }
    hdr_.name.push_back(std::string((const char*)&buf_[off_ + 4], l_name ? l_name - 1 : 0));
    hdr_.len.push_back((int32_t)rd32(&buf_[off_ + 4 + l_name]));
    off_ += 8 + l_name;
  }
  return true;
}

void BamReader::close() {
  if (f_) fclose(f_);
  f_ = nullptr;
}

// read up to ~32 MiB of compressed blocks, inflate them on several threads, append to buf_
bool BamReader::read_block_group(std::string* err) {
  if (eof_) return true;
  std::vector<uint8_t> comp;
  std::vector<Block> blocks;
  size_t utotal = 0;
  const size_t target = (size_t)32 << 20;
  while (comp.size() < target) {
    uint8_t h[18];
    size_t got = fread(h, 1, 18, f_);
    if (got == 0) { eof_ = true; break; }
    if (got != 18 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4) || rd16(h + 10) != 6 || h[12] != 'B' || h[13] != 'C') { *err = "bad BGZF block header"; return false; }
    const size_t clen = (size_t)rd16(h + 16) + 1;
    const size_t at = comp.size();
    comp.resize(at + clen);
    memcpy(&comp[at], h, 18);
    if (fread(&comp[at + 18], 1, clen - 18, f_) != clen - 18) { *err = "truncated BGZF block"; return false; }
    const size_t ulen = rd32(&comp[at + clen - 4]);
    blocks.push_back(Block{at, clen, utotal, ulen});
    utotal += ulen;
  }
  if (blocks.empty()) return true;
  // compact the consumed prefix before growing
  if (off_ > 0) { buf_.erase(buf_.begin(), buf_.begin() + (ptrdiff_t)off_); off_ = 0; }
  const size_t base = buf_.size();
  buf_.resize(base + utotal);
  const int nt = threads_ < 1 ? 1 : threads_;
  std::vector<std::thread> pool;
  std::vector<int> ok(nt, 1);
  for (int t = 0; t < nt; ++t)
    pool.emplace_back([&, t]() {
      for (size_t b = (size_t)t; b < blocks.size(); b += (size_t)nt)
        if (blocks[b].ulen && !inflate_block(&comp[blocks[b].coff], blocks[b].clen, &buf_[base + blocks[b].uoff], blocks[b].ulen)) ok[t] = 0;
    });
  for (auto& th : pool) th.join();
  for (int t = 0; t < nt; ++t) if (!ok[t]) { *err = "BGZF inflate / CRC error"; return false; }
  return true;
}

bool BamReader::fill(size_t want, std::string* err) {
  while (buf_.size() - off_ < want && !eof_) if (!read_block_group(err)) return false;
  return true;
}

bool BamReader::next_contig(ContigReads* out, std::string* err) {
  out->clear();
  out->cigar_off.push_back(0); out->qual_off.push_back(0);
  for (;;) {
    if (!fill(4, err)) return false;
    if (buf_.size() - off_ < 4) break;   // end of file
    const uint32_t bs = rd32(&buf_[off_]);
    if (!fill(4 + (size_t)bs, err)) return false;
    if (buf_.size() - off_ < 4 + (size_t)bs || bs < 32) { *err = "truncated BAM record"; return false; }
    const uint8_t* r = &buf_[off_ + 4];
    const int32_t tid = (int32_t)rd32(r);
    if (tid < 0) { off_ = buf_.size(); eof_ = true; break; }   // unplaced reads come last in a sorted BAM
    if (out->tid >= 0 && tid != out->tid) break;             // next contig starts: leave the record for the next call
    out->tid = tid;
    const uint32_t bmq = rd32(r + 8), fnc = rd32(r + 12);
    const uint32_t l_name = bmq & 0xff, n_cig = fnc & 0xffff, l_seq = rd32(r + 16);
    if (32 + (size_t)l_name + 4 * (size_t)n_cig + (l_seq + 1) / 2 + l_seq > bs) { *err = "corrupt BAM record"; return false; }
    out->pos.push_back((int32_t)rd32(r + 4));
    out->mapq.push_back((uint8_t)((bmq >> 8) & 0xff));
    out->flag.push_back((uint16_t)(fnc >> 16));
    out->mtid.push_back((int32_t)rd32(r + 20));
    out->mpos.push_back((int32_t)rd32(r + 24));
    out->isize.push_back((int32_t)rd32(r + 28));
    const uint8_t* cg = r + 32 + l_name;
    for (uint32_t k = 0; k < n_cig; ++k) out->cigar.push_back(rd32(cg + 4 * k));
    out->cigar_off.push_back((uint32_t)out->cigar.size());
    const uint8_t* q = cg + 4 * n_cig + (l_seq + 1) / 2;
    out->qual.insert(out->qual.end(), q, q + l_seq);
    out->qual_off.push_back((uint64_t)out->qual.size());
    off_ += 4 + (size_t)bs;
  }
  return out->tid >= 0;
}

bool read_bam_header(const std::string& path, BamHeader* hdr, long long* rec_coff, long long* rec_skip, std::string* err) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) { *err = "cannot open " + path; return false; }
  std::vector<uint8_t> dec;
  std::vector<std::pair<long long, size_t>> starts;   // (file offset, decoded offset) of every block inflated so far
  long long foff = 0;
  bool ok = true;
  auto more = [&]() {
    uint8_t h[18];
    if (fread(h, 1, 18, f) != 18 || h[0] != 0x1f || h[1] != 0x8b || h[2] != 8 || !(h[3] & 4) || rd16(h + 10) != 6 || h[12] != 'B' || h[13] != 'C') return false;
    const size_t clen = (size_t)rd16(h + 16) + 1;
    std::vector<uint8_t> comp(clen);
    memcpy(comp.data(), h, 18);
    if (clen < 26 || fread(&comp[18], 1, clen - 18, f) != clen - 18) return false;
    const size_t ulen = rd32(&comp[clen - 4]);
    starts.push_back(std::make_pair(foff, dec.size()));
    const size_t at = dec.size();
    dec.resize(at + ulen);
    if (ulen && !inflate_block(comp.data(), clen, &dec[at], ulen)) return false;
    foff += (long long)clen;
    return true;
  };
  auto need = [&](size_t n) { while (ok && dec.size() < n) ok = more(); return ok; };
  size_t p = 0;
  if (!need(12) || memcmp(dec.data(), "BAM\1", 4) != 0) { *err = path + " is not a BAM file"; fclose(f); return false; }
  const uint32_t l_text = rd32(&dec[4]);
  if (!need(12 + (size_t)l_text)) { *err = "truncated BAM header"; fclose(f); return false; }
  const uint32_t n_ref = rd32(&dec[8 + l_text]);
  p = 12 + (size_t)l_text;
  hdr->name.clear(); hdr->len.clear();
  for (uint32_t i = 0; i < n_ref; ++i) {
    if (!need(p + 4)) break;
    const uint32_t l_name = rd32(&dec[p]);
    if (!need(p + 8 + (size_t)l_name)) break;
    hdr->name.push_back(std::string((const char*)&dec[p + 4], l_name ? l_name - 1 : 0));
    hdr->len.push_back((int32_t)rd32(&dec[p + 4 + l_name]));
    p += 8 + l_name;
  }
  fclose(f);
  if (!ok) { *err = "truncated BAM header"; return false; }
  // the block that holds decoded offset p; when the header ends exactly at a block end, the next block
  *rec_coff = foff; *rec_skip = 0;
  for (size_t i = 0; i < starts.size(); ++i) {
    const size_t lo = starts[i].second, hi = i + 1 < starts.size() ? starts[i + 1].second : dec.size();
    if (lo <= p && p < hi) { *rec_coff = starts[i].first; *rec_skip = (long long)(p - lo); }
  }
  return true;
}

// BAI layout (SAM spec 5.2; written by bam_index_save, bam_index.c:262-319): magic "BAI\1", n_ref, then per reference
// n_bin x {bin, n_chunk, n_chunk x (beg, end)}, n_intv, n_intv x ioffset.  Bin 37450 is samtools' pseudo-bin (file range of
// the reference + mapped/unmapped counts), not a chunk list.
bool read_bai(const std::string& bam_path, size_t n_ref, std::vector<BaiRef>* out) {
  FILE* f = fopen((bam_path + ".bai").c_str(), "rb");
  if (!f) {
    std::string alt = bam_path;
    if (alt.size() > 4 && alt.compare(alt.size() - 4, 4, ".bam") == 0) { alt.replace(alt.size() - 4, 4, ".bai"); f = fopen(alt.c_str(), "rb"); }
    if (!f) return false;
  }
  std::vector<uint8_t> d;
  uint8_t tmp[1 << 16];
  size_t got;
  while ((got = fread(tmp, 1, sizeof tmp, f)) > 0) d.insert(d.end(), tmp, tmp + got);
  fclose(f);
  size_t p = 0;
  auto u32 = [&](uint32_t* v) { if (p + 4 > d.size()) return false; *v = rd32(&d[p]); p += 4; return true; };
  auto u64 = [&](uint64_t* v) { if (p + 8 > d.size()) return false; *v = (uint64_t)rd32(&d[p]) | ((uint64_t)rd32(&d[p + 4]) << 32); p += 8; return true; };
  uint32_t nref = 0;
  if (d.size() < 8 || memcmp(d.data(), "BAI\1", 4) != 0) return false;
  p = 4;
  if (!u32(&nref) || nref != n_ref) return false;
  out->assign(n_ref, BaiRef());
  for (uint32_t r = 0; r < nref; ++r) {
    uint32_t nbin = 0;
    if (!u32(&nbin)) return false;
    uint64_t first = ~0ull, last = 0;
    for (uint32_t b = 0; b < nbin; ++b) {
      uint32_t bin = 0, nchunk = 0;
      if (!u32(&bin) || !u32(&nchunk)) return false;
      for (uint32_t k = 0; k < nchunk; ++k) {
        uint64_t beg = 0, end = 0;
        if (!u64(&beg) || !u64(&end)) return false;
        if (bin != 37450 && beg < first) first = beg;
        if (bin == 37450 && k == 0) last = end;          // (ref_beg, ref_end) of the pseudo-bin
      }
    }
    uint32_t nintv = 0;
    if (!u32(&nintv)) return false;
    if (p + 8 * (size_t)nintv > d.size()) return false;
    p += 8 * (size_t)nintv;
    if (first != ~0ull) { (*out)[r].has_reads = true; (*out)[r].first_voff = first; (*out)[r].end_voff = last >= first ? last : 0; }
  }
  return true;
}

}  // namespace rsihost
