// bam_reader.hpp -- host-side BGZF / BAM decoder for the `rsicnv rsi -b` path.
//
// Replaces what the reference takes from its vendored samtools-0.1.18:
//   BGZF blocks        bgzf.c:56-70, 401-411, 471-523   (18-byte gzip header with the BC extra field, raw deflate, CRC32 + ISIZE)
//   BAM header/records bam.c:69-110, 179-210, bam.h:131-155
// The reference walks one contig at a time through the BAI index (bam_iter_query over [0, 2^31-1), which
// yields every record of the contig in file order); a coordinate-sorted BAM holds each contig's records
// contiguously, so this reader streams the file once, inflates groups of BGZF blocks on several host
// threads, and hands back one position-sorted structure-of-arrays batch per contig -- exactly the layout
// rsigpu_read_batch takes.  No index is needed.
#pragma once
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

namespace rsihost {

struct ContigReads {   // SoA of the records of one contig, file (= position) order
  int tid = -1;
  std::vector<int32_t> pos, mpos, isize, mtid;
  std::vector<uint16_t> flag;
  std::vector<uint8_t> mapq, qual;
  std::vector<uint32_t> cigar_off, cigar;
  std::vector<uint64_t> qual_off;
  size_t n() const { return pos.size(); }
  void clear() {
    tid = -1; pos.clear(); mpos.clear(); isize.clear(); mtid.clear(); flag.clear(); mapq.clear(); qual.clear(); cigar_off.clear(); cigar.clear(); qual_off.clear();
  }
};

struct BamHeader {
  std::vector<std::string> name;
  std::vector<int32_t> len;
};

class BamReader {
 public:
  explicit BamReader(int threads = 8) : threads_(threads) {}
  ~BamReader() { close(); }
  bool open(const std::string& path, std::string* err);
  void close();
  const BamHeader& header() const { return hdr_; }
  // next contig that has at least one record (records with refID < 0 end the stream); false at end of file
  bool next_contig(ContigReads* out, std::string* err);

 private:
  bool fill(size_t want, std::string* err);   // make at least `want` decoded bytes available (if the file has them)
  bool read_block_group(std::string* err);
  int threads_;
  FILE* f_ = nullptr;
  BamHeader hdr_;
  std::vector<uint8_t> buf_;   // decoded bytes not yet consumed
  size_t off_ = 0;             // consume offset into buf_
  bool eof_ = false;
};

// Only the header (bam_header_read, bam.c:69-110), inflating block by block: names and lengths, plus where the alignment
// records start -- rec_coff = file offset of the BGZF block that holds the first record, rec_skip = decoded bytes of that
// block in front of it.  This is what the GPU decoder (rsigpu_bam_feed) needs from the host.
bool read_bam_header(const std::string& path, BamHeader* hdr, long long* rec_coff, long long* rec_skip, std::string* err);

// The BAI index next to the BAM (bam_index.c:321-465 loads it; the reference REQUIRES it: rsi.cpp:2112-2113).  Only what the
// chromosome-sharded GPU path needs: for every reference sequence whether it has records (the reference's "populated" test,
// rsi.cpp:2121-2126) and the virtual file offset (BGZF block offset << 16 | offset inside the decoded block) of its first
// record, so that every GPU can start decoding ITS contigs in the middle of the file.  false if the file is absent or malformed.
struct BaiRef { bool has_reads = false; uint64_t first_voff = 0, end_voff = 0; };   // end_voff: where the last record ends (samtools' pseudo-bin), 0 = unknown
bool read_bai(const std::string& bam_path, size_t n_ref, std::vector<BaiRef>* out);

}  // namespace rsihost
