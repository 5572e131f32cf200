// bamio.h — host-side BAM/BAI/BGZF container logic of the B200 coverage path (no decompression here: every
// inflate, including the header's, runs on the device through K1).
//
// Mirrors what the reference gets from noodles + BamStats (bamnado/src/bam_utils.rs:311-436): header references,
// BAI metadata pseudo-bin counts, the chunk length estimate and the genomic chunk geometry.
#pragma once
#include <stdint.h>

#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

namespace bnd {

enum ErrClass { ERR_RUNTIME = 1, ERR_NOT_FOUND = 2, ERR_INVALID = 3 };
struct Error : std::runtime_error {
  int cls;
  Error(int c, const std::string &m) : std::runtime_error(m), cls(c) {}
};
[[noreturn]] inline void fail(const std::string &m, int cls = ERR_RUNTIME) { throw Error(cls, m); }

struct BamHeader {
  std::vector<std::string> names;
  std::vector<uint64_t> lens;
  std::unordered_map<std::string, int> name_to_id;  // last occurrence wins, like a HashMap insert
  uint64_t first_record_voff = 0;                   // virtual offset of the first alignment record
  std::string text;                                 // SAM header text (the BAM writers copy it and append their @PG line)
};

struct BaiChunk {
  uint64_t beg, end;
};
struct BaiRef {
  std::vector<BaiChunk> chunks;  // all chunks of all real bins (file order)
  std::vector<uint64_t> linear;  // 16 kb linear index
  bool has_meta = false;
  uint64_t mapped = 0, unmapped = 0;
  uint64_t min_voff = ~0ull, max_voff = 0;  // span of the reference's records in the file
};
struct Bai {
  std::vector<BaiRef> refs;
  bool has_no_coor = false;
  uint64_t n_no_coor = 0;
};

Bai read_bai(const std::string &path);

// BGZF block table of a byte range that starts at a block boundary.
struct BlockTable {
  std::vector<uint64_t> coff;  // [n + 1] offsets relative to the range start; coff[n] = bytes covered by whole blocks
  std::vector<uint64_t> uoff;  // [n + 1] prefix sums of ISIZE
};
// Walks the BSIZE chain over buf[0, n).  Stops before a block that is not entirely inside the range.
// Throws on a malformed block header.
void scan_bgzf_blocks(const uint8_t *buf, uint64_t n, BlockTable &out, uint64_t max_blocks = ~0ull);

// Parses the BAM header from inflated bytes.  Returns false when `data` does not yet hold the whole header.
// `consumed` is the header length in inflated bytes.
bool parse_bam_header(const uint8_t *data, uint64_t n, BamHeader &out, uint64_t &consumed);

// BamStats::new + estimate_genome_chunk_length + genome_chunks (bam_utils.rs:311-436)
struct Geometry {
  std::vector<uint8_t> in_stats;         // references that have chunk statistics (zip of header and index refs)
  uint64_t n_mapped = 0, n_unmapped = 0, genome_len = 0;
  uint64_t bin = 0, L = 0, nbf = 0;
  std::vector<uint64_t> ref_first_bin;   // [n_ref + 1]
  std::vector<uint64_t> ref_first_chunk; // [n_ref + 1]
  uint64_t total_bins() const { return ref_first_bin.empty() ? 0 : ref_first_bin.back(); }
  uint64_t total_chunks() const { return ref_first_chunk.empty() ? 0 : ref_first_chunk.back(); }
  uint64_t chunks_of(int r) const { return ref_first_chunk[r + 1] - ref_first_chunk[r]; }
  // global bin index of the first bin of chunk ordinal `c` (c may equal total_chunks())
  uint64_t first_bin_of_chunk(uint64_t c) const;
  int ref_of_chunk(uint64_t c) const;
};
// drop_scaffold_refs: scaffold contigs get no chunks at all, so their reads are neither counted nor visited
// (bin_counts.rs:190-200 filters the chunk list; BamPileup only drops their rows at output time)
Geometry make_geometry(const BamHeader &h, const Bai &bai, uint64_t bin_size, bool drop_scaffold_refs = false);

// is_scaffold (bam_utils.rs:38-43): ^chrUn | _alt$ | _random$
bool is_scaffold(const std::string &name);

// Smallest virtual offset that can hold a record overlapping 1-based position `pos1` of reference `ref` or any later
// record (linear index with conservative fallbacks); ~0ull when nothing at or after that point exists in the index.
uint64_t voff_lower_bound(const Bai &bai, int ref, uint64_t pos1);

// ryu-style shortest round-trip text of an f64, as polars' CsvWriter prints a Float64 column
std::string format_f64(double v);

}  // namespace bnd
