// kernel_args.h — plain-data argument blocks shared by the CUDA kernels (kernels.cu) and the C++ host engine.
// No device code here, so host translation units can include it without nvcc.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BND_HD __host__ __device__
#else
#define BND_HD
#endif

namespace bnd {

// status codes written per block
enum InflateStatus : int32_t {
  INF_OK = 0,
  INF_BAD_BTYPE = 1,
  INF_BAD_STORED = 2,
  INF_BAD_CODELEN = 3,
  INF_BAD_SYMBOL = 4,
  INF_BAD_DISTANCE = 5,
  INF_OUTPUT_OVERRUN = 6,
  INF_INPUT_OVERRUN = 7,
  INF_SIZE_MISMATCH = 8,
  INF_CRC_MISMATCH = 9,
  INF_BAD_HEADER = 10,
};

struct CrcTables {
  uint32_t t[256];        // byte table of the reflected CRC-32 (poly 0xEDB88320)
  uint32_t z[4][256];     // "advance by 32*G zero bits" operator, sliced by byte
};

// K1 stages its input with bulk copies of one round (1.3 KB) that may start up to a round before the end of the last block:
// every buffer of compressed bytes is allocated with this many readable bytes behind the data
constexpr size_t COMP_PAD = 2048;

struct InflateArgs {
  const uint8_t *comp;        // compressed bytes of the segment (padded by COMP_PAD readable bytes)
  const uint64_t *blk_coff;   // [n+1] offset of each BGZF block (gzip header) within comp
  const uint64_t *blk_uoff;   // [n+1] offset of each block's output within out
  uint8_t *out;               // inflated bytes
  int32_t *status;            // [n] per-block status, or nullptr
  int32_t *error_flag;        // sticky: first non-zero status (atomicCAS)
  uint32_t *work_counter;     // dynamic block scheduler (zeroed before launch)
  uint32_t *out_len;          // zlib streams only: [n] inflated size of each stream (blk_uoff then holds capacities)
  uint32_t coff_stride;       // zlib streams only: 2 = blk_coff holds (begin, end) pairs instead of consecutive offsets
  const CrcTables *crc;       // device-resident tables for this G
  uint32_t n_blocks;
  int verify_crc;
};

constexpr uint32_t REC_TILE = 16384;   // bytes of inflated stream per tile
constexpr uint32_t REC_SLOTS = 456;    // >= ceil(REC_TILE / 36): a record is at least 4 + 32 bytes
constexpr uint64_t REC_NONE = ~0ull;   // "no / invalid offset"
constexpr uint32_t REC_MAX_SIZE = 1u << 28;

enum SegError : int32_t {
  SEG_OK = 0,
  SEG_ERR_BAD_RECORD = 1,      // block_size < 32 on the true chain
  SEG_ERR_CARRY_TOO_LARGE = 2, // a record larger than the carry area straddles two segments
};

// Device-resident state chained from segment to segment of one pipeline.
struct SegState {
  uint64_t entry_off;      // buffer offset of the first record of this segment
  uint64_t next_entry_off; // where the next segment's first record starts in *its* buffer (carry_cap - carry_len)
  uint64_t n_records;      // records fully contained in the segment just indexed
  uint64_t carry_len;      // bytes of the trailing incomplete record of the segment just indexed
  uint64_t carry_src;      // buffer offset of those bytes
  uint64_t total_records;  // running total over all segments
  uint64_t last_ref_pos;   // (ref_id + 1) << 32 | (pos + 1) of the last complete record seen (range termination)
  int32_t error;           // sticky SegError
  uint32_t fix_rounds;     // diagnostics: fixed-point rounds and re-walked tiles
  uint32_t rewalked;
  uint32_t pad;
};

struct RecArgs {
  const uint8_t *buf;      // segment buffer: [carry area][inflated bytes]
  uint64_t tile0;          // buffer offset of tile 0 (tiles cover [tile0, data_end))
  uint64_t data_end;       // buffer offset one past the last inflated byte
  uint32_t n_tiles;
  int32_t n_ref;
  uint64_t *tile_entry;    // [n_tiles]
  uint64_t *tile_exit;     // [n_tiles]
  uint32_t *tile_count;    // [n_tiles]
  uint8_t *tile_flag;      // [n_tiles] scratch of the fix pass
  uint64_t *tile_base;     // [n_tiles] first record index of the tile (after the scan)
  uint16_t *tile_slots;    // [n_tiles][REC_SLOTS]
  uint64_t *rec_off;       // [capacity] buffer offsets of the records, in stream order
  SegState *st;
  uint64_t carry_cap;      // size of the carry area of the *next* buffer
};

enum {
  ST_N_TOTAL = 0, ST_FAIL_PROPER_PAIR, ST_FAIL_MAPQ, ST_FAIL_LENGTH, ST_FAIL_BLACKLIST, ST_FAIL_BARCODE,
  ST_FAIL_READ_GROUP, ST_FAIL_STRAND, ST_FAIL_TAG, ST_FAIL_FRAGMENT_LENGTH, ST_FAIL_DUPLICATE, ST_FAIL_SECONDARY,
  ST_FAIL_SUPPLEMENTARY, ST_N_COUNTERS
};

struct FilterDev {
  int32_t strand;  // 0 both, 1 forward, 2 reverse
  int32_t proper_pair;
  uint32_t min_mapq, min_length, max_length;
  int32_t has_min_frag, has_max_frag;
  uint32_t min_frag, max_frag;
  int32_t ex_dup, ex_sec, ex_sup;
  int32_t has_rg, rg_len;
  const char *rg;
  int32_t has_tag;      // both --tag and --tag-value were given
  int32_t tag_name_ok;  // tag name has exactly two characters
  char tag0, tag1;
  int32_t tagv_len;
  const char *tagv;
  int32_t has_blacklist;
  const uint64_t *bl_off;     // [n_ref + 1] slice of bl_starts / bl_stops per reference
  const uint64_t *bl_starts;  // per reference: interval starts, sorted
  const uint64_t *bl_stops;   // per reference: interval stops, sorted independently
  int32_t has_barcodes;
  uint32_t bc_mask;           // table size - 1 (power of two)
  const uint32_t *bc_table;   // open addressing: 0 = empty, else barcode index + 1
  const uint32_t *bc_off;     // [n + 1] offsets into bc_pool
  const char *bc_pool;
};

struct PileDev {
  int32_t use_fragment;
  int32_t shift_on;
  int32_t shift[4];
  int32_t has_trunc, t_has5, t_has3;
  uint64_t t5, t3;
};

// Chunk / bin geometry (bamnado/src/bam_utils.rs:384-436 + bin_intervals).  Bins of all chunks of all references
// are laid out back to back in header order; chunk c of reference r starts at ref_first_bin[r] + c * nbf.
struct GeomDev {
  int32_t n_ref;
  const uint64_t *ref_len;          // [n_ref]
  const uint64_t *ref_first_bin;    // [n_ref + 1]
  const uint64_t *ref_first_chunk;  // [n_ref + 1] ordinal of the reference's first chunk (refs without chunks repeat)
  uint64_t L;                       // chunk length
  uint64_t bin;                     // bin size
  uint64_t nbf;                     // bins of a full chunk = ceil((L - 1) / bin)
  uint64_t chunk_lo, chunk_hi;      // chunk ordinals owned by this shard: [lo, hi)
  uint64_t bin_lo, bin_hi;          // global bin range of those chunks; the diff array has bin_hi - bin_lo + 1 slots
};

constexpr int PILE_THREADS = 128;
constexpr int PILE_RPT = 4;                      // records per thread per tile
constexpr uint32_t PILE_WIN = 2048;              // bins in the CTA-private window
constexpr uint32_t PILE_WIN_MARGIN = PILE_WIN / 8;

struct PileArgs {
  const uint8_t *buf;
  const uint64_t *rec_off;
  const SegState *st;            // n_records of this segment
  int32_t *diff;                 // [bin_hi - bin_lo + 1]
  unsigned long long *stats;     // [ST_N_COUNTERS]
  unsigned long long *range_probe;  // [2]: max (ref+1)<<32|(pos+1) seen, records visited
  FilterDev f;
  PileDev p;
  GeomDev g;
  int use_window;
};

constexpr int FIN_THREADS = 256;
constexpr int FIN_IPT = 16;  // bins per thread
constexpr uint32_t FIN_TILE = FIN_THREADS * FIN_IPT;

struct FinState {
  unsigned long long n_runs;
  unsigned int max_val;
  unsigned int pad;
  unsigned long long text_bytes;
};

struct FinArgs {
  const int32_t *diff;   // [n_bins (+1)]
  uint64_t n_bins;       // bins of this shard (bin_hi - bin_lo)
  GeomDev g;
  int collapse;
  int32_t *tile_sum;     // [n_tiles] -> exclusive carry after the tile scan
  uint64_t *tile_heads;  // [n_tiles] -> exclusive run base after the tile scan
  uint32_t n_tiles;
  uint64_t *run_bin;     // [n_runs] global bin index of each run head
  uint32_t *run_val;     // [n_runs] count
  uint32_t *counts;      // optional dense per-bin counts [n_bins]
  FinState *fs;
};

struct RunRow {  // == bnd_run of the C ABI
  int32_t ref_id;
  uint32_t val;
  uint64_t start, stop;
};

struct RunArgs {
  const uint64_t *run_bin;
  const uint32_t *run_val;
  uint64_t n_runs;
  GeomDev g;
  RunRow *rows;
};

struct SignalArgs {
  const uint32_t *counts;  // dense per-bin counts of the shard
  GeomDev g;
  int32_t ref;
  uint64_t chrom_size;
  float scale;
  float *out;
};

struct TextArgs {
  const uint64_t *run_bin;
  const uint32_t *run_val;
  uint64_t n_runs;
  GeomDev g;
  const char *names;         // concatenated reference names
  const uint32_t *name_off;  // [n_ref + 1]
  const uint8_t *ref_skip;   // [n_ref] 1 = drop rows of this reference (--ignore-scaffold)
  const char *score_text;    // [n_scores][SCORE_W]: byte 0 = length, then the text
  uint32_t n_scores;
  uint16_t *row_len;         // [n_rows]
  uint64_t *tile_off;        // [n_text_tiles] -> exclusive offsets after the scan
  uint32_t n_tiles;
  char *text;
  FinState *fs;
  uint64_t *ref_text_off;    // [n_ref + 1] byte offset of each reference's first row
};
constexpr int SCORE_W = 32;
constexpr int TEXT_THREADS = 256;

// collapse-bedgraph rows (sorted by chrom rank, start) and their collapsed groups
struct BdgArgs {
  const uint32_t *chrom;  // rank of the chromosome name in byte-string order
  const long long *start, *end;
  const double *score;
  uint64_t n;
  uint32_t n_tiles;
  uint64_t *tile_heads;
  unsigned int *unsorted;
  uint32_t *out_chrom;
  long long *out_start, *out_end;  // pre-filled with +inf / -inf
  double *out_score;
};

// ---- collapse-bedgraph reader on the device (bedgraph_parse.cuh) ------------------------------------------------------------
constexpr int COL_TILE = 256;  // rows per CTA pass of the head-flag / compaction kernels (collapse.cuh: COL_THREADS)
constexpr int BDG_THREADS = 256;
constexpr uint32_t BDG_PER_THREAD = 16;
constexpr uint32_t BDG_TILE = BDG_THREADS * BDG_PER_THREAD;  // bytes per newline tile

struct BdgSlow {  // a score the device does not convert itself: the host parses text[off, off + len)
  uint64_t line, off;
  uint32_t len, pad;
};

struct BdgParseArgs {
  const uint8_t *text;
  uint64_t n_bytes;
  uint64_t *tile_count;  // [n_tiles] newlines per tile of BDG_TILE bytes -> exclusive offsets after the scan
  uint32_t n_tiles;
  uint64_t *line_off;    // [n_lines + 1]
  uint64_t n_lines;
  // per line
  uint8_t *state;        // 0: skipped (blank, '#', fewer than four fields), 1: row, 2: row whose score the host converts, 3: bad integer
  long long *start, *end;
  double *score;
  uint64_t *name_off;
  uint32_t *name_len;
  BdgSlow *slow;
  unsigned long long *n_slow;
  uint64_t slow_cap;
  unsigned long long *bad_line;  // smallest line index with an unparsable start / end (initialised to ~0)
};

struct BdgRowArgs {  // lines -> rows (skipped lines dropped)
  const uint8_t *state;
  uint64_t n_lines;
  uint64_t *tile_count;
  uint32_t n_tiles;
  const long long *start, *end;
  const double *score;
  const uint64_t *name_off;
  const uint32_t *name_len;
  long long *o_start, *o_end;
  double *o_score;
  uint64_t *o_name_off;
  uint32_t *o_name_len;
};

struct BdgNameArgs {  // chromosome names -> ranks
  const uint8_t *text;
  const uint64_t *name_off;
  const uint32_t *name_len;
  uint64_t n;
  uint8_t *flag;           // [n] 1: the name differs from the previous row's
  uint64_t *tile_count;
  uint32_t n_tiles;
  uint64_t *seg_name_off;  // [n_segments]
  uint32_t *seg_name_len;
  const uint32_t *seg_rank;  // [n_segments] rank of the segment's name in byte-string order
  uint32_t *chrom;           // [n]
};

constexpr int MULTI_MAX_BAMS = 64;
struct MultiArgs {
  const uint32_t *counts[MULTI_MAX_BAMS];  // dense per-bin counts of each BAM (identical geometry)
  int n_bams;
  uint64_t n_bins;
  GeomDev g;
  uint32_t n_tiles;
  uint64_t *tile_heads;
  uint64_t *run_bin;
  uint32_t *out_vals;  // [n_bams][n_runs_cap]
  uint64_t n_runs_cap;
  const uint32_t *bits;  // optional: precomputed head flags, one bit per bin (OR-ed over ranks when every rank holds one BAM)
  uint32_t *bits_out;    // multi_bits_kernel output: ceil(n_bins / 32) words
};

// TSV rows of MultiBamPileup::to_tsv formatted on the device: name \t start \t end (\t count)* \n
struct MultiTextArgs {
  const uint64_t *run_bin;   // [n_runs] head bin of every row of the table
  uint64_t n_runs;
  uint64_t row_lo, n_rows;   // the slice of rows formatted by this call: [row_lo, row_lo + n_rows)
  const uint32_t *vals;      // [n_cols][stride]: counts of the slice's rows, one column per BAM
  uint64_t stride;
  int n_cols;
  uint8_t col_pos[MULTI_MAX_BAMS];  // column c of the output (BAM c) is vals[col_pos[c]][..]
  GeomDev g;
  const char *names;         // concatenated reference names
  const uint32_t *name_off;  // [n_ref + 1]
  uint16_t *row_len;         // [n_rows]
  uint64_t *tile_off;        // [n_tiles] -> exclusive offsets after the scan
  uint32_t n_tiles;
  char *text;
  uint64_t *ref_text_off;    // [n_ref + 1] byte offset of each reference's first row
};

// bin_counts::count_bins (bin_counts.rs:254-262): one sample's dense per-bin counts scattered into its column of the
// shared (n_bins x n_samples) matrix; rows follow the grid of chromosomes sorted by name
struct BinMatrixArgs {
  const uint32_t *counts;         // dense counts of bins [g.bin_lo, g.bin_hi)
  GeomDev g;
  const uint64_t *grid_row0;      // [n_ref] first matrix row of each reference, ~0 = not on the grid
  unsigned long long *matrix;     // [n_rows][n_samples]
  uint64_t bins_per_chunk_grid;   // L / bin: grid bins spanned by one genomic chunk
  uint32_t n_samples, sample;
};


// ---- BigWig on the device (bigwig.cuh) ---------------------------------------------------------------------------------
constexpr uint32_t BW_ITEMS = 1024;                          // items per section (itemsPerSlot of bigtools / UCSC writers)
constexpr uint32_t BW_ITEM_BYTES = 12;                       // bedGraph item: u32 start, u32 end, f32 value
constexpr uint32_t BW_SEC_HEADER = 24;
constexpr uint32_t BW_ZOOM_BYTES = 32;                       // zoom record: chrom, start, end, valid, min, max, sum, sumsq
constexpr uint32_t BW_RAW_MAX = BW_ZOOM_BYTES * BW_ITEMS;    // largest uncompressed section (zoom): 32 KiB
// worst case of the fixed-Huffman stream: 2 (zlib header) + (3 + 9 bits per byte + 7 end-of-block, rounded up) + 4 (Adler-32)
BND_HD constexpr uint32_t bw_comp_bound(uint32_t raw) { return 2 + (3 + 9 * raw + 7 + 7) / 8 + 4; }
constexpr int BW_THREADS = 256;

struct BwItem {  // what a bedGraph section stores per interval; coordinates are 0-based (both ends - 1)
  uint32_t start, end;
  float val;
};

struct BwItemArgs {  // runs -> items of the kept references
  const uint64_t *run_bin;
  const uint32_t *run_val;
  uint64_t n_runs;
  GeomDev g;
  const uint64_t *row_first;   // [n_ref + 1] first run of each reference
  const uint64_t *item_first;  // [n_ref + 1] first item of each reference (dropped references own no items)
  const uint8_t *ref_skip;     // [n_ref]
  double factor;
  BwItem *items;
  unsigned int *range_error;   // set when a coordinate does not fit the format's u32
};

struct BwSection {  // one section to encode: `count` units starting at unit `first`
  uint64_t first;
  uint32_t count;
  uint32_t chrom;  // dense chromosome id
};

struct BwDeflateArgs {
  const uint8_t *units;         // BwItem[] or zoom records
  const BwSection *sections;
  uint32_t n_sections;
  uint32_t unit_bytes;          // 12 or 32
  uint32_t header_bytes;        // 24 (bedGraph section header, built here) or 0
  uint32_t stride;              // bytes reserved per section in `staging`
  uint32_t dist[4];             // match distances tried at every byte (0 = unused): record-structured data repeats at fixed lags
  int compress;                 // 0: sections are stored raw
  uint8_t *staging;             // [n_sections][stride]
  uint32_t *comp_size;          // [n_sections]
  uint32_t *bounds;             // [n_sections][2]: first start, last end (for the R-tree)
  double *stats;                // [n_sections][5]: covered bases, min, max, sum, sum of squares (bedGraph sections only)
};

struct BwCompactArgs {
  const uint8_t *staging;
  const uint32_t *comp_size;
  const uint64_t *offset;  // [n_sections] exclusive scan of comp_size
  uint32_t n_sections, stride;
  uint8_t *blob;
};

struct BwZoomArgs {
  const BwItem *items;
  const uint64_t *item_first;  // [n_chrom + 1] per dense chromosome
  const uint64_t *win_first;   // [n_chrom + 1] first window of each chromosome at this level
  uint32_t n_chrom;
  uint32_t reduction;
  uint64_t n_windows;
  uint8_t *recs;               // [n_windows][32] dense, one slot per window
  uint8_t *flags;              // [n_windows] 1 = the window holds data
  uint64_t *tile_count;        // [n_tiles] windows with data per tile of BW_THREADS -> exclusive offsets after the scan
  uint32_t n_tiles;
  uint8_t *out;                // compacted records
  uint64_t *chrom_first_rec;   // [n_chrom + 1] first compacted record of each chromosome
  const uint8_t *prev_recs;    // level below (dense slots, reduction / 4), or null: build this level from the items
  const uint8_t *prev_flags;
  const uint64_t *prev_win_first;
};

// ---- bigwig-compare / bigwig-aggregate (bigwig_tools.cuh) --------------------------------------------------------------------
struct BwrEmitArgs {  // inflated sections -> items
  const uint8_t *data;
  const uint64_t *sec_off;   // [n_secs] byte offset of every section in `data`
  const uint64_t *item_off;  // [n_secs] first item of every section
  uint32_t n_secs;
  BwItem *items;
  unsigned int *bad;         // bit 0: empty interval, bit 1: overlapping / unsorted intervals
};

constexpr int BWT_MAX_FILES = 64;
enum : int { BWT_SUBTRACTION = 0, BWT_RATIO = 1, BWT_LOGRATIO = 2, BWT_SUM = 10, BWT_MEAN = 11, BWT_MAX = 12, BWT_MIN = 13, BWT_MEDIAN = 14 };

struct BwtArgs {
  int n_files, mode;
  const BwItem *items[BWT_MAX_FILES];
  const uint64_t *range[BWT_MAX_FILES];  // [n_chrom][2]: items of output chromosome c in file f
  double scale[BWT_MAX_FILES];
  double pseudocount;
  uint32_t n_chrom, bin_size;
  const uint64_t *bin_first;   // [n_chrom + 1] first output bin of every chromosome
  const uint32_t *chrom_len;   // [n_chrom]
  uint64_t n_bins;
  float *bins;                 // [n_bins]
  // weighted median: pairs per bin (counts -> exclusive offsets after the scan) and their storage
  uint64_t *pair_off;
  float *pair_val;
  uint32_t *pair_w;
  // bins -> items
  uint64_t *tile_count;
  uint32_t n_tiles;
  BwItem *out_items;
  uint64_t *chrom_item_first;  // [n_chrom]
};

// ---- BAM writers: split / split-exogenous / modify (bam_write.cuh) -------------------------------------------------------------
constexpr uint32_t BGZF_RAW = 0xff00;    // payload bytes per BGZF block (htslib's block size)
constexpr uint32_t BGZF_SLOT = 65536;    // bytes reserved per compressed block (a BGZF block holds at most 64 KiB)
constexpr int BAMW_MAX_CLASSES = 4;
constexpr int BAMW_TILE = 256;           // records per tile of the class scans
enum : int { BAMW_SPLIT = 0, BAMW_MODIFY = 1, BAMW_SPLIT_EXOGENOUS = 2 };
// counters of a writer job (SplitStats of spike_in_analysis.rs:28-58 come first, in its order)
enum : int {
  BW_N_TOTAL = 0, BW_N_UNMAPPED, BW_N_QCFAIL, BW_N_DUPLICATE, BW_N_SECONDARY, BW_N_FILTERED, BW_N_LOW_MAPQ, BW_N_BOTH, BW_N_EXOGENOUS, BW_N_ENDOGENOUS,
  BW_N_WRITTEN, BW_N_DROPPED_BY_SHIFT, BW_N_ERRORS, BW_N_COUNTERS = 16
};

struct BamwArgs {
  const uint8_t *buf;         // segment buffer (records of the segment through rec_off)
  const uint64_t *rec_off;
  const SegState *st;
  FilterDev f;
  int32_t mode, tn5_shift, n_classes;
  int32_t n_ref;
  const uint64_t *ref_len;    // [n_ref]
  const uint8_t *ref_is_exo;  // [n_ref] split-exogenous: reference name starts with the exogenous prefix
  uint8_t *rec_cls;           // [records] class of every record of the segment (0xff = not written)
  uint64_t *tile_sum;         // [tiles][BAMW_MAX_CLASSES] bytes of every class per tile -> exclusive stream offsets after the scan
  uint64_t *class_len;        // [BAMW_MAX_CLASSES] bytes appended to every class stream so far (chained from segment to segment)
  unsigned long long *counts; // [BW_N_COUNTERS]
  uint8_t *stream[BAMW_MAX_CLASSES];      // class streams
  uint32_t *first_rec[BAMW_MAX_CLASSES];  // per BGZF_RAW block of a stream: offset of the first record starting in it
};

struct BamwDeflateArgs {
  const uint8_t *raw;         // stream to cut into blocks of BGZF_RAW bytes (16-byte aligned)
  uint64_t raw_len;
  uint32_t n_blocks;          // ceil(raw_len / BGZF_RAW)
  const uint32_t *first_rec;  // [n_blocks] or nullptr (no record structure: header text)
  uint8_t *out;               // [n_blocks][BGZF_SLOT]
  uint32_t *out_size;         // [n_blocks]
  uint32_t *tokens;           // scratch [gridDim.x][BGZF_RAW]
  const CrcTables *crc;
  int32_t store_only;         // 1 = stored blocks (no compression)
};

}  // namespace bnd
