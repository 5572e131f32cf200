// outputs.cpp — see outputs.h: BigWig encoding (host, north_star keeps it there), the multi-BAM TSV, the
// collapse-bedgraph file front end and the CLI mirror of src/cli.rs for the three coverage subcommands.
#include "outputs.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>
#include <thread>

namespace bnd {

namespace {

std::string lower(std::string s) {
  for (auto &c : s) c = (char)tolower((unsigned char)c);
  return s;
}

}  // namespace

// ---- to_bigwig (coverage_analysis.rs:642-759): items, sections, zlib streams and zoom summaries are built on the device
// (bigwig.cuh, Pileup::to_bigwig in engine.cpp); bigwig_layout.h holds the host-side file layout (B+ tree, R-tree).
void pileup_to_bigwig(Pileup &p, const std::string &path) { p.to_bigwig(path); }

// ---- MultiBamPileup::to_tsv (coverage_analysis.rs:783-864), intended semantics ------------------------------------
void multi_to_tsv(const bnd_pileup_cfg *cfgs, const bnd_filter_cfg *filters, int n, const char *out) {
  if (n < 1) fail("At least one BAM pileup is required", ERR_INVALID);
  std::vector<std::unique_ptr<Pileup>> owned;
  std::vector<Pileup *> bams;
  for (int i = 0; i < n; i++) {
    owned.emplace_back(new Pileup(&cfgs[i], filters ? &filters[i] : nullptr));
    bams.push_back(owned.back().get());
  }
  Pileup::multi_to_tsv(bams, out);
}

// ---- collapse-bedgraph (src/cli.rs:1149-1205) -----------------------------------------------------------------------------------
// The input's bytes are handed to the device as they are (Pileup engine: collapse_bedgraph_text); what comes back are the
// collapsed groups.  Their text (name, start, end, shortest round-trip score like polars' CsvWriter) is formatted here by all
// host cores at once and written with one positional write per thread.
namespace {
void write_collapsed(const std::vector<std::string> &names, const BdgRows &res, int fd, bool seekable) {
  const size_t m = res.chrom.size();
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const int threads = (int)std::max<size_t>(1, std::min<size_t>(hw, m / 65536 + 1));
  std::vector<std::string> parts((size_t)threads);
  std::vector<std::thread> th;
  for (int t = 0; t < threads; t++)
    th.emplace_back([&, t] {
      const size_t lo = m * (size_t)t / (size_t)threads, hi = m * ((size_t)t + 1) / (size_t)threads;
      std::string &text = parts[(size_t)t];
      text.reserve((hi - lo) * 32);
      char buf[48];
      for (size_t i = lo; i < hi; i++) {
        text += names[res.chrom[i]];
        text.push_back('\t');
        text.append(buf, (size_t)snprintf(buf, sizeof buf, "%lld", res.start[i]));
        text.push_back('\t');
        text.append(buf, (size_t)snprintf(buf, sizeof buf, "%lld", res.end[i]));
        text.push_back('\t');
        text += format_f64(res.score[i]);
        text.push_back('\n');
      }
    });
  for (auto &t : th) t.join();
  auto write_all = [&](const std::string &text, off_t at) {
    size_t done = 0;
    while (done < text.size()) {
      const ssize_t w = seekable ? pwrite(fd, text.data() + done, text.size() - done, at + (off_t)done) : write(fd, text.data() + done, text.size() - done);
      if (w <= 0) fail("write error on the collapsed bedGraph");
      done += (size_t)w;
    }
  };
  off_t at = 0;
  for (auto &p : parts) {
    write_all(p, at);
    at += (off_t)p.size();
  }
}
}  // namespace

void collapse_bedgraph_file(const char *in_path, const char *out_path) {
  std::vector<std::string> names;
  BdgRows res;
  if (in_path) {
    const int fd = ::open(in_path, O_RDONLY);
    if (fd < 0) fail(std::string("cannot open ") + in_path);
    struct stat st;
    if (fstat(fd, &st) != 0) {
      close(fd);
      fail(std::string("cannot stat ") + in_path);
    }
    const uint64_t n = (uint64_t)st.st_size;
    void *m = n && S_ISREG(st.st_mode) ? mmap(nullptr, n, PROT_READ, MAP_SHARED, fd, 0) : MAP_FAILED;
    try {
      if (m != MAP_FAILED) {
        collapse_bedgraph_text((const uint8_t *)m, n, names, res, -1);
      } else {  // pipes, character devices, empty files
        std::string all;
        char buf[1 << 16];
        ssize_t r;
        while ((r = read(fd, buf, sizeof buf)) > 0) all.append(buf, (size_t)r);
        if (r < 0) fail(std::string("read error on ") + in_path);
        collapse_bedgraph_text((const uint8_t *)all.data(), all.size(), names, res, -1);
      }
    } catch (...) {
      if (m != MAP_FAILED) munmap(m, n);
      close(fd);
      throw;
    }
    if (m != MAP_FAILED) munmap(m, n);
    close(fd);
  } else {
    std::string all;
    char buf[1 << 16];
    size_t r;
    while ((r = fread(buf, 1, sizeof buf, stdin)) > 0) all.append(buf, r);
    collapse_bedgraph_text((const uint8_t *)all.data(), all.size(), names, res, -1);
  }
  if (out_path) {
    const int fd = ::open(out_path, O_WRONLY | O_CREAT | O_TRUNC, 0644);
    if (fd < 0) fail(std::string("cannot create ") + out_path);
    try {
      write_collapsed(names, res, fd, true);
    } catch (...) {
      close(fd);
      throw;
    }
    if (close(fd) != 0) fail(std::string("write error on ") + out_path);
  } else {
    fflush(stdout);
    write_collapsed(names, res, 1, false);
  }
}

// ---- run_cli (src/cli.rs:737-757) for bam-coverage / multi-bam-coverage / collapse-bedgraph -------------------------------------
namespace {

const char *VERSION = "bamnado 0.9.0";

struct CliError {
  int code;
  std::string msg;
};
[[noreturn]] void usage_error(const std::string &m) { throw CliError{2, "error: " + m + "\n\nFor more information, try '--help'.\n"}; }

const char *TOP_HELP =
    "Fast BAM and BigWig utilities for genomics workflows.\n\n"
    "Usage: bamnado [OPTIONS] <COMMAND>\n\n"
    "Commands (B200 coverage path):\n"
    "  bam-coverage        Generate a coverage track from one BAM file [aliases: coverage]\n"
    "  multi-bam-coverage  Merge coverage from multiple BAM files into one track [aliases: multi-coverage]\n"
    "  split               Split a BAM file using the supplied filters\n"
    "  split-exogenous     Split a BAM file into endogenous and exogenous reads [aliases: split-spikein]\n"
    "  modify              Filter and/or adjust reads in a BAM file\n"
    "  bigwig-compare      Compare two BigWig files bin by bin [aliases: compare-bigwigs]\n"
    "  bigwig-aggregate    Aggregate multiple BigWig files into one track [aliases: aggregate-bigwigs]\n"
    "  collapse-bedgraph   Collapse adjacent bins with equal scores in a bedGraph file [aliases: collapse]\n\n"
    "Options:\n"
    "  -v, --verbose <LEVEL>  Log verbosity from 0 (quiet) to 4 (debug) [default: 2]\n"
    "  -h, --help             Print help\n"
    "  -V, --version          Print version\n";

const char *COVERAGE_HELP =
    "Options:\n"
    "  -b, --bam <BAM> / --bams <BAM>...   Input BAM file(s)\n"
    "  -o, --output <FILE>                 Output bedGraph, BigWig or TSV path\n\n"
    "Coverage Options:\n"
    "  -s, --bin-size <BP>                 Bin size for the output track\n"
    "  -n, --normalize <METHOD>            raw | rpkm | cpm [default: raw] [aliases: --norm-method]\n"
    "  -f, --scale-factor <FLOAT>          Multiply the final signal by this factor\n"
    "      --fragment-counts               Count fragments instead of reads [aliases: --use-fragment]\n"
    "      --shift <L,R,FL,FR>             Shift read or fragment ends before pileup [default: 0,0,0,0]\n"
    "      --truncate <L,R,FL,FR>          Trim read or fragment ends before pileup\n"
    "      --ignore-scaffolds              Skip scaffold or unplaced chromosomes [aliases: --ignore-scaffold]\n"
    "      --threads <N>                   Threads to use while writing BigWig output [default: 6]\n\n"
    "Read Filters:\n"
    "      --strand <STRAND>               both | forward | reverse [default: both]\n"
    "      --proper-pairs                  Keep only properly paired reads [aliases: --proper-pair]\n"
    "      --min-mapq <INT>                [default: 20]\n"
    "      --min-length <BP>               [default: 20]\n"
    "      --max-length <BP>               [default: 1000]\n"
    "      --blacklist <BED>               [aliases: --blacklisted-locations]\n"
    "      --barcode-allowlist <FILE>      [aliases: --whitelisted-barcodes]\n"
    "      --read-group <RG>\n"
    "      --tag <TAG>                     [aliases: --filter-tag]\n"
    "      --tag-value <VALUE>             [aliases: --filter-tag-value]\n"
    "      --min-fragment-len <BP>         [aliases: --min-fragment-length]\n"
    "      --max-fragment-len <BP>         [aliases: --max-fragment-length]\n"
    "      --ignore-duplicates             [aliases: --no-duplicates]\n"
    "      --ignore-secondary\n"
    "      --ignore-supplementary\n"
    "  -h, --help                          Print help\n";

struct Opts {
  std::vector<std::string> bams;
  bool has_output = false;
  std::string output;
  bool has_input = false;
  std::string input;
  bool has_bin = false;
  uint64_t bin = 50;
  int norm = 0;
  float scale = 1.0f;
  bool use_fragment = false, ignore_scaffold = false;
  bool has_shift = true;
  int32_t shift[4] = {0, 0, 0, 0};
  bool has_truncate = false, t_has5 = false, t_has3 = false;
  uint64_t t5 = 0, t3 = 0;
  int threads = 6;
  int strand = 0;
  bool proper_pair = false;
  uint32_t min_mapq = 20, min_length = 20, max_length = 1000;
  bool has_blacklist = false, has_barcodes = false, has_rg = false, has_tag = false, has_tag_value = false;
  std::string blacklist, barcodes, rg, tag, tag_value;
  bool has_min_frag = false, has_max_frag = false;
  uint32_t min_frag = 0, max_frag = 0;
  bool ex_dup = false, ex_sec = false, ex_sup = false;
};

uint64_t parse_uint(const std::string &opt, const std::string &v, uint64_t max) {
  if (v.empty()) usage_error("invalid value '" + v + "' for '" + opt + "'");
  uint64_t x = 0;
  for (char c : v) {
    if (c < '0' || c > '9') usage_error("invalid value '" + v + "' for '" + opt + "': invalid digit found in string");
    x = x * 10 + (uint64_t)(c - '0');
    if (x > max) usage_error("invalid value '" + v + "' for '" + opt + "': number too large to fit in target type");
  }
  return x;
}

std::vector<std::string> split_commas(const std::string &s) {
  std::vector<std::string> out;
  std::string cur;
  for (char c : s) {
    if (c == ',') out.push_back(cur), cur.clear();
    else cur.push_back(c);
  }
  out.push_back(cur);
  return out;
}

bool parse_i32(const std::string &s, int32_t &v) {
  if (s.empty()) return false;
  char *e = nullptr;
  errno = 0;
  long long x = strtoll(s.c_str(), &e, 10);
  if (*e || errno || x < INT32_MIN || x > INT32_MAX) return false;
  v = (int32_t)x;
  return true;
}

// Parses the flags shared by the coverage subcommands.  `multi`: -b/--bams collects a list.
void parse_coverage_args(const std::vector<std::string> &args, bool multi, Opts &o, bool &help) {
  for (size_t i = 0; i < args.size(); i++) {
    std::string a = args[i], inline_val;
    bool has_inline = false;
    if (a.rfind("--", 0) == 0) {
      size_t eq = a.find('=');
      if (eq != std::string::npos) {
        inline_val = a.substr(eq + 1);
        a = a.substr(0, eq);
        has_inline = true;
      }
    } else if (a.size() > 2 && a[0] == '-' && a[1] != '-') {  // -s50
      inline_val = a.substr(2);
      if (!inline_val.empty() && inline_val[0] == '=') inline_val = inline_val.substr(1);
      a = a.substr(0, 2);
      has_inline = true;
    }
    auto value = [&]() -> std::string {
      if (has_inline) return inline_val;
      if (i + 1 >= args.size()) usage_error("a value is required for '" + a + "' but none was supplied");
      return args[++i];
    };
    auto flag = [&]() {
      if (has_inline) usage_error("unexpected value '" + inline_val + "' for '" + a + "' found; no more were expected");
    };
    if (a == "-h" || a == "--help") {
      help = true;
      return;
    } else if (a == "-v" || a == "--verbose") {
      parse_uint(a, value(), 255);
    } else if ((!multi && (a == "-b" || a == "--bam")) || (multi && (a == "-b" || a == "--bams"))) {
      o.bams.push_back(value());
      if (multi)
        while (!has_inline && i + 1 < args.size() && !args[i + 1].empty() && args[i + 1][0] != '-') o.bams.push_back(args[++i]);
    } else if (a == "-o" || a == "--output") {
      o.output = value();
      o.has_output = true;
    } else if (a == "-s" || a == "--bin-size") {
      o.bin = parse_uint(a, value(), UINT64_MAX / 16);
      o.has_bin = true;
    } else if (a == "-n" || a == "--normalize" || a == "--norm-method") {
      std::string v = lower(value());
      if (v == "raw") o.norm = 0;
      else if (v == "rpkm") o.norm = 1;
      else if (v == "cpm") o.norm = 2;
      else usage_error("invalid value '" + v + "' for '--normalize <METHOD>'\n  [possible values: raw, rpkm, cpm]");
    } else if (a == "-f" || a == "--scale-factor") {
      std::string v = value();
      char *e = nullptr;
      o.scale = strtof(v.c_str(), &e);
      if (v.empty() || *e) usage_error("invalid value '" + v + "' for '--scale-factor <FLOAT>': invalid float literal");
    } else if (a == "--fragment-counts" || a == "--use-fragment") {
      flag();
      o.use_fragment = true;
    } else if (a == "--shift") {
      auto parts = split_commas(value());
      if (parts.size() != 4) usage_error("invalid value for '--shift <L,R,FL,FR>': Shift must be in the format '5,3,5r,3r'");
      for (int k = 0; k < 4; k++)
        if (!parse_i32(parts[k], o.shift[k])) o.shift[k] = 0;  // unparseable parts silently become 0
      o.has_shift = true;
    } else if (a == "--truncate") {
      auto parts = split_commas(value());
      if (parts.size() != 2) usage_error("invalid value for '--truncate <L,R,FL,FR>': Truncate must be in the format '5,3'");
      o.has_truncate = true;
      auto parse_usize = [](const std::string &s, uint64_t &v) {
        if (s.empty()) return false;
        v = 0;
        for (char c : s) {
          if (c < '0' || c > '9') return false;
          v = v * 10 + (uint64_t)(c - '0');
        }
        return true;
      };
      o.t_has5 = parse_usize(parts[0], o.t5);
      o.t_has3 = parse_usize(parts[1], o.t3);
    } else if (a == "--ignore-scaffolds" || a == "--ignore-scaffold") {
      flag();
      o.ignore_scaffold = true;
    } else if (a == "--threads") {
      o.threads = (int)parse_uint(a, value(), 1u << 20);
    } else if (a == "--strand") {
      std::string v = lower(value());
      if (v == "both") o.strand = 0;
      else if (v == "forward") o.strand = 1;
      else if (v == "reverse") o.strand = 2;
      else usage_error("invalid value '" + v + "' for '--strand <STRAND>': Invalid strand value: " + v);
    } else if (a == "--proper-pairs" || a == "--proper-pair") {
      flag();
      o.proper_pair = true;
    } else if (a == "--min-mapq") {
      o.min_mapq = (uint32_t)parse_uint(a, value(), 255);
    } else if (a == "--min-length") {
      o.min_length = (uint32_t)parse_uint(a, value(), UINT32_MAX);
    } else if (a == "--max-length") {
      o.max_length = (uint32_t)parse_uint(a, value(), UINT32_MAX);
    } else if (a == "--blacklist" || a == "--blacklisted-locations") {
      o.blacklist = value();
      o.has_blacklist = true;
    } else if (a == "--barcode-allowlist" || a == "--whitelisted-barcodes") {
      o.barcodes = value();
      o.has_barcodes = true;
    } else if (a == "--read-group") {
      o.rg = value();
      o.has_rg = true;
    } else if (a == "--tag" || a == "--filter-tag") {
      o.tag = value();
      o.has_tag = true;
    } else if (a == "--tag-value" || a == "--filter-tag-value") {
      o.tag_value = value();
      o.has_tag_value = true;
    } else if (a == "--min-fragment-len" || a == "--min-fragment-length") {
      o.min_frag = (uint32_t)parse_uint(a, value(), UINT32_MAX);
      o.has_min_frag = true;
    } else if (a == "--max-fragment-len" || a == "--max-fragment-length") {
      o.max_frag = (uint32_t)parse_uint(a, value(), UINT32_MAX);
      o.has_max_frag = true;
    } else if (a == "--ignore-duplicates" || a == "--no-duplicates") {
      flag();
      o.ex_dup = true;
    } else if (a == "--ignore-secondary") {
      flag();
      o.ex_sec = true;
    } else if (a == "--ignore-supplementary") {
      flag();
      o.ex_sup = true;
    } else {
      usage_error("unexpected argument '" + args[i] + "' found");
    }
  }
}

void fill_cfgs(const Opts &o, const std::string &bam, bool collapse, bool merge_blacklist, int barcode_column, bnd_pileup_cfg &pc, bnd_filter_cfg &fc) {
  memset(&pc, 0, sizeof pc);
  memset(&fc, 0, sizeof fc);
  pc.bam_path = bam.c_str();
  pc.bin_size = o.bin;
  pc.norm_method = o.norm;
  pc.scale_factor = o.scale;
  pc.use_fragment = o.use_fragment;
  pc.collapse = collapse;
  pc.ignore_scaffold = o.ignore_scaffold;
  pc.has_shift = o.has_shift;
  for (int k = 0; k < 4; k++) pc.shift[k] = o.shift[k];
  pc.has_truncate = o.has_truncate;
  pc.trunc_has5 = o.t_has5, pc.trunc5 = o.t5, pc.trunc_has3 = o.t_has3, pc.trunc3 = o.t3;
  pc.n_threads = o.threads;
  pc.device = -1;
  fc.strand = o.strand;
  fc.proper_pair = o.proper_pair;
  fc.min_mapq = o.min_mapq, fc.min_length = o.min_length, fc.max_length = o.max_length;
  fc.has_min_fragment_length = o.has_min_frag, fc.min_fragment_length = o.min_frag;
  fc.has_max_fragment_length = o.has_max_frag, fc.max_fragment_length = o.max_frag;
  fc.exclude_duplicates = o.ex_dup, fc.exclude_secondary = o.ex_sec, fc.exclude_supplementary = o.ex_sup;
  fc.read_group = o.has_rg ? o.rg.c_str() : nullptr;
  fc.filter_tag = o.has_tag ? o.tag.c_str() : nullptr;
  fc.filter_tag_value = o.has_tag_value ? o.tag_value.c_str() : nullptr;
  fc.blacklist_bed = o.has_blacklist ? o.blacklist.c_str() : nullptr;
  fc.blacklist_merge_overlaps = merge_blacklist;
  fc.barcode_csv = o.has_barcodes ? o.barcodes.c_str() : nullptr;
  fc.barcode_csv_column = barcode_column;
}

bool file_exists(const std::string &p) {
  std::ifstream f(p);
  return (bool)f;
}

void validate_bam_file(const std::string &bam) {  // src/cli.rs:553-570
  if (!file_exists(bam)) fail("BAM file does not exist: " + bam);
  std::string idx = bam;
  size_t dot = idx.find_last_of('.');
  size_t slash = idx.find_last_of('/');
  if (dot != std::string::npos && (slash == std::string::npos || dot > slash)) idx = idx.substr(0, dot);
  idx += ".bam.bai";  // Path::with_extension("bam.bai")
  if (!file_exists(idx))
    fail("BAM index file does not exist for " + bam + ". Please create the index file using samtools index command");
}

int file_type(const std::string &path) {  // 0 bedgraph, 1 bigwig, 2 tsv (bam_utils.rs:595-609)
  size_t dot = path.find_last_of('.');
  size_t slash = path.find_last_of('/');
  if (dot == std::string::npos || (slash != std::string::npos && dot < slash)) fail("No file extension found");
  std::string ext = lower(path.substr(dot + 1));
  if (ext == "bedgraph" || ext == "bdg") return 0;
  if (ext == "bigwig" || ext == "bw") return 1;
  if (ext == "tsv" || ext == "txt") return 2;
  fail("Unsupported file extension: " + ext + "\n\nCaused by:\n    Unknown file type: " + ext);
}

int run(const std::vector<std::string> &argv) {
  size_t i = 1;
  // global options before the subcommand
  while (i < argv.size() && !argv[i].empty() && argv[i][0] == '-') {
    const std::string &a = argv[i];
    if (a == "-h" || a == "--help") {
      fputs(TOP_HELP, stdout);
      return 0;
    }
    if (a == "-V" || a == "--version") {
      printf("%s\n", VERSION);
      return 0;
    }
    if (a == "-v" || a == "--verbose") {
      if (i + 1 >= argv.size()) usage_error("a value is required for '--verbose <LEVEL>' but none was supplied");
      parse_uint(a, argv[i + 1], 255);
      i += 2;
      continue;
    }
    if (a.rfind("--verbose=", 0) == 0 || (a.rfind("-v", 0) == 0 && a.size() > 2)) {
      i++;
      continue;
    }
    usage_error("unexpected argument '" + a + "' found");
  }
  if (i >= argv.size()) {  // arg_required_else_help
    fputs(TOP_HELP, stderr);
    return 2;
  }
  const std::string cmd = argv[i++];
  std::vector<std::string> rest(argv.begin() + (long)i, argv.end());
  if (cmd == "bam-coverage" || cmd == "coverage") {
    Opts o;
    bool help = false;
    parse_coverage_args(rest, false, o, help);
    if (help) {
      printf("Generate a coverage track from one BAM file\n\nUsage: bamnado bam-coverage [OPTIONS] --bam <BAM>\n\n%s\nExample:\n  bamnado bam-coverage --bam sample.bam --output sample.bw --bin-size 50 --normalize cpm\n", COVERAGE_HELP);
      return 0;
    }
    if (o.bams.empty()) usage_error("the following required arguments were not provided:\n  --bam <BAM>");
    if (o.bams.size() > 1) usage_error("the argument '--bam <BAM>' cannot be used multiple times");
    const std::string &bam = o.bams[0];
    validate_bam_file(bam);
    bnd_pileup_cfg pc;
    bnd_filter_cfg fc;
    fill_cfgs(o, bam, true, true, -1, pc, fc);
    Pileup pile(&pc, &fc);
    if ((o.has_min_frag || o.has_max_frag) && !pile.is_paired_end())
      fail("Fragment length filtering (--min-fragment-length / --max-fragment-length) requires paired-end reads, but the BAM file does not appear to contain paired-end data.");
    std::string outfile = o.output;
    if (!o.has_output) {
      outfile = bam;
      size_t dot = outfile.find_last_of('.');
      size_t slash = outfile.find_last_of('/');
      if (dot != std::string::npos && (slash == std::string::npos || dot > slash)) outfile = outfile.substr(0, dot);
      outfile += ".bedgraph";
    }
    switch (file_type(outfile)) {
      case 0: pile.to_bedgraph(outfile); break;
      case 1: pileup_to_bigwig(pile, outfile); break;
      default: fail("TSV output is not supported for single BAM coverage");
    }
    return 0;
  }
  if (cmd == "multi-bam-coverage" || cmd == "multi-coverage") {
    Opts o;
    bool help = false;
    parse_coverage_args(rest, true, o, help);
    if (help) {
      printf("Merge coverage from multiple BAM files into one track\n\nUsage: bamnado multi-bam-coverage [OPTIONS]\n\n%s", COVERAGE_HELP);
      return 0;
    }
    for (auto &b : o.bams) validate_bam_file(b);
    std::string outfile = o.has_output ? o.output : "output.tsv";
    if (file_type(outfile) != 2) fail("Unsupported output format. Currently only TSV is supported for multi-BAM coverage");
    std::vector<bnd_pileup_cfg> pcs(o.bams.size());
    std::vector<bnd_filter_cfg> fcs(o.bams.size());
    for (size_t k = 0; k < o.bams.size(); k++) fill_cfgs(o, o.bams[k], false, false, o.has_barcodes ? (int)k : -1, pcs[k], fcs[k]);
    multi_to_tsv(pcs.data(), fcs.data(), (int)o.bams.size(), outfile.c_str());
    return 0;
  }
  if (cmd == "collapse-bedgraph" || cmd == "collapse") {
    std::string in, out;
    bool has_in = false, has_out = false;
    for (size_t k = 0; k < rest.size(); k++) {
      const std::string &a = rest[k];
      auto val = [&]() -> std::string {
        if (k + 1 >= rest.size()) usage_error("a value is required for '" + a + "' but none was supplied");
        return rest[++k];
      };
      if (a == "-h" || a == "--help") {
        printf("Collapse adjacent bins with equal scores in a bedGraph file\n\nUsage: bamnado collapse-bedgraph [OPTIONS]\n\nOptions:\n  -i, --input <BEDGRAPH>   Input bedGraph path. Reads from stdin if omitted\n  -o, --output <BEDGRAPH>  Output bedGraph path. Writes to stdout if omitted\n  -h, --help               Print help\n");
        return 0;
      } else if (a == "-i" || a == "--input") in = val(), has_in = true;
      else if (a == "-o" || a == "--output") out = val(), has_out = true;
      else if (a.rfind("--input=", 0) == 0) in = a.substr(8), has_in = true;
      else if (a.rfind("--output=", 0) == 0) out = a.substr(9), has_out = true;
      else if (a == "-v" || a == "--verbose") val();
      else usage_error("unexpected argument '" + a + "' found");
    }
    collapse_bedgraph_file(has_in ? in.c_str() : nullptr, has_out ? out.c_str() : nullptr);
    return 0;
  }
  if (cmd == "split" || cmd == "modify" || cmd == "split-exogenous" || cmd == "split-spikein") {
    // src/cli.rs:343-395 (flags) and :962-1079 (dispatch): --input / --output plus the read filters; the pass and the BAM writer
    // run on the device (engine: Pileup::write_bam)
    const bool exo = cmd == "split-exogenous" || cmd == "split-spikein", modify = cmd == "modify";
    std::vector<std::string> pass;
    std::string input, prefix, stats;
    bool has_input = false, has_prefix = false, has_stats = false, tn5 = false;
    for (size_t k = 0; k < rest.size(); k++) {
      std::string a = rest[k], inl;
      bool has_inl = false;
      if (a.rfind("--", 0) == 0 && a.find('=') != std::string::npos) inl = a.substr(a.find('=') + 1), a = a.substr(0, a.find('=')), has_inl = true;
      else if (a.size() > 2 && a[0] == '-' && a[1] != '-') inl = a.substr(a[2] == '=' ? 3 : 2), a = a.substr(0, 2), has_inl = true;
      auto val = [&]() -> std::string {
        if (has_inl) return inl;
        if (k + 1 >= rest.size()) usage_error("a value is required for '" + a + "' but none was supplied");
        return rest[++k];
      };
      if (a == "-i" || a == "--input") input = val(), has_input = true;
      else if (exo && (a == "-e" || a == "--exogenous-prefix")) prefix = val(), has_prefix = true;
      else if (exo && (a == "-s" || a == "--stats")) stats = val(), has_stats = true;
      else if (modify && a == "--tn5-shift") tn5 = true;
      else pass.push_back(rest[k]);
    }
    Opts o;
    bool help = false;
    parse_coverage_args(pass, false, o, help);
    if (help) {
      printf("%s\n\nUsage: bamnado %s [OPTIONS] --input <BAM> --output <PREFIX>%s\n\nOptions:\n  -i, --input <BAM>                   Input BAM file\n"
             "  -o, --output <PREFIX>               Output prefix\n%s%s\nRead Filters: the same flags as bam-coverage (--strand, --proper-pairs, --min-mapq, --min-length, --max-length,\n"
             "  --blacklist, --barcode-allowlist, --read-group, --tag, --tag-value, --min-fragment-len, --max-fragment-len,\n  --ignore-duplicates, --ignore-secondary, --ignore-supplementary)\n",
             exo ? "Split a BAM file into endogenous and exogenous reads" : modify ? "Filter and/or adjust reads in a BAM file" : "Split a BAM file using the supplied filters",
             cmd.c_str(), exo ? " --exogenous-prefix <PREFIX>" : "",
             exo ? "  -e, --exogenous-prefix <PREFIX>     Reference-name prefix used to identify exogenous sequences\n  -s, --stats <FILE>                  Optional path for summary statistics output\n" : "",
             modify ? "      --tn5-shift                     Apply the standard Tn5 offset\n" : "");
      return 0;
    }
    if (o.has_bin || !o.bams.empty()) usage_error("unexpected argument found");
    std::string missing;
    if (!has_input) missing += "\n  --input <BAM>";
    if (!o.has_output) missing += "\n  --output <PREFIX>";
    if (exo && !has_prefix) missing += "\n  --exogenous-prefix <PREFIX>";
    if (!missing.empty()) usage_error("the following required arguments were not provided:" + missing);
    validate_bam_file(input);
    bnd_pileup_cfg pc;
    bnd_filter_cfg fc;
    fill_cfgs(o, input, true, true, -1, pc, fc);
    std::string cl;  // bamnado_command_line (bam_utils.rs:782-788): the process arguments joined by spaces
    for (size_t k = 0; k < argv.size(); k++) cl += (k ? " " : "") + argv[k];
    uint64_t cnt[BW_N_COUNTERS] = {0};
    if (exo) {
      std::vector<std::string> paths;
      for (const char *ext : {"endogenous.bam", "exogenous.bam", "both.bam", "unmapped.bam"}) paths.push_back(path_with_extension(o.output, ext));
      bam_write_tool(input.c_str(), &fc, BAMW_SPLIT_EXOGENOUS, 0, prefix.c_str(), paths, cl.c_str(), -1, cnt, nullptr);
      fprintf(stderr, "Successfully split BAM file\n");
      if (has_stats) {  // serde_json of SplitStats (spike_in_analysis.rs:101-125); non-finite floats are null
        const size_t slash = input.find_last_of('/');
        const std::string fname = slash == std::string::npos ? input : input.substr(slash + 1);
        auto f64 = [](double v) { return std::isfinite(v) ? format_f64(v) : std::string("null"); };
        double factor = 1.0 / ((double)cnt[BW_N_EXOGENOUS] / 1e6);
        if (std::isinf(factor)) factor = 1.0;
        const double pct = (double)cnt[BW_N_EXOGENOUS] / (double)cnt[BW_N_ENDOGENOUS] * 100.0;
        std::string js = "{\"filename\":\"";
        for (char c : fname) {
          if (c == '"' || c == '\\') js += '\\';
          js += c;
        }
        js += "\"";
        const char *keys[10] = {"n_total_reads", "n_unmapped_reads", "n_qcfail_reads", "n_duplicate_reads", "n_secondary_reads", "n_filtered_reads", "n_low_maq",
                                "n_both_genomes", "n_exogenous", "n_endogenous"};
        for (int k = 0; k < 10; k++) js += std::string(",\"") + keys[k] + "\":" + std::to_string(cnt[k]);
        js += ",\"spike_in_factor\":" + f64(factor) + ",\"spike_in_percentage\":" + f64(pct) + "}";
        std::ofstream f(stats, std::ios::binary);
        if (!f) fail("Failed to create stats file");
        f << js;
        if (!f.flush()) fail("Failed to write stats file");
        fprintf(stderr, "Successfully wrote stats file to %s\n", stats.c_str());
      }
    } else {
      bam_write_tool(input.c_str(), &fc, modify ? BAMW_MODIFY : BAMW_SPLIT, tn5 ? 1 : 0, nullptr, {o.output}, cl.c_str(), -1, cnt, nullptr);
      fprintf(stderr, modify ? "Successfully modified BAM file\n" : "Successfully split BAM file\n");
    }
    return 0;
  }
  if (cmd == "bigwig-compare" || cmd == "compare-bigwigs" || cmd == "bigwig-aggregate" || cmd == "aggregate-bigwigs") {
    // src/cli.rs:259-330 (flags) and :1081-1147 (dispatch)
    const bool compare = cmd == "bigwig-compare" || cmd == "compare-bigwigs";
    std::string bw1, bw2, out, method;
    std::vector<std::string> inputs;
    std::vector<double> sfs;
    uint32_t bin_size = 50;
    bool has_pc = false, has_sf1 = false, has_sf2 = false;
    double pc = 0, sf1 = 1, sf2 = 1;
    auto parse_f64 = [&](const std::string &v, const std::string &opt) {
      char *e = nullptr;
      const double x = strtod(v.c_str(), &e);
      if (v.empty() || *e) usage_error("invalid value '" + v + "' for '" + opt + " <FLOAT>': invalid float literal");
      return x;
    };
    for (size_t k = 0; k < rest.size(); k++) {
      const std::string &a = rest[k];
      auto val = [&]() -> std::string {
        if (k + 1 >= rest.size()) usage_error("a value is required for '" + a + "' but none was supplied");
        return rest[++k];
      };
      auto more = [&]() { return k + 1 < rest.size() && (rest[k + 1].empty() || rest[k + 1][0] != '-' || (rest[k + 1].size() > 1 && (isdigit((unsigned char)rest[k + 1][1]) || rest[k + 1][1] == '.'))); };
      if (a == "-h" || a == "--help") {
        if (compare)
          printf("Compare two BigWig files bin by bin\n\nUsage: bamnado bigwig-compare [OPTIONS] --bw1 <BW> --bw2 <BW> --output <BW> --comparison <METHOD>\n\nOptions:\n"
                 "      --bw1 <BW>                  First BigWig input\n      --bw2 <BW>                  Second BigWig input\n  -o, --output <BW>               Output BigWig path\n"
                 "  -c, --comparison <METHOD>       Comparison method [possible values: subtraction, ratio, log-ratio]\n  -s, --bin-size <BP>             Bin size for comparison [default: 50]\n"
                 "      --pseudocount <FLOAT>       Value added to both tracks before comparison\n      --scale-factor-bw1 <FLOAT>  Scale factor applied to bw1 before comparison\n"
                 "      --scale-factor-bw2 <FLOAT>  Scale factor applied to bw2 before comparison\n      --threads <THREADS>         Number of threads to use for BigWig writing [default: 6]\n  -h, --help                      Print help\n");
        else
          printf("Aggregate multiple BigWig files into one track\n\nUsage: bamnado bigwig-aggregate [OPTIONS] --output <BW> --method <METHOD>\n\nOptions:\n"
                 "      --bigwigs <BW>...           BigWig files to aggregate\n  -o, --output <BW>               Output BigWig path\n"
                 "  -m, --method <METHOD>           Aggregation method [possible values: sum, mean, median, max, min]\n  -s, --bin-size <BP>             Bin size for aggregation [default: 50]\n"
                 "      --pseudocount <FLOAT>       Value added to all inputs before aggregation\n      --scale-factors [<FLOAT>...]  Per-file scale factors (one per --bigwigs input, in order)\n"
                 "      --threads <THREADS>         Number of threads to use for BigWig writing [default: 6]\n  -h, --help                      Print help\n");
        return 0;
      } else if (compare && a == "--bw1") bw1 = val();
      else if (compare && a == "--bw2") bw2 = val();
      else if (!compare && a == "--bigwigs") {
        inputs.push_back(val());
        while (more()) inputs.push_back(rest[++k]);
      } else if (a == "-o" || a == "--output") out = val();
      else if ((compare && (a == "-c" || a == "--comparison")) || (!compare && (a == "-m" || a == "--method"))) method = lower(val());
      else if (a == "-s" || a == "--bin-size") bin_size = (uint32_t)parse_uint("--bin-size <BP>", val(), 0xffffffffull);
      else if (a == "--pseudocount") pc = parse_f64(val(), "--pseudocount"), has_pc = true;
      else if (compare && a == "--scale-factor-bw1") sf1 = parse_f64(val(), "--scale-factor-bw1"), has_sf1 = true;
      else if (compare && a == "--scale-factor-bw2") sf2 = parse_f64(val(), "--scale-factor-bw2"), has_sf2 = true;
      else if (!compare && a == "--scale-factors") {
        while (more()) sfs.push_back(parse_f64(rest[++k], "--scale-factors"));
      } else if (a == "--threads") val();  // sizes the reference's tokio writer; the device encoder has no use for it
      else if (a == "-v" || a == "--verbose") val();
      else usage_error("unexpected argument '" + a + "' found");
    }
    if (out.empty()) usage_error("the following required arguments were not provided:\n  --output <BW>");
    if (compare) {
      if (bw1.empty() || bw2.empty()) usage_error("the following required arguments were not provided:\n  --bw1 <BW>\n  --bw2 <BW>");
      int m = method == "subtraction" ? 0 : method == "ratio" ? 1 : method == "log-ratio" ? 2 : -1;
      if (m < 0) usage_error("invalid value '" + method + "' for '--comparison <METHOD>'\n  [possible values: subtraction, ratio, log-ratio]");
      bigwig_compare(bw1.c_str(), bw2.c_str(), out.c_str(), m, bin_size, has_pc ? &pc : nullptr, has_sf1 ? &sf1 : nullptr, has_sf2 ? &sf2 : nullptr, -1);
      fprintf(stderr, "Successfully compared BigWig files and wrote output to %s\n", out.c_str());
    } else {
      int m = method == "sum" ? 0 : method == "mean" ? 1 : method == "max" ? 2 : method == "min" ? 3 : method == "median" ? 4 : -1;
      if (m < 0) usage_error("invalid value '" + method + "' for '--method <METHOD>'\n  [possible values: sum, mean, median, max, min]");
      if (inputs.empty()) fail("At least one BigWig file must be provided");
      if (!sfs.empty() && sfs.size() != inputs.size())
        fail("--scale-factors count (" + std::to_string(sfs.size()) + ") must match input BigWig count (" + std::to_string(inputs.size()) + ")");
      std::vector<const char *> ps;
      for (auto &x : inputs) ps.push_back(x.c_str());
      bigwig_aggregate(ps.data(), (int)ps.size(), out.c_str(), 10 + m, bin_size, has_pc ? &pc : nullptr, sfs.empty() ? nullptr : sfs.data(), -1);
      fprintf(stderr, "Successfully aggregated %zu BigWig files and wrote output to %s\n", inputs.size(), out.c_str());
    }
    return 0;
  }
  usage_error("unrecognized subcommand '" + cmd + "'");
}

}  // namespace

int cli_main(int argc, const char *const *argv) {
  std::vector<std::string> args;
  for (int i = 0; i < argc; i++) args.emplace_back(argv[i] ? argv[i] : "");
  if (args.empty()) args.push_back("bamnado");
  try {
    return run(args);
  } catch (const CliError &e) {  // clap parse error: message + exit code 2
    fputs(e.msg.c_str(), stderr);
    return e.code;
  } catch (const std::exception &e) {  // anyhow error: "Error: ..." on stderr, exit code 1 (src/cli.rs:750-756)
    fprintf(stderr, "Error: %s\n", e.what());
    return 1;
  }
}

}  // namespace bnd
