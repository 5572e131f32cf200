// filter_host.cpp — see filter_host.h.
#include "filter_host.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <unordered_set>

namespace bnd {

namespace {

std::vector<std::string> read_text_lines(const std::string &path, const char *what) {
  std::ifstream in(path, std::ios::binary);
  if (!in) fail(std::string("Failed to read ") + what + ": " + path);
  std::vector<std::string> lines;
  std::string l;
  while (std::getline(in, l)) {
    if (!l.empty() && l.back() == '\r') l.pop_back();
    lines.push_back(l);
  }
  return lines;
}

std::vector<std::string> split_fields(const std::string &l) {
  std::vector<std::string> out;
  size_t i = 0;
  const size_t n = l.size();
  while (i < n) {
    while (i < n && (l[i] == ' ' || l[i] == '\t')) i++;
    size_t j = i;
    while (j < n && l[j] != ' ' && l[j] != '\t') j++;
    if (j > i) out.push_back(l.substr(i, j - i));
    i = j;
  }
  return out;
}

bool parse_u64(const std::string &s, uint64_t &v) {
  if (s.empty()) return false;
  v = 0;
  for (char c : s) {
    if (c < '0' || c > '9') return false;
    v = v * 10 + (uint64_t)(c - '0');
  }
  return true;
}

struct Span {
  uint64_t start, stop;
};

// rust-lapper merge_overlaps: intervals sorted by (start, stop); an interval is folded into the previous one unless
// the previous stop is strictly below its start
void merge_overlaps(std::vector<Span> &v) {
  std::vector<Span> out;
  for (const Span &iv : v) {
    if (out.empty() || out.back().stop < iv.start) out.push_back(iv);
    else if (out.back().stop < iv.stop) out.back().stop = iv.stop;
  }
  v.swap(out);
}

// bed_to_lapper + convert_lapper_chrom_names_to_ids: noodles BED feature_start = start + 1, feature_end = end; kept
// as half-open [start + 1, end)
void load_blacklist(FilterHost &F, const std::string &bed, const BamHeader &h, bool merge) {
  std::map<std::string, std::vector<Span>> by_chrom;
  for (const std::string &l : read_text_lines(bed, "blacklisted locations")) {
    if (l.empty() || l[0] == '#') continue;
    auto f = split_fields(l);
    if (f.empty() || f[0] == "track" || f[0] == "browser") continue;
    if (f.size() < 4) fail("Failed to read blacklisted locations: a BED record needs at least 4 fields");
    uint64_t s, e;
    if (!parse_u64(f[1], s) || !parse_u64(f[2], e)) fail("Failed to read blacklisted locations: invalid BED coordinates");
    if (e == 0) fail("Failed to get feature end");
    by_chrom[f[0]].push_back(Span{s + 1, e});
  }
  if (by_chrom.empty()) fail("Lapper is empty");
  const size_t n_ref = h.names.size();
  std::vector<std::vector<Span>> per_ref(n_ref);
  for (auto &kv : by_chrom) {
    auto it = h.name_to_id.find(kv.first);
    if (it == h.name_to_id.end()) fail("Chromosome " + kv.first + " not found in BAM file");
    auto &v = per_ref[it->second];
    v = kv.second;
    std::sort(v.begin(), v.end(), [](const Span &a, const Span &b) { return a.start != b.start ? a.start < b.start : a.stop < b.stop; });
    if (merge) merge_overlaps(v);
  }
  F.bl_off.assign(n_ref + 1, 0);
  for (size_t r = 0; r < n_ref; r++) {
    F.bl_off[r + 1] = F.bl_off[r] + per_ref[r].size();
    std::vector<uint64_t> st, en;
    for (const Span &s : per_ref[r]) st.push_back(s.start), en.push_back(s.stop);
    std::sort(st.begin(), st.end());
    std::sort(en.begin(), en.end());
    F.bl_starts.insert(F.bl_starts.end(), st.begin(), st.end());
    F.bl_stops.insert(F.bl_stops.end(), en.begin(), en.end());
  }
  F.has_blacklist = true;
}

std::vector<std::string> split_csv(const std::string &l) {
  std::vector<std::string> out;
  std::string cur;
  bool quoted = false;
  for (size_t i = 0; i < l.size(); i++) {
    const char c = l[i];
    if (quoted) {
      if (c == '"' && i + 1 < l.size() && l[i + 1] == '"') cur.push_back('"'), i++;
      else if (c == '"') quoted = false;
      else cur.push_back(c);
    } else if (c == '"') {
      quoted = true;
    } else if (c == ',') {
      out.push_back(cur);
      cur.clear();
    } else {
      cur.push_back(c);
    }
  }
  out.push_back(cur);
  return out;
}

// CellBarcodes::from_csv (column `barcode`, nulls are an error) / CellBarcodesMulti::from_csv (column by index, nulls
// dropped)
void load_barcode_csv(std::unordered_set<std::string> &set, const std::string &csv, int column) {
  auto lines = read_text_lines(csv, "barcodes");
  if (lines.empty()) fail("Failed to read barcodes: empty file " + csv);
  auto hdr = split_csv(lines[0]);
  int col = column;
  if (col < 0) {
    for (size_t i = 0; i < hdr.size(); i++)
      if (hdr[i] == "barcode") col = (int)i;
    if (col < 0) fail("Missing 'barcode' column in barcode CSV");
  } else if ((size_t)col >= hdr.size()) {
    fail("barcode CSV has too few columns");
  }
  for (size_t i = 1; i < lines.size(); i++) {
    if (lines[i].empty()) continue;
    auto f = split_csv(lines[i]);
    if ((size_t)col >= f.size() || f[col].empty()) {
      if (column < 0) fail("Null value in 'barcode' column");
      continue;
    }
    set.insert(f[col]);
  }
}

}  // namespace

FilterHost make_filter_host(const bnd_filter_cfg *c, const BamHeader &h) {
  FilterHost F;
  F.cfg = *c;
  if (c->read_group) F.has_rg = true, F.rg = c->read_group;
  if (c->filter_tag) F.has_tag = true, F.tag = c->filter_tag;
  if (c->filter_tag_value) F.has_tag_value = true, F.tag_value = c->filter_tag_value;
  if (c->blacklist_bed) load_blacklist(F, c->blacklist_bed, h, c->blacklist_merge_overlaps != 0);
  std::unordered_set<std::string> set;
  if (c->barcode_csv) {
    load_barcode_csv(set, c->barcode_csv, c->barcode_csv_column);
    F.has_barcodes = true;
  }
  if (c->has_barcode_list) {
    for (uint64_t i = 0; i < c->n_barcodes; i++) set.insert(c->barcodes[i]);
    F.has_barcodes = true;
  }
  if (F.has_barcodes) {
    std::vector<std::string> v(set.begin(), set.end());
    std::sort(v.begin(), v.end());
    F.bc_off.push_back(0);
    for (auto &s : v) {
      F.bc_pool += s;
      F.bc_off.push_back((uint32_t)F.bc_pool.size());
    }
    size_t cap = 16;
    while (cap < v.size() * 2 + 1) cap <<= 1;
    F.bc_table.assign(cap, 0);
    for (size_t i = 0; i < v.size(); i++) {
      uint64_t hsh = 0xcbf29ce484222325ull;  // FNV-1a 64, same as the device probe (pileup.cuh)
      for (unsigned char ch : v[i]) hsh = (hsh ^ ch) * 0x100000001b3ull;
      size_t slot = (size_t)((uint32_t)(hsh >> 17) & (uint32_t)(cap - 1));
      while (F.bc_table[slot]) slot = (slot + 1) & (cap - 1);
      F.bc_table[slot] = (uint32_t)i + 1;
    }
  }
  F.cfg.read_group = F.cfg.filter_tag = F.cfg.filter_tag_value = F.cfg.blacklist_bed = F.cfg.barcode_csv = nullptr;
  F.cfg.barcodes = nullptr;
  return F;
}

}  // namespace bnd
