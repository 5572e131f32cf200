// bigwig_layout.h — host side of the BigWig writer: little-endian buffer, chromosome B+ tree and cirTree (R-tree) index.
// The sections themselves (items, zlib streams, zoom records) are produced on the device (bigwig.cuh); this is the file
// layout around them (what bigtools' BigWigWrite does after its workers return, bamnado/src/coverage_analysis.rs:748-755).
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <tuple>
#include <vector>

namespace bnd {
namespace bwl {

// ---- little-endian writer ---------------------------------------------------------------------------------------------
struct Buf {
  std::vector<uint8_t> d;
  void u8(uint8_t v) { d.push_back(v); }
  void u16(uint16_t v) {
    for (int i = 0; i < 2; i++) d.push_back((uint8_t)(v >> (8 * i)));
  }
  void u32(uint32_t v) {
    for (int i = 0; i < 4; i++) d.push_back((uint8_t)(v >> (8 * i)));
  }
  void u64(uint64_t v) {
    for (int i = 0; i < 8; i++) d.push_back((uint8_t)(v >> (8 * i)));
  }
  void f32(float v) {
    uint32_t b;
    memcpy(&b, &v, 4);
    u32(b);
  }
  void f64(double v) {
    uint64_t b;
    memcpy(&b, &v, 8);
    u64(b);
  }
  void bytes(const void *p, size_t n) { d.insert(d.end(), (const uint8_t *)p, (const uint8_t *)p + n); }
  void pad_to(size_t n) { d.resize(std::max(d.size(), n), 0); }
  void put_u64_at(size_t off, uint64_t v) {
    for (int i = 0; i < 8; i++) d[off + i] = (uint8_t)(v >> (8 * i));
  }
};

struct Section {  // one BigWig data section (<= ITEMS_PER_SLOT bedGraph items of one chromosome)
  uint32_t chrom, start, end;
  uint64_t offset, size;
};
constexpr uint32_t ITEMS_PER_SLOT = 1024, BLOCK = 256, MAX_ZOOM = 10;

// B+ tree over (name -> id, size), keys sorted bytewise
inline void write_chrom_tree(Buf &b, std::vector<std::tuple<std::string, uint32_t, uint32_t>> items) {
  std::sort(items.begin(), items.end());
  uint32_t key_size = 1;
  for (auto &it : items) key_size = std::max<uint32_t>(key_size, (uint32_t)std::get<0>(it).size());
  const uint64_t n = items.size();
  b.u32(0x78CA8C91);
  b.u32(BLOCK);
  b.u32(key_size);
  b.u32(8);
  b.u64(n);
  b.u64(0);
  // levels: leaves hold up to BLOCK items; each upper node up to BLOCK children
  std::vector<uint64_t> level_nodes;  // nodes per level, leaf level first
  uint64_t cnt = std::max<uint64_t>(n, 1);
  level_nodes.push_back((cnt + BLOCK - 1) / BLOCK);
  while (level_nodes.back() > 1) level_nodes.push_back((level_nodes.back() + BLOCK - 1) / BLOCK);
  const int levels = (int)level_nodes.size();
  const uint64_t leaf_item = key_size + 8, inner_item = key_size + 8;
  // items covered by one node at level l (0 = leaf)
  auto span = [&](int l) {
    uint64_t s = BLOCK;
    for (int i = 0; i < l; i++) s *= BLOCK;
    return s;
  };
  // offsets of each level (root first in the file)
  std::vector<uint64_t> level_off(levels);
  uint64_t off = b.d.size();
  for (int l = levels - 1; l >= 0; l--) {
    level_off[l] = off;
    uint64_t nodes = level_nodes[l];
    // every node is written at full size so child offsets are arithmetic
    off += nodes * (4 + (uint64_t)BLOCK * (l == 0 ? leaf_item : inner_item));
  }
  auto key = [&](const std::string &s) {
    std::string k = s;
    k.resize(key_size, '\0');
    return k;
  };
  for (int l = levels - 1; l >= 0; l--) {
    const uint64_t nodes = level_nodes[l], per = span(l), item_sz = l == 0 ? leaf_item : inner_item;
    for (uint64_t nd = 0; nd < nodes; nd++) {
      const uint64_t first = nd * per;
      uint64_t children;
      if (l == 0) children = n > first ? std::min<uint64_t>(BLOCK, n - first) : 0;
      else {
        const uint64_t child_span = span(l - 1);
        children = n > first ? std::min<uint64_t>(BLOCK, (n - first + child_span - 1) / child_span) : 0;
      }
      b.u8(l == 0 ? 1 : 0);
      b.u8(0);
      b.u16((uint16_t)children);
      for (uint64_t c = 0; c < BLOCK; c++) {
        if (c < children) {
          if (l == 0) {
            auto &it = items[first + c];
            b.bytes(key(std::get<0>(it)).data(), key_size);
            b.u32(std::get<1>(it));
            b.u32(std::get<2>(it));
          } else {
            const uint64_t child_span = span(l - 1), child = nd * BLOCK + c;
            b.bytes(key(std::get<0>(items[first + c * child_span])).data(), key_size);
            b.u64(level_off[l - 1] + child * (4 + (uint64_t)BLOCK * (l - 1 == 0 ? leaf_item : inner_item)));
          }
        } else {
          for (uint64_t z = 0; z < item_sz; z++) b.u8(0);
        }
      }
    }
  }
}

// cirTree (R-tree) over the sections, built bottom-up with full-size nodes.  Child offsets inside the tree are absolute file
// offsets: `base` is where b's first byte will sit in the file.
inline void write_rtree(Buf &b, const std::vector<Section> &secs, uint64_t end_file_offset, uint64_t base) {
  const uint64_t n = secs.size();
  b.u32(0x2468ACE0);
  b.u32(BLOCK);
  b.u64(n);
  b.u32(n ? secs.front().chrom : 0);
  b.u32(n ? secs.front().start : 0);
  b.u32(n ? secs.back().chrom : 0);
  b.u32(n ? secs.back().end : 0);
  b.u64(end_file_offset);
  b.u32(ITEMS_PER_SLOT);
  b.u32(0);
  struct Node {
    uint32_t sc, sb, ec, eb;
  };
  // level 0 = leaf nodes over sections
  std::vector<std::vector<Node>> bounds;  // per level: bounding box of each node
  std::vector<uint64_t> level_nodes;
  uint64_t cnt = std::max<uint64_t>(n, 1);
  level_nodes.push_back((cnt + BLOCK - 1) / BLOCK);
  while (level_nodes.back() > 1) level_nodes.push_back((level_nodes.back() + BLOCK - 1) / BLOCK);
  const int levels = (int)level_nodes.size();
  bounds.resize(levels);
  auto merge = [](Node &a, const Node &x, bool first) {
    if (first) {
      a = x;
      return;
    }
    if (x.sc < a.sc || (x.sc == a.sc && x.sb < a.sb)) a.sc = x.sc, a.sb = x.sb;
    if (x.ec > a.ec || (x.ec == a.ec && x.eb > a.eb)) a.ec = x.ec, a.eb = x.eb;
  };
  for (int l = 0; l < levels; l++) {
    bounds[l].resize(level_nodes[l]);
    for (uint64_t nd = 0; nd < level_nodes[l]; nd++) {
      Node box{0, 0, 0, 0};
      bool first = true;
      const uint64_t lo = nd * BLOCK, hi = l == 0 ? std::min<uint64_t>(n, lo + BLOCK) : std::min<uint64_t>(level_nodes[l - 1], lo + BLOCK);
      for (uint64_t c = lo; c < hi; c++) {
        Node x = l == 0 ? Node{secs[c].chrom, secs[c].start, secs[c].chrom, secs[c].end} : bounds[l - 1][c];
        merge(box, x, first);
        first = false;
      }
      bounds[l][nd] = box;
    }
  }
  std::vector<uint64_t> level_off(levels);
  uint64_t off = base + b.d.size();
  for (int l = levels - 1; l >= 0; l--) {
    level_off[l] = off;
    off += level_nodes[l] * (4 + (uint64_t)BLOCK * (l == 0 ? 32 : 24));
  }
  for (int l = levels - 1; l >= 0; l--) {
    for (uint64_t nd = 0; nd < level_nodes[l]; nd++) {
      const uint64_t lo = nd * BLOCK, hi = l == 0 ? std::min<uint64_t>(n, lo + BLOCK) : std::min<uint64_t>(level_nodes[l - 1], lo + BLOCK);
      const uint64_t children = hi > lo ? hi - lo : 0;
      b.u8(l == 0 ? 1 : 0);
      b.u8(0);
      b.u16((uint16_t)children);
      for (uint64_t c = 0; c < BLOCK; c++) {
        if (c < children) {
          if (l == 0) {
            const Section &s = secs[lo + c];
            b.u32(s.chrom), b.u32(s.start), b.u32(s.chrom), b.u32(s.end);
            b.u64(s.offset), b.u64(s.size);
          } else {
            const Node &x = bounds[l - 1][lo + c];
            b.u32(x.sc), b.u32(x.sb), b.u32(x.ec), b.u32(x.eb);
            b.u64(level_off[l - 1] + (lo + c) * (4 + (uint64_t)BLOCK * (l - 1 == 0 ? 32 : 24)));
          }
        } else {
          for (int z = 0; z < (l == 0 ? 32 : 24); z++) b.u8(0);
        }
      }
    }
  }
}


}  // namespace bwl
}  // namespace bnd
