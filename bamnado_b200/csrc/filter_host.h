// filter_host.h — host-side construction of the read filter's lookup tables (what the reference's CLI / Python
// layers load before BamReadFilter::new: src/cli.rs:648-720, bam_utils.rs:155-270, 650-703).
#pragma once
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/bamnado_b200.h"
#include "bamio.h"

namespace bnd {

struct FilterHost {
  bnd_filter_cfg cfg;  // scalar fields only; strings are copied below
  bool has_rg = false, has_tag = false, has_tag_value = false;
  std::string rg, tag, tag_value;
  // blacklist, per reference slices of sorted starts / independently sorted stops (Lapper semantics)
  bool has_blacklist = false;
  std::vector<uint64_t> bl_off, bl_starts, bl_stops;
  // barcode allowlist
  bool has_barcodes = false;
  std::vector<uint32_t> bc_table, bc_off;
  std::string bc_pool;
};

FilterHost make_filter_host(const bnd_filter_cfg *cfg, const BamHeader &h);

}  // namespace bnd
