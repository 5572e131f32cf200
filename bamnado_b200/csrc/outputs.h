// outputs.h — the output-side callers of the pileup: BigWig, multi-BAM TSV, collapse-bedgraph and the CLI mirror.
#pragma once
#include <string>

#include "../../include/bamnado_b200.h"
#include "engine.h"

namespace bnd {
void pileup_to_bigwig(Pileup &p, const std::string &path);                                           // coverage_analysis.rs:642-759
void multi_to_tsv(const bnd_pileup_cfg *cfgs, const bnd_filter_cfg *filters, int n, const char *out);  // :766-865
void collapse_bedgraph_file(const char *in_path, const char *out_path);                                // cli.rs:1149-1205
int cli_main(int argc, const char *const *argv);                                                       // cli.rs:737-757
}  // namespace bnd
