"""Seeded random sweeps over the option space of the coverage path: every drawn (pile-up, filter) configuration must give
the oracle's rows, counters and factor bit for bit.  The fixed tests in test_parity.py pin the BASELINE configs; this one
looks for interactions between options (fragment mode with shift, strand with proper pairs, odd bin sizes at chunk ends,
filters that reject everything ...) that nobody thought of writing down.  Reference: `BamPileup::new` and
`BamReadFilter::new` argument lists, bamnado/src/coverage_analysis.rs:298-324 and bamnado/src/read_filter.rs:438-492.
"""
import numpy as np
import pytest

from test_parity import assert_same


def draw(rng, bam):
    pile = dict(bam_path=bam, bin_size=int(rng.choice([1, 7, 10, 50, 100, 333, 1000, 4999, 100000])) if rng.random() < 0.9 else int(rng.integers(2, 3000)),
                norm=str(rng.choice(["raw", "cpm", "rpkm"])), scale_factor=float(rng.choice([1.0, 0.5, 2.25])),
                use_fragment=bool(rng.random() < 0.4), collapse=bool(rng.random() < 0.7), ignore_scaffold=bool(rng.random() < 0.3))
    if pile["bin_size"] == 1:
        pile["collapse"] = True  # 4 M rows otherwise; the uncollapsed path is drawn at the other bin sizes
    if rng.random() < 0.3:
        pile["shift"] = [int(x) for x in rng.integers(-40, 41, 4)]
    if rng.random() < 0.3:
        pile["truncate"] = (int(rng.integers(0, 60)) if rng.random() < 0.7 else None, int(rng.integers(0, 60)) if rng.random() < 0.7 else None)
    filt = dict(strand=str(rng.choice(["both", "forward", "reverse"], p=[0.6, 0.2, 0.2])), proper_pair=bool(rng.random() < 0.5),
                min_mapq=int(rng.choice([0, 1, 20, 30, 60, 255])), min_length=int(rng.choice([0, 20, 99, 101])),
                max_length=int(rng.choice([100, 150, 1000, 0xFFFFFFFF])), exclude_duplicates=bool(rng.random() < 0.4),
                exclude_secondary=bool(rng.random() < 0.4), exclude_supplementary=bool(rng.random() < 0.4))
    if rng.random() < 0.3:
        filt["min_fragment_length"] = int(rng.integers(0, 300))
    if rng.random() < 0.3:
        filt["max_fragment_length"] = int(rng.integers(100, 700))
    return pile, filt


def test_random_option_combinations_match_the_oracle(backend, oracle, tmp_path, make_bam):
    bams = [make_bam("rnd_pe", "--pairs", 6000, "--seed", 77, "--odd", "--contigs", "chr1:1300000,chr2:400001,chrUn_a:60000,chrM:16569"),
            make_bam("rnd_se", "--pairs", 5000, "--seed", 78, "--mode", "se", "--odd", "--contigs", "chr1:900000,chr7_random:50000,chrM:16569")]
    rng = np.random.default_rng(424242)
    for i in range(60):
        bam = bams[int(rng.integers(0, 2))]
        pile, filt = draw(rng, bam)
        try:
            assert_same(backend, oracle, pile, filt, tmp_path, text=pile["collapse"] and pile["bin_size"] > 1)
        except AssertionError as e:
            raise AssertionError(f"configuration {i}: {pile} {filt}") from e
