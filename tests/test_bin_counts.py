"""bin_counts::count_bins (bamnado/src/bin_counts.rs:100-289) — SURVEY §8(f) rank 1, the first widening of the hot path.

The CUDA path (pile-up without collapsing + one scatter kernel per sample) is compared with the oracle's restatement
(`oracle_py.count_bins`: the reference's grid and row arithmetic on top of the oracle's per-chunk bins), and the
reference's own tests (bin_counts.rs:303-338) are restated: the same BAM twice gives identical columns and equal,
non-zero library sizes; mismatched argument lists are an error.
"""
import numpy as np
import pytest

import bamnado_b200
from conftest import TEST_BAM

DEFAULT = dict(strand="both", proper_pair=True, min_mapq=0, min_length=0, max_length=1000)  # BamReadFilter::default()


def _piles(bams, bin_size, **kw):
    return [dict(bam_path=str(b), bin_size=bin_size, collapse=False, **kw) for b in bams]


def test_reference_tests_same_bam_twice(backend):
    # test_count_bins_same_bam_twice_gives_identical_columns (bin_counts.rs:303-325): bin 10 000, default filters,
    # reads, scaffolds ignored
    bc = backend.bin_counts(_piles([TEST_BAM, TEST_BAM], 10_000, ignore_scaffold=True), [DEFAULT, DEFAULT])
    assert bc["counts"].shape[0] > 0 and bc["counts"].shape[1] == 2
    assert bc["library_sizes"][0] == bc["library_sizes"][1] > 0
    assert int(bc["counts"].sum()) > 0
    assert (bc["counts"][:, 0] == bc["counts"][:, 1]).all()
    assert bc["chroms"] == sorted(bc["chroms"]) and not any(c.startswith("chrUn") or c.endswith(("_alt", "_random")) for c in bc["chroms"])


@pytest.mark.parametrize("bin_size,ignore_scaffold,use_fragment", [(10_000, True, False), (1000, False, False), (777, True, True)])
def test_matches_oracle_on_test_bam(backend, oracle, bin_size, ignore_scaffold, use_fragment):
    filt = dict(DEFAULT, min_mapq=10)
    bc = backend.bin_counts(_piles([TEST_BAM], bin_size, ignore_scaffold=ignore_scaffold, use_fragment=use_fragment), [filt])
    exp_counts, grid, libs = oracle.count_bins([str(TEST_BAM)], bin_size, [filt], use_fragment, ignore_scaffold)
    assert bc["chroms"] == [g[0] for g in grid]
    assert list(bc["chrom_len"]) == [g[1] for g in grid]
    assert list(bc["chrom_first_row"][:-1]) == [g[2] for g in grid] and int(bc["chrom_first_row"][-1]) == exp_counts.shape[0]
    assert bc["library_sizes"] == libs
    assert bc["counts"].dtype == np.uint64 and (bc["counts"] == exp_counts).all()


def test_two_samples_different_filters_and_scaffolds(backend, oracle, make_bam):
    # two synthetic BAMs over the same contigs (scaffold contigs included), different seeds and filters per sample
    contigs = "chr1:3000000,chr2:1200001,chrUn_x:90000,chr3_random:50000,chrM:16569"
    a = make_bam("bc_a", "--pairs", 30000, "--seed", 5, "--contigs", contigs, "--odd")
    b = make_bam("bc_b", "--pairs", 20000, "--seed", 6, "--contigs", contigs, "--odd")
    filts = [dict(DEFAULT, proper_pair=False, min_mapq=20), dict(DEFAULT, min_mapq=30, exclude_duplicates=True)]
    for ignore_scaffold in (False, True):
        bc = backend.bin_counts(_piles([a, b], 500, ignore_scaffold=ignore_scaffold), filts)
        exp_counts, grid, libs = oracle.count_bins([a, b], 500, filts, False, ignore_scaffold)
        assert bc["chroms"] == [g[0] for g in grid]
        assert bc["library_sizes"] == libs
        assert (bc["counts"] == exp_counts).all()
        # chr2 is 1 200 001 long: its last grid bin is the single base after the last chunk's bins and stays zero
        r = bc["chroms"].index("chr2")
        assert int(bc["chrom_first_row"][r + 1] - bc["chrom_first_row"][r]) == 2401 and int(bc["counts"][bc["chrom_first_row"][r + 1] - 1].sum()) == 0


def test_argument_and_grid_errors(backend, make_bam):
    # test_count_bins_mismatched_bam_paths_and_sample_names_errors (bin_counts.rs:327-338) + the shared-grid check (:134-146)
    with pytest.raises(RuntimeError, match="does not match number of sample names"):
        bamnado_b200.count_bins([str(TEST_BAM)], ["a", "b"], 10_000)
    with pytest.raises(RuntimeError, match="At least one BAM"):
        bamnado_b200.count_bins([], [], 10_000)
    other = make_bam("bc_other", "--pairs", 2000, "--contigs", "chr1:100000,chr2:50000")
    with pytest.raises(RuntimeError, match="Chromosome name/length mismatch"):
        backend.bin_counts(_piles([TEST_BAM, other], 10_000), [DEFAULT, DEFAULT])


@pytest.mark.gpu
def test_python_mirror_on_gpu(native_gpu):
    bc = bamnado_b200.count_bins([str(TEST_BAM), str(TEST_BAM)], ["a", "b"], 10_000, ignore_scaffold_chromosomes=True)
    assert bc.sample_names == ["a", "b"] and bc.bin_size == 10_000
    assert bc.counts.shape == (bc.n_bins, 2) and (bc.counts[:, 0] == bc.counts[:, 1]).all()
    chrom_idx, start, end = bc.bin_arrays()
    assert len(start) == bc.n_bins and int(start[0]) == 1 and int(end[0]) == 10_000
    first = next(iter(bc.bins))
    assert first == (bc.chroms[0], 1, 10_000)
