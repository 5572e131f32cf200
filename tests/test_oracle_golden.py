"""Pins the CPU oracle against the reference's own known-answer tests and fixtures.

Every expected value below is taken from a `#[test]` of the reference (file:line cited) or from the fixture
test/data/test_deeptools_raw.bw, copied to tests/golden/ref_data/.
"""
import numpy as np
import pytest

from bigwig_reader import read_bigwig
from conftest import GOLDEN, TEST_BAM

DEFAULT_FILTER = dict(proper_pair=True, min_mapq=0, min_length=0, max_length=1000)  # BamReadFilter::default(), read_filter.rs:403-420


def test_bam_stats_of_fixture(oracle):
    st = oracle.bam_stats(TEST_BAM, 1000)
    assert st == dict(n_mapped=11545, n_unmapped=0, genome_len=3213927757, chunk_len=5000000, n_chunks=1064, n_refs=456)


def test_known_answer_raw_max_117(oracle):
    # coverage_analysis.rs:989-1022: test.bam, bin 1000, default filter, collapse, Raw -> BigWig min 0.0, max 117.0
    runs, stats, factor = oracle.pileup_runs(dict(bam_path=TEST_BAM, bin_size=1000, norm="raw"), DEFAULT_FILTER)
    assert factor == 1.0
    assert runs["val"].max() == 117 and runs["val"].min() == 0
    assert stats["n_total"] == 11545 and sum(v for k, v in stats.items() if k != "n_total") == 0
    assert len(runs) == 4728


def test_known_answer_cpm(oracle):
    # coverage_analysis.rs:1024-1058: CPM -> max 10134.2568359375 after the f32 cast of to_bigwig (:705-710)
    runs, _, factor = oracle.pileup_runs(dict(bam_path=TEST_BAM, bin_size=1000, norm="cpm"), DEFAULT_FILTER)
    assert factor == 1e6 / 11545
    assert float(np.float32(runs["val"].max() * factor)) == 10134.2568359375


def test_known_answer_rpkm_no_collapse(oracle):
    # coverage_analysis.rs:1060-1094: RPKM at bin 1000, collapse = false -> the same 10134.2568359375
    runs, _, factor = oracle.pileup_runs(dict(bam_path=TEST_BAM, bin_size=1000, norm="rpkm", collapse=False), DEFAULT_FILTER)
    assert float(np.float32(runs["val"].max() * factor)) == 10134.2568359375
    assert len(runs) == 3214155


def test_deeptools_fixture_matches_except_chunk_terminal_base(oracle):
    # test/data/test_deeptools_raw.bw: identical to the reference semantics except that the reference leaves the last
    # base of every 5 Mb chunk uncovered (bam_utils.rs:57-58): +1 on each chunk-terminal end restores it
    bw = read_bigwig(GOLDEN / "ref_data" / "test_deeptools_raw.bw")
    assert bw["summary"]["max_val"] == 117.0 and bw["summary"]["min_val"] == 0.0
    runs, _, _ = oracle.pileup_runs(dict(bam_path=TEST_BAM, bin_size=1000, norm="raw"), DEFAULT_FILTER)
    mine = [(int(r["ref_id"]), int(r["start"]) - 1, int(r["stop"]) - 1, float(r["val"])) for r in runs]  # to_bigwig: -1 on both ends
    theirs = bw["intervals"]
    assert len(mine) == len(theirs) == 4728
    n_terminal = 0
    for a, b in zip(mine, theirs):
        assert a[0] == b[0] and a[1] == b[1] and a[3] == b[3]
        if a[2] != b[2]:
            assert a[2] + 1 == b[2]
            n_terminal += 1
    assert n_terminal == 1064  # one per chunk


def test_scale_factor_vectors(oracle):
    # signal_normalization.rs:81-103
    assert 1000.0 * oracle.scale_factor("raw", 1.0, 1000, 10_000_000) == 1000.0
    assert 1000.0 * oracle.scale_factor("cpm", 1.0, 1000, 10_000_000) == 100.0
    assert 1000.0 * oracle.scale_factor("rpkm", 1.0, 1000, 10_000_000) == 100.0
    assert oracle.scale_factor("cpm", 1.0, 50, 0) == 0.0


def _collapse(oracle, tmp_path, rows):
    src, dst = tmp_path / "in.bdg", tmp_path / "out.bdg"
    src.write_text("".join(f"{c}\t{s}\t{e}\t{v!r}\n" for c, s, e, v in rows))
    oracle.collapse_bedgraph(src, dst)
    out = []
    for l in dst.read_text().splitlines():
        c, s, e, v = l.split("\t")
        out.append((c, int(s), int(e), float(v)))
    return out


def test_collapse_adjacent_bins_contiguity_and_chrom_scoping(oracle, tmp_path):
    # bedgraph_utils.rs:82-108
    rows = [("chr1", 0, 10, 1.0), ("chr1", 10, 20, 1.0), ("chr1", 21, 30, 1.0), ("chr1", 30, 40, 2.0), ("chr2", 0, 5, 1.0), ("chr2", 5, 10, 1.0)]
    assert _collapse(oracle, tmp_path, rows) == [("chr1", 0, 20, 1.0), ("chr1", 21, 30, 1.0), ("chr1", 30, 40, 2.0), ("chr2", 0, 10, 1.0)]


def test_collapse_adjacent_bins_within_epsilon(oracle, tmp_path):
    # bedgraph_utils.rs:110-128: epsilon 1e-6 against the previous row; the group keeps its first score
    rows = [("chr1", 0, 10, 1.0), ("chr1", 10, 20, 1.0000004), ("chr1", 20, 30, 1.000002)]
    assert _collapse(oracle, tmp_path, rows) == [("chr1", 0, 20, 1.0), ("chr1", 20, 30, 1.000002)]


def test_collapse_adjacent_bins_empty(oracle, tmp_path):
    assert _collapse(oracle, tmp_path, []) == []  # bedgraph_utils.rs:66-79


def test_collapse_equal_bins_keeps_first_score(oracle):
    # coverage_analysis.rs:947-968: sum of the collapsed score column == 30
    c, s, e, sc = oracle.collapse_equal_bins([0, 0, 0, 0], [0, 1000, 2000, 3000], [1000, 2000, 3000, 4000], [10, 10, 20, 20])
    assert len(c) < 4 and int(sc.sum()) == 30
    assert list(s) == [0, 2000] and list(e) == [2000, 4000]


@pytest.mark.parametrize("value,text", [(117.0, "117.0"), (0.0, "0.0"), (86.617583369424, "86.617583369424"), (1e16, "1e16"),
                                        (1.5e-7, "1.5e-7"), (0.00001, "0.00001"), (0.000001, "1e-6"), (123456789012345680.0, "1.2345678901234568e17"),
                                        (2.5, "2.5"), (1e15, "1000000000000000.0")])
def test_float_text_follows_ryu(oracle, value, text):
    assert oracle.format_f64(value) == text


def test_filter_counters_and_order(oracle, make_bam):
    # read_filter.rs:838-899 pins: a duplicate / secondary / supplementary read bumps exactly its own counter when the
    # exclusion is on, the duplicate test comes first, and the three counters stay 0 when the exclusions are off
    bam = make_bam("flags", "--pairs", 4000, "--odd")
    off = oracle.pileup_runs(dict(bam_path=bam, bin_size=100), dict())[1]
    assert off["n_failed_duplicate"] == off["n_failed_secondary"] == off["n_failed_supplementary"] == 0
    on = oracle.pileup_runs(dict(bam_path=bam, bin_size=100), dict(exclude_duplicates=True, exclude_secondary=True, exclude_supplementary=True,
                                                                  min_mapq=30, min_length=50, max_length=500, strand="forward"))[1]
    assert on["n_failed_duplicate"] > 0 and on["n_failed_secondary"] > 0 and on["n_failed_supplementary"] > 0
    only_dup = oracle.pileup_runs(dict(bam_path=bam, bin_size=100), dict(exclude_duplicates=True))[1]
    assert only_dup["n_failed_duplicate"] == on["n_failed_duplicate"]  # checked before every other criterion
    assert on["n_total"] == off["n_total"]
