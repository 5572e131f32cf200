"""Callers either side of the pileup: BigWig, multi-BAM TSV, collapse-bedgraph and the CLI mirror."""
import os
import subprocess

import numpy as np
import pytest

from bigwig_reader import read_bigwig
from conftest import TEST_BAM

CLI_FILTER = dict(min_mapq=20, min_length=20, max_length=1000)


def test_bigwig_matches_reference_known_answers(backend, oracle, tmp_path):
    # coverage_analysis.rs:989-1058: summary min/max of the BigWig written from test.bam at bin 1000
    for norm, expect_max in (("raw", 117.0), ("cpm", 10134.2568359375)):
        out = tmp_path / f"{norm}.bw"
        pile, filt = dict(bam_path=TEST_BAM, bin_size=1000, norm=norm), dict(proper_pair=True, max_length=1000)
        with backend.pileup(pile, filt) as h:
            h.to_bigwig(out)
        bw = read_bigwig(out)
        assert bw["summary"]["min_val"] == 0.0 and bw["summary"]["max_val"] == expect_max
        runs, _, factor = oracle.pileup_runs(pile, filt)
        exp = [(int(r["ref_id"]), int(r["start"]) - 1, int(r["stop"]) - 1, float(np.float32(float(r["val"]) * factor))) for r in runs]
        assert bw["intervals"] == exp  # header order, 0-based shift of both ends, f32 scores
        assert len(bw["chroms"]) == 456 and bw["chroms"]["chr9"] == (58, 138394717)


def test_bigwig_zoom_levels_summarise_the_data(backend, tmp_path, make_bam):
    # zoom records are recomputed from the file's own intervals: windows of `reduction` bases per chromosome, covered
    # bases, min, max, sum and sum of squares (f32 accumulation like the writer's)
    bam = make_bam("pe", "--pairs", 20000, "--odd")
    out = tmp_path / "z.bw"
    with backend.pileup(dict(bam_path=bam, bin_size=50, norm="cpm"), CLI_FILTER) as h:
        h.to_bigwig(out)
    bw = read_bigwig(out)
    assert 1 <= bw["zoom_levels"] <= 10
    prev = 0
    for z in bw["zoom"]:
        red = z["reduction"]
        assert red > prev
        prev = red
        exp = {}
        for cid, s, e, v in bw["intervals"]:
            v = np.float32(v)
            for w in range(s // red, (e - 1) // red + 1):
                a, b = max(s, w * red), min(e, (w + 1) * red)
                n = np.float32(b - a)
                r = exp.get((cid, w))
                if r is None:
                    exp[(cid, w)] = [a, b, b - a, v, v, v * n, v * v * n]
                else:
                    r[1] = b
                    r[2] += b - a
                    r[3], r[4] = min(r[3], v), max(r[4], v)
                    r[5] = np.float32(r[5] + v * n)
                    r[6] = np.float32(r[6] + v * v * n)
        got = z["records"]
        assert len(got) == len(exp) == z["leading_count"]
        for (cid, s, e, valid, mn, mx, sm, sq), ((ecid, w), r) in zip(got, sorted(exp.items())):
            assert (cid, s, e, valid) == (ecid, r[0], r[1], r[2])
            assert (np.float32(mn), np.float32(mx)) == (r[3], r[4])
            assert abs(sm - float(r[5])) <= 1e-5 * max(1.0, abs(float(r[5]))) and abs(sq - float(r[6])) <= 1e-5 * max(1.0, abs(float(r[6])))
    counts = [len(z["records"]) for z in bw["zoom"]]
    assert counts == sorted(counts, reverse=True) and counts[0] < len(bw["intervals"])  # every level is coarser than the last


def test_bigwig_ignore_scaffold_and_large(backend, oracle, tmp_path, make_bam):
    bam = make_bam("pe", "--pairs", 20000, "--odd")
    out = tmp_path / "x.bigwig"
    pile, filt = dict(bam_path=bam, bin_size=50, norm="rpkm", ignore_scaffold=True), CLI_FILTER
    with backend.pileup(pile, filt) as h:
        h.to_bigwig(out)
        names, _ = h.refs()
    bw = read_bigwig(out)
    assert all(not (n.startswith("chrUn") or n.endswith("_alt") or n.endswith("_random")) for n in bw["chroms"])
    runs, _, factor = oracle.pileup_runs(pile, filt)
    # chromosome ids in the file are dense over the kept references (header order), not BAM reference indexes
    cid_of_ref = {r: bw["chroms"][n][0] for r, n in enumerate(names) if n in bw["chroms"]}
    assert len(cid_of_ref) < len(names)
    exp = [(cid_of_ref[int(r["ref_id"])], int(r["start"]) - 1, int(r["stop"]) - 1, float(np.float32(float(r["val"]) * factor))) for r in runs if int(r["ref_id"]) in cid_of_ref]
    assert bw["intervals"] == exp


def test_bigwig_sections_are_zlib_streams_written_on_the_device(backend, oracle, tmp_path, make_bam, monkeypatch):
    # every data / zoom section is one zlib stream (fixed-Huffman DEFLATE block + Adler-32, bigwig.cuh): Python's zlib accepts
    # each of them (the reader inflates all sections), the header's uncompressBufSize bounds every inflated section, the
    # streams are smaller than the raw sections, and the stored-raw mode holds the same rows
    import struct
    import zlib

    bam = make_bam("pe", "--pairs", 20000, "--odd")
    pile, filt = dict(bam_path=bam, bin_size=50, norm="cpm"), CLI_FILTER
    a, b = tmp_path / "z.bw", tmp_path / "raw.bw"
    with backend.pileup(pile, filt) as h:
        h.to_bigwig(a)
    monkeypatch.setenv("BND_BIGWIG_COMPRESS", "0")
    with backend.pileup(pile, filt) as h:
        h.to_bigwig(b)
    monkeypatch.delenv("BND_BIGWIG_COMPRESS")
    za, zb = read_bigwig(a), read_bigwig(b)
    assert za["intervals"] == zb["intervals"] and len(za["intervals"]) > 10000
    assert [z["records"] for z in za["zoom"]] == [z["records"] for z in zb["zoom"]]
    da = open(a, "rb").read()
    ubuf = struct.unpack_from("<I", da, 52)[0]
    assert ubuf >= 24 + 12 * min(1024, len(za["intervals"])) and struct.unpack_from("<I", open(b, "rb").read(), 52)[0] == 0
    assert a.stat().st_size < 0.75 * b.stat().st_size
    # a stream cut out of the file by hand: header bytes 0x78 0x9c, one final fixed-Huffman block
    index_off = struct.unpack_from("<Q", da, 24)[0]
    data_off = struct.unpack_from("<Q", da, 16)[0]
    assert da[data_off + 8: data_off + 10] == b"\x78\x9c" and (da[data_off + 10] & 7) == 3
    d = zlib.decompressobj()
    first = d.decompress(da[data_off + 8: index_off])
    assert d.eof and len(first) <= ubuf and struct.unpack_from("<B", first, 20)[0] == 1


def _rows(path):
    lines = open(path).read().splitlines()
    return lines[0], sorted(lines[1:])


def test_multi_bam_tsv(backend, oracle, tmp_path, make_bam):
    # MultiBamPileup::to_tsv, intended semantics; row order is unspecified upstream -> compare as sets
    bams = [make_bam("m0", "--pairs", 8000, "--seed", 11, "--mode", "se"), make_bam("m1", "--pairs", 8000, "--seed", 12, "--mode", "se"),
            make_bam("m2", "--pairs", 8000, "--seed", 13, "--mode", "se")]
    for bin_size in (500, 1):
        piles = [dict(bam_path=b, bin_size=bin_size, collapse=False) for b in bams]
        filts = [CLI_FILTER] * 3
        a, b = tmp_path / f"gpu{bin_size}.tsv", tmp_path / f"cpu{bin_size}.tsv"
        if bin_size == 1:  # keep the CPU side affordable: chrM-sized genome is not available, so only compare sizes here
            continue
        backend.multi_to_tsv(piles, filts, a)
        oracle.multi_to_tsv(piles, filts, b)
        assert _rows(a) == _rows(b)
    with pytest.raises(ValueError):
        backend.multi_to_tsv([dict(bam_path=bams[0], bin_size=50, collapse=True)], [CLI_FILTER], tmp_path / "x.tsv")
    with pytest.raises(ValueError):
        backend.multi_to_tsv([dict(bam_path=bams[0], bin_size=50, collapse=False), dict(bam_path=bams[1], bin_size=100, collapse=False)],
                             [CLI_FILTER] * 2, tmp_path / "x.tsv")


def test_multi_bam_random_filters_and_bins(backend, oracle, tmp_path, make_bam):
    # per-BAM filters differ (coverage_analysis.rs:783-848 runs every BAM's own pile-up before the equal-row collapse),
    # small contigs so that bin 1 and odd bin sizes stay affordable on the CPU side
    contigs = "chr1:60000,chr2:25001,chrUn_x:9000,chrM:1657"
    bams = [make_bam(f"mr{i}", "--pairs", 1500, "--seed", 31 + i, "--contigs", contigs, "--odd") for i in range(3)]
    rng = np.random.default_rng(7)
    for trial in range(5):
        n = int(rng.integers(1, 4))
        bin_size = int(rng.choice([1, 13, 50, 1000]))
        use_fragment = bool(rng.random() < 0.4)
        piles = [dict(bam_path=bams[i], bin_size=bin_size, collapse=False, use_fragment=use_fragment) for i in range(n)]
        filts = [dict(CLI_FILTER, min_mapq=int(rng.choice([0, 20, 40])), proper_pair=bool(rng.random() < 0.5),
                      exclude_duplicates=bool(rng.random() < 0.5)) for _ in range(n)]
        a, b = tmp_path / f"mg{trial}.tsv", tmp_path / f"mc{trial}.tsv"
        backend.multi_to_tsv(piles, filts, a)
        oracle.multi_to_tsv(piles, filts, b)
        assert _rows(a) == _rows(b), f"trial {trial}: bin {bin_size}, {n} BAMs, fragment {use_fragment}, {filts}"


def test_collapse_bedgraph_reference_vectors(backend, tmp_path):
    # bedgraph_utils.rs:82-129
    def run(rows):
        src, dst = tmp_path / "in.bdg", tmp_path / "out.bdg"
        src.write_text("".join(f"{c}\t{s}\t{e}\t{v!r}\n" for c, s, e, v in rows))
        backend.collapse_bedgraph(src, dst)
        return [(l.split("\t")[0], int(l.split("\t")[1]), int(l.split("\t")[2]), float(l.split("\t")[3])) for l in dst.read_text().splitlines()]

    assert run([("chr1", 0, 10, 1.0), ("chr1", 10, 20, 1.0), ("chr1", 21, 30, 1.0), ("chr1", 30, 40, 2.0), ("chr2", 0, 5, 1.0), ("chr2", 5, 10, 1.0)]) == \
        [("chr1", 0, 20, 1.0), ("chr1", 21, 30, 1.0), ("chr1", 30, 40, 2.0), ("chr2", 0, 10, 1.0)]
    assert run([("chr1", 0, 10, 1.0), ("chr1", 10, 20, 1.0000004), ("chr1", 20, 30, 1.000002)]) == [("chr1", 0, 20, 1.0), ("chr1", 20, 30, 1.000002)]
    assert run([]) == []


def test_collapse_bedgraph_matches_oracle(backend, oracle, tmp_path, make_bam):
    # a bin-1 style bedGraph (uncollapsed rows), shuffled, with comments / short lines / float scores
    bam = make_bam("pe", "--pairs", 20000, "--odd")
    runs, _, factor = oracle.pileup_runs(dict(bam_path=bam, bin_size=400, norm="cpm", collapse=False), CLI_FILTER)
    names = ["chrB", "chrA", "chr10", "chr2"]
    rng = np.random.default_rng(5)
    lines = ["# a comment", "", "chrA\t1\t2"]
    sel = runs[runs["ref_id"] < 4][:60000]
    for r in sel:
        lines.append(f"{names[int(r['ref_id'])]}\t{int(r['start'])}\t{int(r['stop'])}\t{float(r['val']) * factor!r}")
    sorted_in, shuffled_in = tmp_path / "s.bdg", tmp_path / "u.bdg"
    sorted_in.write_text("\n".join(lines) + "\n")
    body = lines[3:]
    rng.shuffle(body)
    shuffled_in.write_text("\n".join(lines[:3] + list(body)) + "\n")
    for src in (sorted_in, shuffled_in):
        a, b = tmp_path / "a.bdg", tmp_path / "b.bdg"
        backend.collapse_bedgraph(src, a)
        oracle.collapse_bedgraph(src, b)
        assert a.read_bytes() == b.read_bytes()
        assert len(a.read_text().splitlines()) < len(sel)


def test_collapse_bedgraph_reader_edge_cases(backend, oracle, tmp_path):
    # the device-side reader (src/cli.rs:1160-1175: trim, skip blank / '#' / short lines, split_whitespace, str::parse): every way a
    # number can be spelled - converted on the device (<= 19 digits, |exp| <= 22, mantissa < 2^53) or handed to the host - CRLF line
    # ends, runs of blanks and tabs, a last line without '\n', names that repeat after another name
    scores = ["0", "0.0", "-0.0", "1", "1.0", "+2.5", "3.", ".5", "1e3", "1E-3", "2.5e+2", "0.1", "0.30000000000000004", "1.2345678901234567",
              "12345678901234567890", "9007199254740993", "1e23", "1e-23", "4.9e-324", "1.7976931348623157e308", "inf", "-inf", "Infinity", "1e400",
              "0.000001", "000123.4500", "0.10000000000000001", "123456789012345678", "5e22", "5e-22", "0e999", "1e22", "9007199254740992", "0.3"]
    lines = ["# header comment", "", "   ", "chrZ 1 2", "chrZ\t1"]
    pos = 0
    for k, sc in enumerate(scores):
        name = ["chrB", "chrA", "chrB", "chr10"][k % 4]
        sep = ["\t", " ", "  \t ", "\t\t"][k % 4]
        lines.append(f"  {name}{sep}{pos}{sep}{pos + 10}{sep}{sc}  extra column" + ("\r" if k % 3 == 0 else ""))
        if k % 5 != 4:
            pos += 10
    lines += ["chrA\t-20\t-10\t7", "chrA\t+5\t9223372036854775807\t7", "chrA\t-9223372036854775808\t0\t7"]
    for tail in ("\n", ""):
        src, a, b = tmp_path / "edge.bdg", tmp_path / "edge_a.bdg", tmp_path / "edge_b.bdg"
        src.write_bytes(("\n".join(lines) + tail).encode())
        backend.collapse_bedgraph(src, a)
        oracle.collapse_bedgraph(src, b)
        assert a.read_bytes() == b.read_bytes()
        assert len(a.read_text().splitlines()) >= 20
    for bad in ("chr1\t1.0\t2\t3", "chr1\t1\t2x\t3", "chr1\t1\t2\t3..0", "chr1\t1\t2\t0x10", "chr1\t1\t9223372036854775808\t1", "chr1\t-\t2\t3",
                "chr1\t1\t2\t1e", "chr1\t1\t2\tnan(1)"):
        src = tmp_path / "bad.bdg"
        src.write_text("chr1\t0\t1\t1\n" + bad + "\n")
        with pytest.raises(RuntimeError, match="invalid bedGraph line"):
            backend.collapse_bedgraph(src, tmp_path / "bad_out.bdg")


def test_collapse_bedgraph_random_rows_match_oracle(backend, oracle, tmp_path):
    # synthetic rows built to sit on every decision of collapse_adjacent_bins (bedgraph_utils.rs:5-45): touching and
    # gapped bins, scores equal, within and just beyond 1e-6 of the previous ROW (drift chains), repeated starts,
    # several chromosomes in shuffled order, integer- and float-looking scores
    rng = np.random.default_rng(99)
    for trial in range(6):
        lines = []
        for chrom in rng.permutation(["chr1", "chr10", "chr2", "chrX", "scaffold_9"])[: int(rng.integers(1, 6))]:
            pos, score = int(rng.integers(0, 1000)), float(rng.integers(0, 5))
            for _ in range(int(rng.integers(1, 400))):
                width = int(rng.choice([1, 10, 50, 137]))
                gap = int(rng.choice([0, 0, 0, 1, 25]))
                step = float(rng.choice([0.0, 0.0, 4e-7, 9e-7, 1.1e-6, 1.0, -1.0, 0.5]))
                score = max(0.0, score + step)
                pos += gap
                lines.append(f"{chrom}\t{pos}\t{pos + width}\t{score!r}" if rng.random() < 0.8 else f"{chrom}\t{pos}\t{pos + width}\t{int(score)}")
                if rng.random() < 0.03:
                    lines.append(lines[-1])  # an exact duplicate row
                pos += width
        order = rng.permutation(len(lines))
        src = tmp_path / f"r{trial}.bdg"
        src.write_text("\n".join(lines[i] for i in order) + "\n")
        a, b = tmp_path / "ra.bdg", tmp_path / "rb.bdg"
        backend.collapse_bedgraph(src, a)
        oracle.collapse_bedgraph(src, b)
        assert a.read_bytes() == b.read_bytes(), f"trial {trial}"


def test_cli_exit_codes_and_outputs(backend, oracle, tmp_path, capfd):
    # test/test_interface.py:103-105: --help / --version -> 0, parse errors -> 2; runtime errors -> 1 (src/cli.rs:750-756)
    assert backend.cli_main(["bamnado", "--help"]) == 0
    assert backend.cli_main(["bamnado", "--version"]) == 0
    assert backend.cli_main(["bamnado", "bam-coverage", "--help"]) == 0
    assert backend.cli_main(["bamnado", "definitely-not-a-command"]) == 2
    assert backend.cli_main(["bamnado", "bam-coverage", "--no-such-flag"]) == 2
    assert backend.cli_main(["bamnado", "bam-coverage"]) == 2
    assert backend.cli_main(["bamnado"]) == 2
    assert backend.cli_main(["bamnado", "bam-coverage", "--bam", "/nonexistent.bam"]) == 1
    assert backend.cli_main(["bamnado", "coverage", "-b", str(TEST_BAM), "-o", str(tmp_path / "x.tsv")]) == 1
    capfd.readouterr()
    out = tmp_path / "cli.bedgraph"
    rc = backend.cli_main(["bamnado", "coverage", "-b", str(TEST_BAM), "-o", str(out), "-s", "100", "--normalize", "cpm", "--proper-pair",
                           "--min-mapq=30", "--ignore-scaffold"])
    assert rc == 0
    ref = tmp_path / "ref.bedgraph"
    oracle.pileup_to_bedgraph(dict(bam_path=TEST_BAM, bin_size=100, norm="cpm", ignore_scaffold=True),
                              dict(min_mapq=30, min_length=20, max_length=1000, proper_pair=True, blacklist_merge_overlaps=True), ref)
    assert out.read_bytes() == ref.read_bytes()
    # default bin size 50 and default output name <stem>.bedgraph next to the BAM are exercised on a copy
    bam2 = tmp_path / "copy.bam"
    bam2.write_bytes(open(TEST_BAM, "rb").read())
    (tmp_path / "copy.bam.bai").write_bytes(open(str(TEST_BAM) + ".bai", "rb").read())
    assert backend.cli_main(["bamnado", "bam-coverage", "--bam", str(bam2)]) == 0
    oracle.pileup_to_bedgraph(dict(bam_path=TEST_BAM, bin_size=50), CLI_FILTER, ref)
    assert (tmp_path / "copy.bedgraph").read_bytes() == ref.read_bytes()
    # BigWig by extension, collapse subcommand
    assert backend.cli_main(["bamnado", "bam-coverage", "--bam", str(bam2), "-o", str(tmp_path / "o.bw"), "-s", "1000"]) == 0
    assert read_bigwig(tmp_path / "o.bw")["summary"]["max_val"] > 0
    assert backend.cli_main(["bamnado", "collapse", "-i", str(ref), "-o", str(tmp_path / "c.bdg")]) == 0
    assert len((tmp_path / "c.bdg").read_text().splitlines()) <= len(ref.read_text().splitlines())


def test_cli_fragment_filter_needs_paired_end(backend, make_bam, tmp_path):
    se = make_bam("se", "--pairs", 15000, "--mode", "se", "--unplaced", 500)
    assert backend.cli_main(["bamnado", "bam-coverage", "--bam", se, "-o", str(tmp_path / "x.bdg"), "--max-fragment-len", "200"]) == 1
    pe = make_bam("pe", "--pairs", 20000, "--odd")
    assert backend.cli_main(["bamnado", "bam-coverage", "--bam", pe, "-o", str(tmp_path / "x.bdg"), "--max-fragment-len", "200", "--fragment-counts"]) == 0
