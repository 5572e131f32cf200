"""bigwig-compare / bigwig-aggregate (bamnado/src/bigwig_compare.rs:616-790).

The oracle (oracle/bigwig_compare_py.py) is pinned by the reference's own test vectors first, then the device path is compared
with it on the same inputs: subtraction, ratio, sum, mean, max, min and median bit-exact (same f64 additions in the same order),
log-ratio within 1e-6 (device log vs libm)."""
from __future__ import annotations

import math
import random

import numpy as np
import pytest

from bigwig_reader import read_bigwig
from bigwig_writer import write_bigwig
from oracle import bigwig_compare_py as O

f32 = np.float32


# ---- the oracle against the reference's vectors (bigwig_compare.rs:823-1248) -------------------------------------------------
def test_oracle_merge_of_equal_bins_reference_vector():
    # test_write_bedgraph_bins_collapses_equal_signals :823-862
    assert O.merge_bins([f32(1), f32(1), f32(2), f32(2), f32(1)], 500, 100) == [(0, 200, f32(1)), (200, 400, f32(2)), (400, 500, f32(1))]


def test_oracle_compare_reference_vectors():
    a = [(0, 100, 10.0), (100, 200, 20.0)]
    b = [(0, 100, 5.0), (100, 200, 10.0)]
    sub = O.compare(a, b, 1000, 10, "subtraction")  # :864-923
    assert len(sub) == 100 and abs(sub[0] - 5.0) < 1e-5 and abs(sub[10] - 10.0) < 1e-5 and sub[20] == 0
    ratio = O.compare([(0, 100, 10.0)], [(0, 100, 5.0)], 1000, 10, "ratio")  # :925-971
    assert abs(ratio[0] - 2.0) < 1e-5
    # identical inputs :973-1064: subtraction ~0, ratio ~1 where there is signal, log-ratio ~0
    assert all(abs(x) < 1e-5 for x in O.compare(a, a, 1000, 10, "subtraction"))
    assert all(abs(x - 1.0) < 1e-5 for x in O.compare(a, a, 1000, 10, "ratio")[:20])
    assert all(abs(x) < 1e-5 for x in O.compare(a, a, 1000, 10, "log-ratio"))
    # zeros in the denominator stay finite with a pseudocount :1066-1141
    for mode in ("ratio", "log-ratio"):
        assert all(math.isfinite(x) for x in O.compare([(0, 100, 10.0)], [(0, 100, 0.0)], 200, 10, mode, pseudocount=1e-3))


def test_oracle_aggregate_reference_vector():
    # test_aggregate_bigwigs_mean_preserves_first_input_chromosome_order :1196-1248 (chr2: 2 and 6 -> 4)
    assert O.aggregate([[(0, 100, 2.0)], [(0, 100, 6.0)]], 100, 100, "mean") == [f32(4.0)]
    assert O.aggregate([[(0, 100, 8.0)], [(0, 100, 12.0)]], 100, 100, "sum") == [f32(20.0)]


# ---- inputs ------------------------------------------------------------------------------------------------------------------
def random_track(rng, chrom_len, density=0.02, zero_frac=0.1, ints=False):
    """Sorted, disjoint (start, end, value) intervals with gaps, abutting runs, exact zeros and a piece that ends at chrom_len."""
    ivs, pos = [], rng.randrange(0, 40)
    while pos < chrom_len:
        ln = rng.choice([1, 2, 7, 10, 25, 50, 133, 400]) if rng.random() < 0.9 else rng.randrange(1, 3000)
        end = min(pos + ln, chrom_len)
        r = rng.random()
        v = 0.0 if r < zero_frac else (float(rng.randrange(1, 40)) if ints or r < 0.5 else float(f32(rng.uniform(-3.0, 90.0))))
        ivs.append((pos, end, v))
        pos = end if rng.random() < 0.6 else end + int(rng.expovariate(density))
    return ivs


def write_track(path, chroms, per_chrom, flavour, rng):
    """chroms: [(name, id, length)]; per_chrom: {name: intervals}.  flavour: 'zlib' | 'raw' | 'mixed-types'."""
    sections = []
    for name, cid, _ in sorted(chroms, key=lambda c: c[1]):
        ivs = per_chrom.get(name, [])
        k = 0
        while k < len(ivs):
            n = rng.choice([1, 3, 64, 512, 1024])
            chunk = ivs[k:k + n]
            k += n
            typ = 1
            if flavour == "mixed-types":
                spans = {e - s for s, e, _ in chunk}
                if len(spans) == 1:
                    steps = {chunk[i + 1][0] - chunk[i][0] for i in range(len(chunk) - 1)}
                    typ = 3 if len(steps) == 1 and len(chunk) > 1 and rng.random() < 0.7 else 2
            if typ == 1:
                sections.append(dict(chrom_id=cid, type=1, items=chunk))
            elif typ == 2:
                sections.append(dict(chrom_id=cid, type=2, span=chunk[0][1] - chunk[0][0], items=[(s, v) for s, _, v in chunk]))
            else:
                sections.append(dict(chrom_id=cid, type=3, start=chunk[0][0], step=chunk[1][0] - chunk[0][0], span=chunk[0][1] - chunk[0][0],
                                     items=[v for _, _, v in chunk]))
    write_bigwig(path, chroms, sections, compress=flavour != "raw")


def read_per_chrom(path):
    bw = read_bigwig(path)
    by_id = {cid: name for name, (cid, _) in bw["chroms"].items()}
    out = {name: [] for name in bw["chroms"]}
    for cid, s, e, v in bw["intervals"]:
        out[by_id[cid]].append((s, e, f32(v)))
    return bw, out


def same_float(a, b):
    return (a == b) or (math.isnan(a) and math.isnan(b))


CHROMS = [("chrB", 0, 20_011), ("chrA", 1, 7_500), ("chrEmpty", 2, 900), ("chrTiny", 3, 7)]


def make_inputs(tmp_path, seed, n_files, flavours):
    rng = random.Random(seed)
    tracks, paths = [], []
    for k in range(n_files):
        per = {name: random_track(rng, ln, ints=(k == 0)) for name, _, ln in CHROMS if name != "chrEmpty"}
        # other files list their chromosomes under other ids and in another order: names decide, not ids
        chroms = CHROMS if k == 0 else [(n, (i + k) % len(CHROMS), ln) for (n, i, ln) in CHROMS]
        p = tmp_path / f"in{k}.bw"
        write_track(p, chroms, per, flavours[k % len(flavours)], rng)
        tracks.append(per)
        paths.append(str(p))
    return tracks, paths


# ---- device path against the oracle ------------------------------------------------------------------------------------------
@pytest.mark.parametrize("comparison", ["subtraction", "ratio", "log-ratio"])
@pytest.mark.parametrize("bin_size,opts", [(50, {}), (7, dict(pseudocount=0.5, scale_factor_bw1=1.75, scale_factor_bw2=0.3)), (1000, dict(pseudocount=1.0))])
def test_compare_matches_oracle(backend, tmp_path, comparison, bin_size, opts):
    tracks, paths = make_inputs(tmp_path, 1234 + bin_size, 2, ["zlib", "mixed-types"])
    out = tmp_path / "out.bw"
    backend.bigwig_compare(paths[0], paths[1], out, comparison, bin_size, **opts)
    bw, got = read_per_chrom(out)
    assert [n for n, _ in sorted(bw["chroms"].items(), key=lambda kv: kv[1][0])] == [c[0] for c in CHROMS]  # first input's order
    for name, _, ln in CHROMS:
        bins = O.compare(tracks[0].get(name, []), tracks[1].get(name, []), ln, bin_size, comparison, opts.get("pseudocount"), opts.get("scale_factor_bw1"),
                         opts.get("scale_factor_bw2"))
        if comparison != "log-ratio":
            want = O.merge_bins(bins, ln, bin_size)
            assert len(got[name]) == len(want), name
            for g, w in zip(got[name], want):
                assert g[:2] == w[:2] and same_float(g[2], w[2]), (name, g, w)
        else:  # device log vs libm: compare bin by bin
            dense = [None] * len(bins)
            for s, e, v in got[name]:
                for b in range(s // bin_size, -(-e // bin_size)):
                    dense[b] = v
            for b, (g, w) in enumerate(zip(dense, bins)):
                assert g is not None
                assert same_float(g, w) or abs(float(g) - float(w)) <= 1e-6 * max(1.0, abs(float(w))), (name, b, g, w)


@pytest.mark.parametrize("method", ["sum", "mean", "max", "min", "median"])
@pytest.mark.parametrize("bin_size,opts", [(50, {}), (13, dict(pseudocount=0.25, scale_factors=[1.0, 0.5, 3.0])), (2048, dict(scale_factors=[2.0, 2.0, 0.125]))])
def test_aggregate_matches_oracle(backend, tmp_path, method, bin_size, opts):
    tracks, paths = make_inputs(tmp_path, 99 + bin_size, 3, ["zlib", "raw", "mixed-types"])
    out = tmp_path / "out.bw"
    backend.bigwig_aggregate(paths, out, method, bin_size, **opts)
    _, got = read_per_chrom(out)
    for name, _, ln in CHROMS:
        bins = O.aggregate([t.get(name, []) for t in tracks], ln, bin_size, method, opts.get("pseudocount"), opts.get("scale_factors"))
        want = O.merge_bins(bins, ln, bin_size)
        assert len(got[name]) == len(want), (name, len(got[name]), len(want))
        for g, w in zip(got[name], want):
            assert g[:2] == w[:2] and same_float(g[2], w[2]), (name, g, w)


def test_reference_vectors_through_the_library(backend, tmp_path):
    def bw(name, chroms, rows):
        p = tmp_path / name
        backend.bigwig_from_intervals(chroms, [r[0] for r in rows], [r[1] for r in rows], [r[2] for r in rows], [r[3] for r in rows], p)
        return str(p)

    one = [("chr1", 1000)]
    a = bw("a.bw", one, [(0, 0, 100, 10.0), (0, 100, 200, 20.0)])
    b = bw("b.bw", one, [(0, 0, 100, 5.0), (0, 100, 200, 10.0)])
    out = tmp_path / "o.bw"
    backend.bigwig_compare(a, b, out, "subtraction", 10)  # :864-923
    _, got = read_per_chrom(out)
    assert got["chr1"] == [(0, 100, f32(5)), (100, 200, f32(10)), (200, 1000, f32(0))]
    backend.bigwig_compare(a, a, out, "log-ratio", 10)  # identical inputs :973-1064
    assert all(abs(v) < 1e-5 for _, _, v in read_per_chrom(out)[1]["chr1"])
    c = bw("c.bw", [("chr1", 200)], [(0, 0, 100, 10.0)])
    z = bw("z.bw", [("chr1", 200)], [(0, 0, 100, 0.0)])
    for mode in ("ratio", "log-ratio"):  # :1066-1141
        backend.bigwig_compare(c, z, out, mode, 10, pseudocount=1e-3)
        assert all(math.isfinite(v) for _, _, v in read_per_chrom(out)[1]["chr1"])
    # chromosome order of the first input is kept (:1143-1248): chr2 carries id 0
    two = [("chr2", 100), ("chr1", 100)]
    p = bw("p.bw", two, [(0, 0, 100, 2.0), (1, 0, 100, 8.0)])
    q = bw("q.bw", two, [(0, 0, 100, 6.0), (1, 0, 100, 12.0)])
    backend.bigwig_aggregate([p, q], out, "mean", 100)
    info, got = read_per_chrom(out)
    assert info["chroms"] == {"chr2": (0, 100), "chr1": (1, 100)}
    assert got["chr2"] == [(0, 100, f32(4))] and got["chr1"] == [(0, 100, f32(10))]


def test_tool_errors(backend, tmp_path):
    p = tmp_path / "a.bw"
    backend.bigwig_from_intervals([("chr1", 100)], [0], [0], [10], [1.0], p)
    q = tmp_path / "b.bw"
    backend.bigwig_from_intervals([("chrX", 100)], [0], [0], [10], [1.0], q)
    with pytest.raises(RuntimeError, match="chr1"):  # bigtools: the chromosome is not in the second file
        backend.bigwig_compare(p, q, tmp_path / "o.bw", "ratio", 10)
    with pytest.raises((RuntimeError, FileNotFoundError, ValueError)):
        backend.bigwig_compare(p, tmp_path / "missing.bw", tmp_path / "o.bw", "ratio", 10)
    with pytest.raises((RuntimeError, ValueError)):
        backend.bigwig_aggregate([str(p)], tmp_path / "o.bw", "sum", 0)
    (tmp_path / "junk.bw").write_bytes(b"\0" * 200)
    with pytest.raises(RuntimeError, match="BigWig"):
        backend.bigwig_aggregate([str(tmp_path / "junk.bw")], tmp_path / "o.bw", "sum", 10)


def test_cli_bigwig_tools(backend, tmp_path, capfd):
    a, b = tmp_path / "a.bw", tmp_path / "b.bw"
    backend.bigwig_from_intervals([("chr1", 1000)], [0, 0], [0, 100], [100, 200], [10.0, 20.0], a)
    backend.bigwig_from_intervals([("chr1", 1000)], [0, 0], [0, 100], [100, 200], [5.0, 10.0], b)
    out = tmp_path / "o.bw"
    assert backend.cli_main(["bamnado", "bigwig-compare", "--bw1", str(a), "--bw2", str(b), "-o", str(out), "-c", "ratio", "-s", "10"]) == 0
    assert read_per_chrom(out)[1]["chr1"][0] == (0, 200, f32(2))
    assert backend.cli_main(["bamnado", "bigwig-aggregate", "--bigwigs", str(a), str(b), "-o", str(out), "-m", "max", "--scale-factors", "1", "4"]) == 0
    assert read_per_chrom(out)[1]["chr1"][:2] == [(0, 100, f32(20)), (100, 200, f32(40))]
    assert backend.cli_main(["bamnado", "bigwig-aggregate", "--bigwigs", str(a), "-o", str(out), "-m", "mode"]) == 2
    assert backend.cli_main(["bamnado", "bigwig-aggregate", "--bigwigs", str(a), str(b), "-o", str(out), "-m", "sum", "--scale-factors", "1"]) == 1
    capfd.readouterr()


def test_tools_on_coverage_tracks(backend, tmp_path, make_bam):
    # end to end: two coverage tracks written by the device BigWig encoder (bam-coverage -o x.bw), then compared and aggregated
    bams = [make_bam("bwt_a", "--pairs", 4000, "--seed", 31, "--contigs", "chr1:300000,chr2:120000"),
            make_bam("bwt_b", "--pairs", 6000, "--seed", 32, "--contigs", "chr1:300000,chr2:120000")]
    tracks, paths, chroms = [], [], None
    for k, bam in enumerate(bams):
        p = tmp_path / f"cov{k}.bw"
        with backend.pileup(dict(bam_path=bam, bin_size=20, norm="cpm"), dict(min_mapq=10)) as h:
            h.to_bigwig(p)
        info, per = read_per_chrom(p)
        chroms = sorted(((n, cid, ln) for n, (cid, ln) in info["chroms"].items()), key=lambda c: c[1])
        tracks.append({n: [(s, e, float(v)) for s, e, v in ivs] for n, ivs in per.items()})
        paths.append(str(p))
    out = tmp_path / "cmp.bw"
    backend.bigwig_compare(paths[0], paths[1], out, "subtraction", 50, scale_factor_bw2=0.5)
    _, got = read_per_chrom(out)
    for name, _, ln in chroms:
        want = O.merge_bins(O.compare(tracks[0][name], tracks[1][name], ln, 50, "subtraction", scale2=0.5), ln, 50)
        assert len(got[name]) == len(want) > 10
        assert all(g[:2] == w[:2] and same_float(g[2], w[2]) for g, w in zip(got[name], want)), name
    backend.bigwig_aggregate(paths, out, "mean", 100)
    _, got = read_per_chrom(out)
    for name, _, ln in chroms:
        want = O.merge_bins(O.aggregate([t[name] for t in tracks], ln, 100, "mean"), ln, 100)
        assert len(got[name]) == len(want) > 10
        assert all(g[:2] == w[:2] and same_float(g[2], w[2]) for g, w in zip(got[name], want)), name
