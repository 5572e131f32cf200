"""Parity of the CUDA path against the CPU oracle, through the C ABI.

Each test runs on the GPU (`-m gpu`, the gate) and — for development on boxes without a GPU — on the SIMT emulator
build of the same kernel sources (`emul` marker, part of the CPU suite).  Raw counts, intervals, filter counters
and bedGraph rows must be bit-exact; CPM/RPKM factors are compared exactly too (same f64 arithmetic), which is
tighter than the 1e-6 relative tolerance north_star allows.
"""
import os
import zlib

import numpy as np
import pytest

from conftest import TEST_BAM

CLI_FILTER = dict(min_mapq=20, min_length=20, max_length=1000)  # CLI defaults (src/cli.rs FilterOptions)
REL_TOL = 1e-6


def assert_same(backend, oracle, pile, filt, tmp_path, text=True):
    with backend.pileup(pile, filt) as h:
        runs = h.runs()
        stats, factor = h.stats(), h.norm_factor()
        txt = h.bedgraph_text() if text else None
    oruns, ostats, ofactor = oracle.pileup_runs(pile, filt)
    assert stats == ostats
    assert factor == pytest.approx(ofactor, rel=REL_TOL) and (pile.get("norm", "raw") != "raw" or factor == ofactor)
    assert len(runs) == len(oruns)
    assert (runs == oruns).all()
    if text:
        out = tmp_path / "oracle.bdg"
        oracle.pileup_to_bedgraph(pile, filt, out)
        a, b = txt.split(b"\n"), out.read_bytes().split(b"\n")
        assert len(a) == len(b)
        for x, y in zip(a, b):
            if x != y:  # coordinates exact, score within tolerance
                xs, ys = x.split(b"\t"), y.split(b"\t")
                assert xs[:3] == ys[:3] and float(xs[3]) == pytest.approx(float(ys[3]), rel=REL_TOL)
    return runs, stats


def test_config1_test_bam_bin100_raw(backend, oracle, tmp_path):
    runs, stats = assert_same(backend, oracle, dict(bam_path=TEST_BAM, bin_size=100, norm="raw"), CLI_FILTER, tmp_path)
    assert stats["n_total"] == 11545 and len(runs) == 11846 and runs["val"].max() == 27


@pytest.mark.parametrize("norm,collapse", [("raw", True), ("cpm", True), ("rpkm", False)])
def test_reference_known_answers(backend, oracle, tmp_path, norm, collapse):
    # coverage_analysis.rs:989-1094
    pile = dict(bam_path=TEST_BAM, bin_size=1000, norm=norm, collapse=collapse)
    filt = dict(proper_pair=True, max_length=1000)
    runs, _ = assert_same(backend, oracle, pile, filt, tmp_path, text=collapse)
    with backend.pileup(pile, filt) as h:
        h.accumulate()
        f = h.norm_factor()
    assert runs["val"].max() == 117
    if norm != "raw":
        assert float(np.float32(117 * f)) == 10134.2568359375


def test_many_segments_and_carry(backend, oracle, tmp_path, small_segments, make_bam):
    bam = make_bam("seg", "--pairs", 30000, "--odd")
    assert_same(backend, oracle, dict(bam_path=bam, bin_size=50, norm="rpkm"), dict(min_mapq=30, proper_pair=True), tmp_path)


def test_config2_flags(backend, oracle, tmp_path, make_bam):
    bam = make_bam("pe", "--pairs", 20000, "--odd")
    assert_same(backend, oracle, dict(bam_path=bam, bin_size=50, norm="rpkm", ignore_scaffold=True), dict(min_mapq=30, proper_pair=True), tmp_path)


def test_config3_fragments_blacklist(backend, oracle, tmp_path, make_bam):
    bam = make_bam("atac", "--pairs", 20000, "--mode", "atac", "--odd")
    bed = tmp_path / "bl.bed"
    rng = np.random.default_rng(7)
    lines = []
    for i in range(300):
        c = ["chr1", "chr2", "chr9", "chrX", "chr1_KI270706v1_random"][i % 5]
        s = int(rng.integers(0, 150000)) if c.endswith("random") else int(rng.integers(0, 50_000_000))
        lines.append(f"{c}\t{s}\t{s + int(rng.integers(200, 2_000_000))}\tbl{i}\n")
    bed.write_text("".join(lines))
    filt = dict(max_fragment_length=120, blacklist_bed=str(bed), blacklist_merge_overlaps=True, **CLI_FILTER)
    runs, stats = assert_same(backend, oracle, dict(bam_path=bam, bin_size=10, use_fragment=True), filt, tmp_path)
    assert stats["n_failed_blacklist"] > 0 and stats["n_failed_fragment_length"] > 0


def test_config4_barcodes_duplicates_cpm(backend, oracle, tmp_path, make_bam):
    bam = make_bam("cb", "--pairs", 20000, "--barcodes", "--n-barcodes", 500)
    # allowlist = every third barcode seen in the file + some that are absent
    import subprocess
    seen = set()
    raw = open(bam, "rb").read()
    off = 0
    import struct
    data = b""
    while off + 18 <= len(raw):
        bs = struct.unpack_from("<H", raw, off + 16)[0] + 1
        data += zlib.decompress(raw[off + 18: off + bs - 8], -15)
        off += bs
    i = 0
    while True:
        i = data.find(b"CBZ", i)
        if i < 0:
            break
        j = data.find(b"\0", i + 3)
        seen.add(data[i + 3: j].decode())
        i = j
    seen = sorted(seen)
    allow = seen[::3] + ["AAAACCCCGGGGTTTT-1", "NOTPRESENT-1"]
    csv = tmp_path / "bc.csv"
    csv.write_text("barcode,cluster\n" + "".join(f"{b},c{i % 4}\n" for i, b in enumerate(allow)))
    filt = dict(barcode_csv=str(csv), exclude_duplicates=True, **CLI_FILTER)
    runs, stats = assert_same(backend, oracle, dict(bam_path=bam, bin_size=50, norm="cpm"), filt, tmp_path)
    assert stats["n_failed_barcode"] > 0 and stats["n_failed_duplicate"] > 0
    filt2 = dict(barcodes=allow[:50], **CLI_FILTER)  # Python ReadFilter.whitelisted_barcodes path
    assert_same(backend, oracle, dict(bam_path=bam, bin_size=50), filt2, tmp_path, text=False)


def test_read_group_tag_strand_filters(backend, oracle, tmp_path, make_bam):
    bam = make_bam("rg", "--pairs", 15000, "--rg", "--odd")
    for filt in (dict(read_group="rgA"), dict(read_group="missing"), dict(filter_tag="VP", filter_tag_value="BCL2"),
                 dict(filter_tag="VP"), dict(filter_tag="VPX", filter_tag_value="BCL2"), dict(strand="forward"),
                 dict(strand="reverse", exclude_secondary=True, exclude_supplementary=True), dict(min_fragment_length=150, max_fragment_length=400)):
        assert_same(backend, oracle, dict(bam_path=bam, bin_size=200), dict(filt, **CLI_FILTER), tmp_path, text=False)


def test_shift_and_truncate(backend, oracle, tmp_path, make_bam):
    bam = make_bam("pe", "--pairs", 20000, "--odd")
    assert_same(backend, oracle, dict(bam_path=bam, bin_size=25, shift=[4, -5, 5, -4]), CLI_FILTER, tmp_path, text=False)
    assert_same(backend, oracle, dict(bam_path=bam, bin_size=25, truncate=[10, 30]), CLI_FILTER, tmp_path, text=False)
    assert_same(backend, oracle, dict(bam_path=bam, bin_size=25, truncate=[90, 90], use_fragment=True, shift=[0, 0, 1, 0]), CLI_FILTER, tmp_path, text=False)


def test_single_end_and_unplaced(backend, oracle, tmp_path, make_bam):
    bam = make_bam("se", "--pairs", 15000, "--mode", "se", "--unplaced", 500)
    assert_same(backend, oracle, dict(bam_path=bam, bin_size=1000, norm="cpm", collapse=False, ignore_scaffold=True), dict(), tmp_path)


def test_pileup_chromosome_and_errors(backend, oracle):
    pile, filt = dict(bam_path=TEST_BAM, bin_size=100), CLI_FILTER
    with backend.pileup(pile, filt) as h:
        runs = h.chromosome("chr9")
        stats = h.stats()
        with pytest.raises(KeyError):
            h.chromosome("chrNope")
    oruns, ostats = oracle.pileup_chromosome(pile, filt, "chr9")
    assert (runs == oruns).all() and stats == ostats
    with pytest.raises(RuntimeError):
        backend.pileup(dict(bam_path="/nonexistent/x.bam", bin_size=10), {})
    with pytest.raises(ValueError):
        backend.pileup(dict(bam_path=TEST_BAM, bin_size=0), {})


def test_signal_for_chromosome(backend, oracle, make_bam):
    # bamnado-python/src/lib.rs:248-304: dense f32[chrom_size], 1-based coordinates used as indices
    bam = make_bam("pe", "--pairs", 20000, "--odd")
    pile = dict(bam_path=bam, bin_size=37, scale_factor=1.5, use_fragment=False)
    for chrom in ("chr21", "chrM", "chr1_KI270706v1_random"):
        got = backend.signal_for_chromosome(pile, dict(proper_pair=True, max_length=1000), chrom)
        exp = oracle.signal_for_chromosome(pile, dict(proper_pair=True, max_length=1000), chrom)
        assert got.dtype == np.float32 and got.shape == exp.shape
        assert (got == exp).all()
    with pytest.raises(KeyError):
        backend.signal_for_chromosome(pile, {}, "nope")


@pytest.mark.parametrize("world", [2, 3, 4, 8])
def test_contig_sharded_union_equals_whole(backend, oracle, make_bam, world, small_segments):
    # §8(e): shards own disjoint chunk ranges; rows concatenate, counters add up (13 x u64 sum before CPM/RPKM)
    bam = make_bam("pe", "--pairs", 20000, "--odd")
    pile, filt = dict(bam_path=bam, bin_size=50, norm="cpm"), dict(min_mapq=30, proper_pair=True)
    oruns, ostats, ofactor = oracle.pileup_runs(pile, filt)
    parts, total = [], None
    handles = [backend.pileup(dict(pile, shard_rank=r, shard_world=world), filt) for r in range(world)]
    for h in handles:
        h.accumulate()
        s = h.stats()
        total = s if total is None else {k: total[k] + s[k] for k in s}
    assert total == ostats
    for h in handles:
        h.set_total_stats(list(total.values()))
        h.finalize()
        assert h.norm_factor() == ofactor
        parts.append(h.runs())
        h.close()
    runs = np.concatenate(parts)
    assert len(runs) == len(oruns) and (runs == oruns).all()


def test_resident_pass_equals_streaming(backend, make_bam, small_segments):
    bam = make_bam("pe", "--pairs", 20000, "--odd")
    with backend.pileup(dict(bam_path=bam, bin_size=50), CLI_FILTER) as h:
        a = h.runs()
        sa = h.stats()
        h.stage()
        h.accumulate_resident()
        h.finalize()
        assert (h.runs() == a).all() and h.stats() == sa


def _bgzf_block(payload: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, corrupt_crc=False) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    comp = co.compress(payload) + co.flush()
    crc = zlib.crc32(payload) ^ (1 if corrupt_crc else 0)
    bsize = 18 + len(comp) + 8 - 1
    import struct
    return b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", bsize) + comp + struct.pack("<II", crc, len(payload))


def test_inflate_kernel_block_types(backend):
    # K1 on its own against zlib: dynamic, fixed-Huffman and stored deflate blocks, long matches, empty block, bad CRC
    rng = np.random.default_rng(3)
    text = (b"ACGTTGCA" * 4000 + bytes(rng.integers(65, 70, 20000, dtype=np.uint8)))[:65000]
    rnd = bytes(rng.integers(0, 256, 60000, dtype=np.uint8))
    payloads = [text, rnd, b"", b"A" * 65280, bytes(range(256)) * 200, text[:1], rnd[:300] + b"\0" * 30000]
    blocks = [_bgzf_block(p, 6) for p in payloads]
    blocks.append(_bgzf_block(text, 1, zlib.Z_FIXED))
    blocks.append(_bgzf_block(text, 9, zlib.Z_HUFFMAN_ONLY))
    blocks.append(_bgzf_block(rnd, 0))
    expected = b"".join(payloads) + text + text + rnd
    out, status = backend.bgzf_inflate(b"".join(blocks))
    assert list(status) == [0] * len(blocks)
    assert out == expected
    out, status = backend.bgzf_inflate(_bgzf_block(text) + _bgzf_block(text, corrupt_crc=True))
    assert list(status) == [0, 9]  # INF_CRC_MISMATCH


def test_truncated_and_corrupt_inputs_fail_loudly(backend, tmp_path, make_bam):
    bam = make_bam("pe", "--pairs", 20000, "--odd")
    raw = bytearray(open(bam, "rb").read())
    bad = tmp_path / "bad.bam"
    raw[len(raw) // 2] ^= 0x55  # flip bits inside some BGZF payload
    bad.write_bytes(bytes(raw))
    (tmp_path / "bad.bam.bai").write_bytes(open(bam + ".bai", "rb").read())
    with pytest.raises(RuntimeError):
        with backend.pileup(dict(bam_path=str(bad), bin_size=50), {}) as h:
            h.accumulate()
