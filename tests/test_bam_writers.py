"""split / modify / split-exogenous (bamnado/src/bam_splitter.rs:54-120, bam_modifier.rs:55-230,
spike_in_analysis.rs:233-400) and the device BGZF compressor they write with.

Parity target: the BAM header (the input's text plus the bamnado @PG record) and the sequence of record bytes of every output
file, against the oracle's restatement (oracle/bam_writers.inl, zlib).  The compressed bytes themselves are encoder-specific:
what is checked about them is that zlib inflates every block to the right bytes with the right CRC-32 / ISIZE, that the file ends
with the EOF marker, and that the library's own inflate kernel reads them back."""
from __future__ import annotations

import gzip
import json
import random
import struct
import zlib

import pytest

from conftest import TEST_BAM

EOF_BLOCK = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


def bgzf_blocks(b: bytes):
    """[(raw bytes)] of every BGZF block, each checked: gzip member with a BC field, CRC-32 and ISIZE of the inflated bytes."""
    off, out = 0, []
    while off < len(b):
        assert b[off:off + 4] == b"\x1f\x8b\x08\x04" and b[off + 12:off + 16] == b"BC\x02\x00", off
        bsize = struct.unpack_from("<H", b, off + 16)[0] + 1
        blk = b[off:off + bsize]
        raw = zlib.decompress(blk[18:-8], -15)
        crc, isize = struct.unpack_from("<II", blk, bsize - 8)
        assert isize == len(raw) <= 65536 and zlib.crc32(raw) == crc, off
        out.append(raw)
        off += bsize
    return out


def read_bam(path):
    """-> (header text, [(name, length)], [record bytes incl. block_size]); the file must end with the EOF marker."""
    data = open(path, "rb").read()
    assert data.endswith(EOF_BLOCK), "no BGZF EOF marker"
    d = b"".join(bgzf_blocks(data))
    assert d[:4] == b"BAM\x01"
    l_text = struct.unpack_from("<I", d, 4)[0]
    text = d[8:8 + l_text].decode()
    p = 8 + l_text
    n_ref = struct.unpack_from("<I", d, p)[0]
    p += 4
    refs = []
    for _ in range(n_ref):
        ln = struct.unpack_from("<I", d, p)[0]
        refs.append((d[p + 4:p + 4 + ln - 1].decode(), struct.unpack_from("<I", d, p + 4 + ln)[0]))
        p += 8 + ln
    recs = []
    while p < len(d):
        bs = struct.unpack_from("<I", d, p)[0]
        recs.append(d[p:p + 4 + bs])
        p += 4 + bs
    assert p == len(d)
    return text, refs, recs


def assert_same_bam(got_path, want_path):
    gt, gr, grec = read_bam(got_path)
    wt, wr, wrec = read_bam(want_path)
    assert gt == wt
    assert gr == wr
    assert len(grec) == len(wrec), (len(grec), len(wrec))
    for i, (a, b) in enumerate(zip(grec, wrec)):
        assert a == b, f"record {i} differs"
    return len(grec)


# ---- the compressor on its own ---------------------------------------------------------------------------------------------------
def test_bgzf_compress_round_trips(backend):
    rng = random.Random(7)
    bam = gzip.open(TEST_BAM).read()
    cases = {
        "empty": b"", "one": b"a", "tiny": b"abc" * 5, "zeros": bytes(70000), "one block exactly": bytes(rng.choice(b"ACGT") for _ in range(0xff00)),
        "random (stored blocks)": rng.randbytes(70000), "two symbols": bytes(rng.choice(b"\x00\xff") for _ in range(3000)),
        "single symbol": b"\x07" * 1000, "text": open(__file__, "rb").read(), "bam records": bam[:300000],
        "long runs": b"".join(bytes([rng.randrange(256)]) * rng.randrange(1, 700) for _ in range(300)),
    }
    for name, data in cases.items():
        c = backend.bgzf_compress(data)
        assert c.endswith(EOF_BLOCK), name
        blocks = bgzf_blocks(c)
        assert b"".join(blocks) == data, name
        assert all(len(r) == 0xff00 for r in blocks[:-2]), name  # full blocks, a tail, the empty EOF block
        assert len(c) <= len(data) + 31 * len(blocks) + 28, name   # never worse than stored
        back, status = backend.bgzf_inflate(c)  # the library's own inflate kernel reads what its compressor wrote
        assert bytes(back) == data and not status.any(), name
    assert len(backend.bgzf_compress(bytes(70000))) < 600
    assert len(backend.bgzf_compress(cases["text"])) < 0.75 * len(cases["text"])


# ---- split -----------------------------------------------------------------------------------------------------------------------
def test_split_writes_the_bamnado_program_group(backend, oracle, tmp_path):
    # bam_splitter.rs:203-236: BamReadFilter::default() on test.bam; the output header carries @PG ID:bamnado PN:bamnado VN:<version>
    out, want = tmp_path / "filtered.bam", tmp_path / "oracle.bam"
    filt = dict(proper_pair=True, min_mapq=0, min_length=0, max_length=1000)
    st = backend.bam_split(TEST_BAM, out, filt, command_line="bamnado split -i test.bam -o filtered.bam")
    n = oracle.bam_split(TEST_BAM, want, filt, command_line="bamnado split -i test.bam -o filtered.bam")
    text, refs, recs = read_bam(out)
    pg = [ln.split("\t") for ln in text.splitlines() if ln.startswith("@PG\tID:bamnado\t")]
    assert len(pg) == 1 and "PN:bamnado" in pg[0] and "VN:0.9.0" in pg[0] and "CL:bamnado split -i test.bam -o filtered.bam" in pg[0]
    assert len(refs) == 456
    assert assert_same_bam(out, want) == n == st["n_written"] > 10000
    assert st["n_total_reads"] == 11545 and st["n_errors"] == 0


def draw_filter(rng, tmp_path, contigs):
    filt = dict(strand=rng.choice(["both", "both", "forward", "reverse"]), proper_pair=rng.random() < 0.5, min_mapq=rng.choice([0, 1, 20, 30, 60]),
                min_length=rng.choice([0, 20, 101]), max_length=rng.choice([100, 1000, 0xFFFFFFFF]), exclude_duplicates=rng.random() < 0.4,
                exclude_secondary=rng.random() < 0.4, exclude_supplementary=rng.random() < 0.4)
    if rng.random() < 0.3:
        filt["min_fragment_length"] = rng.randrange(0, 300)
    if rng.random() < 0.3:
        filt["max_fragment_length"] = rng.randrange(100, 700)
    if rng.random() < 0.4:
        bed = tmp_path / f"bl{rng.randrange(1 << 30)}.bed"
        with open(bed, "w") as f:
            for _ in range(30):
                name, ln = rng.choice(contigs)
                s = rng.randrange(0, ln - 10)
                f.write(f"{name}\t{s}\t{min(ln, s + rng.randrange(10, 40000))}\tx\n")
        filt["blacklist_bed"] = str(bed)
        filt["blacklist_merge_overlaps"] = True
    return filt


CONTIGS = [("chr1", 1300000), ("chr2", 400001), ("chrUn_a", 60000), ("spike_x", 90000), ("chrM", 16569)]
CONTIG_ARG = ",".join(f"{n}:{ln}" for n, ln in CONTIGS)


def test_split_and_modify_match_oracle(backend, oracle, tmp_path, make_bam, monkeypatch):
    bam = make_bam("w_pe", "--pairs", 9000, "--seed", 5, "--odd", "--unplaced", 300, "--contigs", CONTIG_ARG)
    rng = random.Random(99)
    for i in range(8):
        # small segments, tiny output streams: many segments per pass, a flush (compress + write + carried tail) every few of them
        monkeypatch.setenv("BND_SEGMENT_BYTES", str(rng.choice([65536, 262144, 1 << 20])))
        monkeypatch.setenv("BND_BAM_STREAM_BYTES", str(rng.choice([1 << 20, 3 << 20, 2 << 30])))
        filt = draw_filter(rng, tmp_path, CONTIGS)
        tn5 = rng.random() < 0.6
        try:
            got, want = tmp_path / f"s{i}.bam", tmp_path / f"s{i}.oracle.bam"
            st = backend.bam_split(bam, got, filt, command_line=f"run {i}")
            assert oracle.bam_split(bam, want, filt, command_line=f"run {i}") == assert_same_bam(got, want) == st["n_written"]
            got, want = tmp_path / f"m{i}.bam", tmp_path / f"m{i}.oracle.bam"
            st = backend.bam_modify(bam, got, filt, tn5_shift=tn5)
            assert oracle.bam_modify(bam, want, filt, tn5_shift=tn5) == assert_same_bam(got, want) == st["n_written"]
            assert (st["n_dropped_by_shift"] > 0) == (tn5 and not filt["proper_pair"] and st["n_written"] > 50)  # reads outside proper pairs are not written shifted
        except AssertionError as e:
            raise AssertionError(f"configuration {i}: {filt} tn5={tn5}") from e


def test_modify_tn5_shift_rewrites_the_fields(backend, tmp_path, make_bam):
    # bam_modifier.rs:103-219: forward reads start 4 later, |tlen| shrinks by 9, a reverse read's mate start moves by +4, bin follows
    bam = make_bam("w_tn5", "--pairs", 3000, "--seed", 8, "--contigs", "chr1:500000,chr2:200000")
    plain, shifted = tmp_path / "p.bam", tmp_path / "t.bam"
    backend.bam_modify(bam, plain, dict(proper_pair=True), tn5_shift=False)
    st = backend.bam_modify(bam, shifted, dict(proper_pair=True), tn5_shift=True)
    _, _, a = read_bam(plain)
    _, _, b = read_bam(shifted)
    assert st["n_dropped_by_shift"] == len(a) - len(b) and len(b) > 2000
    by_key = {r[36:36 + r[12]] + bytes([r[18] & 0xC0]): r for r in a}  # name + first/second-in-pair bits
    for r in b:
        o = by_key[r[36:36 + r[12]] + bytes([r[18] & 0xC0])]
        flag = struct.unpack_from("<H", o, 18)[0]
        pos, mpos, tlen = struct.unpack_from("<i", o, 8)[0], struct.unpack_from("<i", o, 28)[0], struct.unpack_from("<i", o, 32)[0]
        npos, nmpos, ntlen = struct.unpack_from("<i", r, 8)[0], struct.unpack_from("<i", r, 28)[0], struct.unpack_from("<i", r, 32)[0]
        rev = bool(flag & 0x10)
        assert npos == (pos if rev else pos + 4)
        assert ntlen == (tlen + 9 if tlen < 0 else tlen - 9)
        assert nmpos == (mpos + 4 if rev else mpos)
        assert r[:8] == o[:8] and r[12:14] == o[12:14] and r[16:28] == o[16:28] and r[36:] == o[36:]


# ---- split-exogenous -------------------------------------------------------------------------------------------------------------
def test_split_exogenous_matches_oracle(backend, oracle, tmp_path, make_bam, monkeypatch):
    monkeypatch.setenv("BND_SEGMENT_BYTES", "262144")
    monkeypatch.setenv("BND_BAM_STREAM_BYTES", str(1 << 20))
    bam = make_bam("w_pe", "--pairs", 9000, "--seed", 5, "--odd", "--unplaced", 300, "--contigs", CONTIG_ARG)
    names = ["endogenous", "exogenous", "both", "unmapped"]
    for i, filt in enumerate([dict(), dict(min_mapq=30, proper_pair=True), dict(strand="forward", exclude_supplementary=True, max_length=100)]):
        prefix = tmp_path / f"x{i}" / "sample.tmp"  # with_extension replaces ".tmp"
        prefix.parent.mkdir()
        st = backend.bam_split_exogenous(bam, prefix, "spike_", filt)
        want = [tmp_path / f"x{i}" / f"oracle.{n}.bam" for n in names]
        ost = oracle.bam_split_exogenous(bam, want, "spike_", filt)
        assert {k: st[k] for k in ost} == ost, i
        assert st["n_total_reads"] == 18300 and st["n_exogenous"] > 0 and st["n_unmapped_reads"] >= 300 and st["n_both_genomes"] == 0
        counts = [assert_same_bam(tmp_path / f"x{i}" / f"sample.{n}.bam", w) for n, w in zip(names, want)]
        assert counts == [st["n_endogenous"], st["n_exogenous"], 0,
                          st["n_unmapped_reads"] + st["n_qcfail_reads"] + st["n_duplicate_reads"] + st["n_secondary_reads"]]


def test_split_exogenous_on_reference_bam(backend, oracle, tmp_path):
    st = backend.bam_split_exogenous(TEST_BAM, tmp_path / "t", "chrUn", dict(min_mapq=20))
    want = [tmp_path / f"o.{n}.bam" for n in ("endogenous", "exogenous", "both", "unmapped")]
    ost = oracle.bam_split_exogenous(TEST_BAM, want, "chrUn", dict(min_mapq=20))
    assert {k: st[k] for k in ost} == ost and st["n_total_reads"] == 11545
    for n, w in zip(("endogenous", "exogenous", "both", "unmapped"), want):
        assert_same_bam(tmp_path / f"t.{n}.bam", w)


def test_writer_errors(backend, tmp_path, make_bam):
    with pytest.raises(RuntimeError):
        backend.bam_split(tmp_path / "missing.bam", tmp_path / "o.bam")
    with pytest.raises(RuntimeError):
        backend.bam_split(TEST_BAM, tmp_path / "no_such_dir" / "o.bam")


def test_cli_bam_writers(backend, oracle, tmp_path, make_bam, capfd):
    bam = make_bam("w_cli", "--pairs", 2000, "--seed", 3, "--unplaced", 50, "--contigs", "chr1:300000,spike_a:50000")
    out = tmp_path / "s.bam"
    assert backend.cli_main(["bamnado", "split", "-i", bam, "-o", str(out), "--min-mapq", "30", "--proper-pair"]) == 0
    want = tmp_path / "s.oracle.bam"
    oracle.bam_split(bam, want, dict(min_mapq=30, proper_pair=True, min_length=20, max_length=1000))
    text, _, recs = read_bam(out)
    assert recs == read_bam(want)[2] and "\tCL:bamnado split -i " in text
    assert backend.cli_main(["bamnado", "modify", "-i", bam, "-o", str(tmp_path / "m.bam"), "--tn5-shift"]) == 0
    oracle.bam_modify(bam, want, dict(min_mapq=20, min_length=20, max_length=1000), tn5_shift=True)
    assert read_bam(tmp_path / "m.bam")[2] == read_bam(want)[2]
    stats = tmp_path / "stats.json"
    assert backend.cli_main(["bamnado", "split-exogenous", "-i", bam, "-o", str(tmp_path / "sp"), "-e", "spike_", "-s", str(stats)]) == 0
    js = json.loads(stats.read_text())
    assert js["filename"] == "w_cli.bam" and js["n_total_reads"] == 4050 and js["n_exogenous"] > 0
    assert js["spike_in_factor"] == 1.0 / (js["n_exogenous"] / 1e6) and js["spike_in_percentage"] == js["n_exogenous"] / js["n_endogenous"] * 100.0
    assert len(read_bam(tmp_path / "sp.exogenous.bam")[2]) == js["n_exogenous"]
    assert backend.cli_main(["bamnado", "split", "-o", str(out)]) == 2
    assert backend.cli_main(["bamnado", "split", "-i", str(tmp_path / "nope.bam"), "-o", str(out)]) == 1
    capfd.readouterr()


def test_long_records_span_bgzf_blocks(backend, oracle, tmp_path, make_bam, monkeypatch):
    # records of ~75 KB: every record is longer than a BGZF block (65 280 bytes), most blocks hold no record start at all, the
    # compressor parses them in pieces without a record lag; segments of 256 KiB make nearly every record straddle two segments
    monkeypatch.setenv("BND_SEGMENT_BYTES", "262144")
    monkeypatch.setenv("BND_BAM_STREAM_BYTES", str(1 << 20))
    bam = make_bam("w_long", "--pairs", 60, "--seed", 21, "--read-len", 50000, "--contigs", "chr1:3000000,spike_x:900000")
    got, want = tmp_path / "l.bam", tmp_path / "l.oracle.bam"
    st = backend.bam_split(bam, got, dict(min_mapq=1))
    assert oracle.bam_split(bam, want, dict(min_mapq=1)) == assert_same_bam(got, want) == st["n_written"] > 20
    assert max(len(r) for r in read_bam(got)[2]) > 0xff00
    st = backend.bam_split_exogenous(bam, tmp_path / "lx", "spike_", dict())
    paths = [tmp_path / f"lo.{n}.bam" for n in ("endogenous", "exogenous", "both", "unmapped")]
    ost = oracle.bam_split_exogenous(bam, paths, "spike_", dict())
    assert {k: st[k] for k in ost} == ost
    for n, w in zip(("endogenous", "exogenous", "both", "unmapped"), paths):
        assert_same_bam(tmp_path / f"lx.{n}.bam", w)


def test_program_group_links_to_the_leaf_of_the_chain(backend, oracle, tmp_path, make_bam):
    # add_bamnado_program_group (bam_utils.rs:757-780): the new @PG record has id `bamnado`, PN, VN and CL; with programs already in
    # the header noodles links it to the leaf of every chain (PP).  One chain (the synthetic BAM: synth_bam) -> one record;
    # test.bam has two leaves (MarkDuplicates and samtools.7) -> `bamnado` and `bamnado-<leaf>`
    bam = make_bam("w_pg", "--pairs", 200, "--seed", 2, "--contigs", "chr1:100000")
    got, want = tmp_path / "g.bam", tmp_path / "w.bam"
    backend.bam_split(bam, got, dict(), command_line="bamnado split -i in.bam -o out.bam")
    oracle.bam_split(bam, want, dict(), command_line="bamnado split -i in.bam -o out.bam")
    text = read_bam(got)[0]
    assert text == read_bam(want)[0]
    pg = [ln for ln in text.splitlines() if ln.startswith("@PG")]
    assert len(pg) == 2 and pg[1] == "@PG\tID:bamnado\tPN:bamnado\tVN:0.9.0\tCL:bamnado split -i in.bam -o out.bam\tPP:" + pg[0].split("\t")[1][3:]
    ref_lines = read_bam_text(TEST_BAM).splitlines()
    backend.bam_split(TEST_BAM, got, dict(), command_line="x")
    new = [ln for ln in read_bam(got)[0].splitlines() if ln not in ref_lines]
    assert [ln.split("\t")[1] for ln in new] == ["ID:bamnado", "ID:bamnado-samtools.7"]  # leaves in header order: MarkDuplicates, samtools.7
    assert [ln.split("\t")[-1] for ln in new] == ["PP:MarkDuplicates", "PP:samtools.7"]
    assert all("\tPN:bamnado\tVN:0.9.0\tCL:x\tPP:" in ln for ln in new)


def read_bam_text(path):
    d = gzip.open(path).read(1 << 20)
    return d[8:8 + struct.unpack_from("<I", d, 4)[0]].decode()
