"""Edge cases of the coverage path: empty inputs, a BAI without the metadata pseudo-bin (chunk length degenerates to the
bin size), tiny contigs, bins larger than contigs, deep pile-ups, and a larger file checked both against
the oracle and through size-independent properties."""
import struct
import zlib

import numpy as np
import pytest

CLI_FILTER = dict(min_mapq=20, min_length=20, max_length=1000)


def same(backend, oracle, pile, filt):
    with backend.pileup(pile, filt) as h:
        runs, stats, factor = h.runs(), h.stats(), h.norm_factor()
    oruns, ostats, ofactor = oracle.pileup_runs(pile, filt)
    assert stats == ostats and factor == ofactor
    assert len(runs) == len(oruns) and (runs == oruns).all()
    return runs, stats


def strip_bai_metadata(src, dst):
    """Rewrites a .bai without the 37450 pseudo-bins (and without the trailing n_no_coor)."""
    d = open(src, "rb").read()
    n_ref = struct.unpack_from("<I", d, 4)[0]
    p, out = 8, [d[:8]]
    for _ in range(n_ref):
        n_bin = struct.unpack_from("<I", d, p)[0]
        p += 4
        bins = []
        for _ in range(n_bin):
            b, n_chunk = struct.unpack_from("<II", d, p)
            raw = d[p: p + 8 + 16 * n_chunk]
            p += 8 + 16 * n_chunk
            if b != 37450:
                bins.append(raw)
        out.append(struct.pack("<I", len(bins)) + b"".join(bins))
        n_intv = struct.unpack_from("<I", d, p)[0]
        out.append(d[p: p + 4 + 8 * n_intv])
        p += 4 + 8 * n_intv
    open(dst, "wb").write(b"".join(out))


def test_small_genome_many_shapes(backend, oracle, make_bam):
    bam = make_bam("tiny", "--pairs", 6000, "--contigs", "c1:120000,c2:777,c3:50,c4:1,c5:300000,c6_random:4000", "--odd", "--peaks", 20)
    for bin_size in (1, 7, 50, 1000, 500000):
        for collapse in (True, False):
            same(backend, oracle, dict(bam_path=bam, bin_size=bin_size, collapse=collapse), CLI_FILTER)
    same(backend, oracle, dict(bam_path=bam, bin_size=25, use_fragment=True, norm="rpkm"), dict(proper_pair=True, **CLI_FILTER))


def test_bai_without_metadata_degenerate_chunks(backend, oracle, make_bam, tmp_path):
    # n_mapped = 0 -> chunk length = bin size -> one (bin - 1)-wide bin per chunk and reads counted once per chunk they touch
    src = make_bam("tiny", "--pairs", 6000, "--contigs", "c1:120000,c2:777,c3:50,c4:1,c5:300000,c6_random:4000", "--odd", "--peaks", 20)
    bam = tmp_path / "nometa.bam"
    bam.write_bytes(open(src, "rb").read())
    strip_bai_metadata(src + ".bai", str(bam) + ".bai")
    assert oracle.bam_stats(bam, 40)["chunk_len"] == 40
    for bin_size in (40, 3):
        runs, stats = same(backend, oracle, dict(bam_path=str(bam), bin_size=bin_size, norm="cpm"), CLI_FILTER)
        assert stats["n_total"] > 12000  # reads are visited once per (tiny) chunk they overlap


def test_header_only_bam(backend, oracle, make_bam):
    bam = make_bam("empty", "--pairs", 0, "--contigs", "c1:100000,c2:2000")
    runs, stats = same(backend, oracle, dict(bam_path=bam, bin_size=100, norm="cpm"), CLI_FILTER)
    assert stats["n_total"] == 0
    with backend.pileup(dict(bam_path=bam, bin_size=100, norm="cpm"), CLI_FILTER) as h:
        assert h.norm_factor() == 0.0  # n == 0 -> factor 0, an all-zero track, not an error
        text = h.bedgraph_text()
    assert text.count(b"\n") == len(runs)


def test_deep_pileup_hot_spot(backend, oracle, make_bam):
    # every read in a handful of peaks: shared-memory window + global atomics under heavy contention
    bam = make_bam("deep", "--pairs", 30000, "--contigs", "c1:5000000,c2:100000", "--peaks", 3)
    runs, _ = same(backend, oracle, dict(bam_path=bam, bin_size=10), dict())
    assert runs["val"].max() > 500


@pytest.mark.gpu
def test_two_million_pairs_properties_and_oracle(native_gpu, oracle, make_bam, tmp_path):
    """4 M records / 865 MB inflated: full comparison with the oracle plus size-independent properties
    (idempotence, resident == streaming, shards == whole, collapsed rows expand to the uncollapsed ones)."""
    N = native_gpu
    bam = make_bam("pe2m", "--pairs", 2000000)
    pile, filt = dict(bam_path=bam, bin_size=50, norm="rpkm"), dict(min_mapq=30, proper_pair=True)
    with N.pileup(pile, filt) as h:
        a, stats = h.runs(), h.stats()
        text1 = h.bedgraph_text()
        h.accumulate()
        h.finalize()
        assert (h.runs() == a).all() and h.bedgraph_text() == text1          # idempotent
        h.stage()
        h.accumulate_resident()
        h.finalize()
        assert (h.runs() == a).all() and h.stats() == stats                   # resident == streaming
    out = tmp_path / "o.bdg"
    ostats, _ = oracle.pileup_to_bedgraph(pile, filt, out)
    assert stats == ostats and text1 == out.read_bytes()                      # byte-identical bedGraph
    parts, total = [], None
    hs = [N.pileup(dict(pile, shard_rank=r, shard_world=5), filt) for r in range(5)]
    for h in hs:
        h.accumulate()
        s = h.stats()
        total = s if total is None else {k: total[k] + s[k] for k in s}
    for h in hs:
        h.set_total_stats(list(total.values()))
        h.finalize()
        parts.append(h.runs())
        h.close()
    assert total == stats and (np.concatenate(parts) == a).all()             # shards == whole
    with N.pileup(dict(pile, collapse=False), filt) as h:
        b = h.runs()
    # expanding collapsed runs bin by bin gives the uncollapsed rows (checksum of value x bins and row count)
    assert int(b["val"].sum()) == int((a["val"].astype(np.int64) * np.ceil((a["stop"] - a["start"]) / 50).astype(np.int64)).sum())
    heads = np.ones(len(b), bool)
    heads[1:] = (b["val"][1:] != b["val"][:-1]) | (b["ref_id"][1:] != b["ref_id"][:-1]) | (b["start"][1:] != b["stop"][:-1])
    assert heads.sum() == len(a)


def _bgzf(payload: bytes) -> bytes:
    co = zlib.compressobj(6, zlib.DEFLATED, -15)
    comp = co.compress(payload) + co.flush()
    return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", 18 + len(comp) + 8 - 1) + comp
            + struct.pack("<II", zlib.crc32(payload), len(payload)))


def write_adversarial_bam(path, n_records=3000, contig_len=400000, seed=1):
    """One contig; every record carries a B:C aux array filled with back-to-back byte patterns that look exactly like
    minimal BAM records, so most speculative record-start guesses of K2 land inside a fake chain."""
    rng = np.random.default_rng(seed)
    fake = struct.pack("<IiiBBHHHIiii", 33, 0, 5, 1, 0, 4680, 0, 4, 0, -1, -1, 0) + b"\0"  # 37 bytes: block_size 33 + payload
    text = b"@HD\tVN:1.6\tSO:coordinate\n@SQ\tSN:c1\tLN:%d\n" % contig_len
    stream = b"BAM\1" + struct.pack("<I", len(text)) + text + struct.pack("<I", 1) + struct.pack("<I", 3) + b"c1\0" + struct.pack("<I", contig_len)
    header_len = len(stream)
    pos = np.sort(rng.integers(0, contig_len - 200, n_records))
    recs = []
    for i, p in enumerate(pos):
        name = b"q%06d\0" % i
        l_seq = 50
        cigar = struct.pack("<I", (l_seq << 4) | 0)
        seq = bytes(rng.integers(0, 256, (l_seq + 1) // 2, dtype=np.uint8))
        qual = bytes(rng.integers(0, 40, l_seq, dtype=np.uint8))
        n_fake = int(rng.integers(5, 90))
        aux = b"XFBC" + struct.pack("<I", 37 * n_fake) + fake * n_fake + b"NMC\x01"
        body = struct.pack("<iiBBHHHIiii", 0, int(p), len(name), int(rng.integers(0, 61)), 4680, 1, 0, l_seq, -1, -1, 0) + name + cigar + seq + qual + aux
        recs.append(struct.pack("<I", len(body)) + body)
    stream += b"".join(recs)
    blocks, coffs, off = [], [], 0
    for i in range(0, len(stream), 65280):
        b = _bgzf(stream[i:i + 65280])
        coffs.append(off)
        off += len(b)
        blocks.append(b)
    eof = _bgzf(b"")
    open(path, "wb").write(b"".join(blocks) + eof)
    first_voff = (coffs[header_len // 65280] << 16) | (header_len % 65280)
    end_voff = off << 16
    n_intv = (contig_len + 16383) // 16384
    bai = b"BAI\1" + struct.pack("<I", 1) + struct.pack("<I", 2)
    bai += struct.pack("<II", 0, 1) + struct.pack("<QQ", first_voff, end_voff)
    bai += struct.pack("<II", 37450, 2) + struct.pack("<QQQQ", first_voff, end_voff, n_records, 0)
    bai += struct.pack("<I", n_intv) + struct.pack("<Q", first_voff) * n_intv + struct.pack("<Q", 0)
    open(str(path) + ".bai", "wb").write(bai)
    return pos


def test_record_discovery_survives_fake_record_chains(backend, oracle, tmp_path, small_segments):
    bam = tmp_path / "adversarial.bam"
    pos = write_adversarial_bam(bam)
    runs, stats = same(backend, oracle, dict(bam_path=str(bam), bin_size=100), dict())
    assert stats["n_total"] == len(pos)
    # independent check of the pile-up itself: every record covers [pos + 1, pos + 51)
    # (1-based half-open), and a bin [a, b) counts #{start < b} - #{stop <= a} (Lapper::count, bam_utils.rs:59); a run's
    # value is the count of its first bin
    starts, stops = np.sort(pos + 1), np.sort(pos + 51)
    a = runs["start"].astype(np.int64)
    b = np.minimum(a + 100, runs["stop"].astype(np.int64))
    expect = np.searchsorted(starts, b, side="left") - np.searchsorted(stops, a, side="right")
    assert (runs["val"].astype(np.int64) == expect).all()
    assert runs["val"].max() > 0 and int((runs["val"].astype(np.int64) > 0).sum()) > 100


def test_chromosome_query_leaves_whole_object_outputs_alone(backend, oracle, tmp_path, make_bam):
    # pileup_chromosome (coverage_analysis.rs:337-430) does not change what a later to_bedgraph writes: the handle must not
    # keep the single-contig accumulators as if they were the whole result
    bam = make_bam("pe", "--pairs", 20000, "--odd")
    pile, filt = dict(bam_path=bam, bin_size=50, norm="cpm"), CLI_FILTER
    ref_out = tmp_path / "oracle.bdg"
    ostats, _ = oracle.pileup_to_bedgraph(pile, filt, ref_out)
    with backend.pileup(pile, filt) as h:  # contig query first, whole output afterwards
        c2 = h.chromosome("chr2")
        assert len(c2) > 0 and (c2["ref_id"] == c2["ref_id"][0]).all()
        out = tmp_path / "after_query.bdg"
        h.to_bedgraph(out)
        assert h.stats() == ostats
        assert out.read_bytes() == ref_out.read_bytes()
    with backend.pileup(pile, filt) as h:  # whole result, contig query, whole result again
        whole = h.bedgraph_text()
        sig = h.chromosome("chr3")
        assert len(sig) > 0
        assert h.bedgraph_text() == whole == ref_out.read_bytes()
        assert h.stats() == ostats


@pytest.mark.parametrize("touchers", ["0", "3"])
def test_to_bedgraph_with_background_page_touching(backend, oracle, tmp_path, make_bam, monkeypatch, touchers):
    # to_bedgraph allocates and touches the output's pages behind the pile-up (engine.cpp: prealloc thread + touchers) and hands
    # that mapping to the writers; small outputs normally skip it, so the threshold is lowered here.  The surplus of the
    # estimate must be cut off and the bytes must be the oracle's.
    monkeypatch.setenv("BND_PREALLOC_MIN", "0")
    monkeypatch.setenv("BND_TOUCHERS", touchers)
    monkeypatch.setenv("BND_PRETOUCH", "1" if touchers != "0" else "0")
    bam = make_bam("pe", "--pairs", 20000, "--odd")
    pile, filt = dict(bam_path=bam, bin_size=50, norm="rpkm"), dict(min_mapq=30, proper_pair=True)
    ref, out = tmp_path / "ref.bdg", tmp_path / "out.bdg"
    oracle.pileup_to_bedgraph(pile, filt, ref)
    for _ in range(2):
        with backend.pileup(pile, filt) as h:
            h.to_bedgraph(out)
        assert out.read_bytes() == ref.read_bytes()


def test_bam_mapping_cache_follows_the_file(backend, oracle, tmp_path, make_bam):
    # the read-only mapping of the last BAM is kept for the next handle on the same file (engine.cpp: MapCache): a path whose
    # file is replaced, rewritten in place or opened twice at once must still give that file's own coverage
    import shutil

    a = make_bam("cache_a", "--pairs", 3000, "--seed", 11, "--contigs", "chr1:400000,chr2:100000")
    b = make_bam("cache_b", "--pairs", 4500, "--seed", 12, "--contigs", "chr1:400000,chr2:100000")
    p = tmp_path / "x.bam"

    def put(src, replace):
        if replace:  # new inode
            shutil.copy(src, str(p) + ".tmp")
            shutil.move(str(p) + ".tmp", p)
        else:  # same inode, rewritten and truncated in place
            data = open(src, "rb").read()
            with open(p, "r+b" if p.exists() else "wb") as f:
                f.write(data)
                f.truncate()
        shutil.copy(src + ".bai", str(p) + ".bai")

    pile, filt = dict(bam_path=str(p), bin_size=25), dict(min_mapq=10)
    for src, replace in [(a, True), (a, True), (b, True), (a, False), (b, False), (b, False)]:
        put(src, replace)
        same(backend, oracle, pile, filt)
    with backend.pileup(pile, filt) as h1, backend.pileup(pile, filt) as h2:  # two handles on one file at once
        r1, r2 = h1.runs(), h2.runs()
    assert (r1 == r2).all() and (r1 == oracle.pileup_runs(pile, filt)[0]).all()
