"""Minimal BigWig reader for the tests (bedGraph / varStep / fixedStep sections, zlib or raw)."""
from __future__ import annotations

import struct
import zlib


def read_bigwig(path):
    """-> dict(chroms={name: (id, size)}, intervals=[(chrom_id, start, end, value)], summary=dict, zoom_levels=int)"""
    d = open(path, "rb").read()
    magic, version, zooms, chrom_off, data_off, index_off, fc, dfc, sql_off, summ_off, ubuf, _ = struct.unpack_from("<IHHQQQHHQQIQ", d, 0)
    assert magic == 0x888FFC26, "not a little-endian BigWig"
    chroms = {}

    def walk_chrom(off, key_size):
        is_leaf, _, count = struct.unpack_from("<BBH", d, off)
        off += 4
        for _ in range(count):
            key = d[off:off + key_size].rstrip(b"\0").decode()
            if is_leaf:
                cid, size = struct.unpack_from("<II", d, off + key_size)
                chroms[key] = (cid, size)
            else:
                (child,) = struct.unpack_from("<Q", d, off + key_size)
                walk_chrom(child, key_size)
            off += key_size + 8

    tmagic, block, key_size, val_size, n_items, _ = struct.unpack_from("<IIIIQQ", d, chrom_off)
    assert tmagic == 0x78CA8C91
    if n_items:
        walk_chrom(chrom_off + 32, key_size)

    # libBigWig (pyBigWig, deepTools) indexes arrays of `n_items` entries by chromosome id: ids must be dense
    assert sorted(cid for cid, _ in chroms.values()) == list(range(len(chroms))), "chromosome ids are not dense"

    leaves = []

    def walk_r(off):
        is_leaf, _, count = struct.unpack_from("<BBH", d, off)
        off += 4
        for _ in range(count):
            if is_leaf:
                sc, sb, ec, eb, doff, dsize = struct.unpack_from("<IIIIQQ", d, off)
                leaves.append((doff, dsize))
                off += 32
            else:
                sc, sb, ec, eb, child = struct.unpack_from("<IIIIQ", d, off)
                walk_r(child)
                off += 24

    rmagic, rblock, ritems = struct.unpack_from("<IIQ", d, index_off)
    assert rmagic == 0x2468ACE0
    if ritems:
        walk_r(index_off + 48)
    intervals = []
    for doff, dsize in sorted(leaves):
        sec = d[doff:doff + dsize]
        if ubuf:
            sec = zlib.decompress(sec)
        cid, cstart, cend, step, span, typ, _, count = struct.unpack_from("<IIIIIBBH", sec, 0)
        p = 24
        for i in range(count):
            if typ == 1:
                s, e, v = struct.unpack_from("<IIf", sec, p)
                p += 12
            elif typ == 2:
                s, v = struct.unpack_from("<If", sec, p)
                e = s + span
                p += 8
            else:
                (v,) = struct.unpack_from("<f", sec, p)
                s = cstart + i * step
                e = s + span
                p += 4
            assert cid < len(chroms), "section refers to a chromosome id outside the tree"
            intervals.append((cid, s, e, v))
    summary = None
    if summ_off:
        cov, mn, mx, sm, sq = struct.unpack_from("<Qdddd", d, summ_off)
        summary = dict(bases_covered=cov, min_val=mn, max_val=mx, sum_data=sm, sum_squares=sq)
    zoom = []
    for z in range(zooms):
        reduction, _, zdata, zindex = struct.unpack_from("<IIQQ", d, 64 + 24 * z)
        (count,) = struct.unpack_from("<I", d, zdata)
        zmagic, _, zitems = struct.unpack_from("<IIQ", d, zindex)
        assert zmagic == 0x2468ACE0
        leaves.clear()
        if zitems:
            walk_r(zindex + 48)
        recs = []
        for doff, dsize in sorted(leaves):
            sec = d[doff:doff + dsize]
            if ubuf:
                sec = zlib.decompress(sec)
            for k in range(len(sec) // 32):
                recs.append(struct.unpack_from("<IIIIffff", sec, 32 * k))  # chrom, start, end, valid, min, max, sum, sumsq
        # UCSC writers put the record count in front of a level's blocks, libBigWig (deepTools) does not: reported, not checked
        zoom.append(dict(reduction=reduction, records=recs, leading_count=count))
    return dict(chroms=chroms, intervals=intervals, summary=summary, zoom_levels=zooms, version=version, zoom=zoom)
