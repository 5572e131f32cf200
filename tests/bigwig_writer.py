"""Minimal BigWig writer for the tests: builds inputs in the flavours other writers produce (raw or zlib sections; bedGraph,
variableStep and fixedStep section types; any items-per-section; chromosome ids in any order).  No zoom levels."""
from __future__ import annotations

import struct
import zlib


def write_bigwig(path, chroms: list, sections: list, compress=True):
    """chroms: [(name, id, length)]; sections: [dict(chrom_id, type=1|2|3, items=[...], start=, step=, span=)] in (id, start) order.
    type 1 items (start, end, value); type 2 items (start, value) with span; type 3 items value with start/step/span."""
    blobs, leaves, max_raw = [], [], 0
    for s in sections:
        t, items = s["type"], s["items"]
        if t == 1:
            body = b"".join(struct.pack("<IIf", a, b, v) for a, b, v in items)
            first, last = items[0][0], items[-1][1]
        elif t == 2:
            body = b"".join(struct.pack("<If", a, v) for a, v in items)
            first, last = items[0][0], items[-1][0] + s["span"]
        else:
            body = b"".join(struct.pack("<f", v) for v in items)
            first, last = s["start"], s["start"] + (len(items) - 1) * s["step"] + s["span"]
        raw = struct.pack("<IIIIIBBH", s["chrom_id"], first, last, s.get("step", 0), s.get("span", 0), t, 0, len(items)) + body
        max_raw = max(max_raw, len(raw))
        blobs.append(zlib.compress(raw) if compress else raw)
        leaves.append((s["chrom_id"], first, s["chrom_id"], last))
    key_size = max(len(c[0]) for c in chroms)
    tree = struct.pack("<IIIIQQ", 0x78CA8C91, max(len(chroms), 1), key_size, 8, len(chroms), 0)
    tree += struct.pack("<BBH", 1, 0, len(chroms))
    for name, cid, ln in sorted(chroms):
        tree += name.encode().ljust(key_size, b"\0") + struct.pack("<II", cid, ln)
    chrom_off = 64 + 40
    data_off = chrom_off + len(tree)
    pos = data_off + 8
    offs = []
    for b in blobs:
        offs.append(pos)
        pos += len(b)
    index_off = pos
    rt = struct.pack("<IIQIIIIQII", 0x2468ACE0, max(len(blobs), 1), len(blobs), leaves[0][0] if leaves else 0, leaves[0][1] if leaves else 0,
                     leaves[-1][2] if leaves else 0, leaves[-1][3] if leaves else 0, index_off, 1024, 0)
    rt += struct.pack("<BBH", 1, 0, len(blobs))
    for (sc, sb, ec, eb), o, b in zip(leaves, offs, blobs):
        rt += struct.pack("<IIIIQQ", sc, sb, ec, eb, o, len(b))
    header = struct.pack("<IHHQQQHHQQIQ", 0x888FFC26, 4, 0, chrom_off, data_off, index_off, 0, 0, 0, 64, max_raw if compress else 0, 0)
    summary = struct.pack("<Qdddd", 0, 0.0, 0.0, 0.0, 0.0)
    with open(path, "wb") as f:
        f.write(header + summary + tree + struct.pack("<Q", len(blobs)) + b"".join(blobs) + rt + struct.pack("<I", 0x888FFC26))
