"""K1 (speculative 32-lane inflate, bamnado_b200/csrc/inflate_spec.cuh) against zlib on streams chosen to hit its corner cases.

The walk cuts the stream into 32 spans per round and trusts only spans that start on a verified code boundary, so the
cases here vary what a span can contain: one-bit codes (hundreds of tokens per span, token array full), 15-bit codes
(second-level table), fixed and stored blocks, several deflate blocks inside one BGZF block (end-of-block inside a
round), streams shorter than one span, and corrupted streams (every error must surface as a status, never as wrong
bytes).  Replaces noodles-bgzf block decoding (bamnado/src/coverage_analysis.rs:456-466).
"""
import struct
import zlib

import numpy as np
import pytest


def bgzf_block(payload: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, mem_level=9, flush_every=0) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15, mem_level, strategy)
    if flush_every:  # Z_FULL_FLUSH ends the deflate block: several blocks (and empty stored ones) per BGZF block
        comp = b""
        for i in range(0, len(payload), flush_every):
            comp += co.compress(payload[i:i + flush_every]) + co.flush(zlib.Z_FULL_FLUSH)
        comp += co.flush()
    else:
        comp = co.compress(payload) + co.flush()
    assert len(comp) + 26 <= 65536, "payload does not fit one BGZF block"
    return (b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", 18 + len(comp) + 8 - 1) + comp
            + struct.pack("<II", zlib.crc32(payload), len(payload)))


def payload_zoo(seed=11):
    rng = np.random.default_rng(seed)
    zoo = {}
    zoo["two_symbols"] = bytes(rng.integers(0, 2, 65000, dtype=np.uint8) + 65)          # ~1-bit literal codes
    zoo["skewed"] = bytes(rng.choice(np.arange(64, dtype=np.uint8), 60000, p=np.r_[0.6, np.full(63, 0.4 / 63)]))
    zoo["zipf"] = bytes(np.minimum(rng.zipf(1.3, 60000), 255).astype(np.uint8))          # long tail -> 13-15 bit codes
    zoo["all_bytes_once_then_run"] = bytes(range(256)) + b"\x07" * 60000
    zoo["random"] = bytes(rng.integers(0, 256, 50000, dtype=np.uint8))
    zoo["long_matches"] = (bytes(rng.integers(0, 256, 300, dtype=np.uint8)) * 220)[:65000]
    zoo["period_1_to_40"] = b"".join(bytes(rng.integers(0, 256, p, dtype=np.uint8)) * (1500 // p) for p in range(1, 41))
    rec = bytes(rng.integers(0, 256, 220, dtype=np.uint8))
    zoo["bam_like"] = b"".join(bytes(a ^ (rng.integers(0, 256) if rng.random() < 0.08 else 0) for a in rec) for _ in range(280))
    zoo["tiny_1"] = b"x"
    zoo["tiny_31"] = bytes(range(31))
    zoo["empty"] = b""
    return zoo


def test_spec_inflate_matches_zlib_over_payload_zoo(backend):
    zoo = payload_zoo()
    blocks, expected, names = [], [], []
    for name, p in zoo.items():
        for level, strategy in [(6, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED),
                                (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE), (6, zlib.Z_FILTERED)]:
            if strategy == zlib.Z_HUFFMAN_ONLY and name == "random":
                p_use = p[:40000]  # incompressible + Huffman-only would overflow one BGZF block
            else:
                p_use = p
            try:
                blocks.append(bgzf_block(p_use, level, strategy))
            except AssertionError:
                continue
            expected.append(p_use)
            names.append((name, level, strategy))
    out, status = backend.bgzf_inflate(b"".join(blocks))
    bad = [n for n, s in zip(names, status) if s != 0]
    assert not bad, f"non-zero status for {bad}"
    assert out == b"".join(expected)


def test_spec_inflate_many_deflate_blocks_per_bgzf_block(backend):
    # end-of-block codes, block headers and empty stored blocks in the middle of rounds; small mem_level = short blocks
    zoo = payload_zoo(5)
    blocks, expected = [], []
    for name in ("bam_like", "skewed", "zipf", "two_symbols", "long_matches"):
        for flush_every in (97, 1000, 4096):
            p = zoo[name][:30000]
            blocks.append(bgzf_block(p, 6, flush_every=flush_every))
            expected.append(p)
        blocks.append(bgzf_block(zoo[name][:60000], 6, mem_level=1))  # 128-token deflate blocks
        expected.append(zoo[name][:60000])
    out, status = backend.bgzf_inflate(b"".join(blocks))
    assert list(status) == [0] * len(blocks)
    assert out == b"".join(expected)


def test_spec_inflate_corruption_is_detected_not_decoded(backend):
    # flip one bit at many places of a valid block: the status must be non-zero (CRC at the very least) or the bytes equal
    rng = np.random.default_rng(17)
    zoo = payload_zoo(23)
    for name in ("bam_like", "zipf"):
        payload = zoo[name][:40000]
        good = bytearray(bgzf_block(payload))
        variants = []
        for _ in range(24):
            b = bytearray(good)
            pos = int(rng.integers(18, len(b) - 8))
            b[pos] ^= 1 << int(rng.integers(0, 8))
            variants.append(bytes(b))
        # each corrupted block is followed by a good one: an error must not leak into the neighbour
        stream = b"".join(v + bytes(good) for v in variants)
        out, status = backend.bgzf_inflate(stream)
        assert len(status) == 2 * len(variants)
        off = 0
        for i, s in enumerate(status):
            chunk = out[off:off + len(payload)]
            off += len(payload)
            if i % 2 == 1:
                assert s == 0 and chunk == payload
            else:
                assert s != 0 or chunk == payload


def test_spec_inflate_rejects_oversized_block(backend):
    # a gzip member that claims (and holds) more than the 64 KiB a BGZF block may inflate to (SAM spec 4.1) is refused by
    # the block scan before anything is launched (the kernel's 16-bit match positions rely on it)
    with pytest.raises(RuntimeError, match="ISIZE exceeds 64 KiB"):
        backend.bgzf_inflate(bgzf_block(b"hello world " * 100) + bgzf_block(b"A" * 70000))


@pytest.mark.gpu
def test_spec_inflate_zoo_on_gpu_many_copies(native_gpu):
    # the same zoo, every block repeated so that all resident warps decode different streams side by side
    zoo = payload_zoo(31)
    blocks, expected = [], []
    for rep in range(40):
        for name, p in zoo.items():
            q = p[rep:] if len(p) > rep else p
            blocks.append(bgzf_block(q, 1 + rep % 9))
            expected.append(q)
    out, status = native_gpu.bgzf_inflate(b"".join(blocks))
    assert int(np.count_nonzero(status)) == 0
    assert out == b"".join(expected)


@pytest.mark.parametrize("variant", ["1", "2", "3", "4"])
def test_other_kernel_shapes_stay_correct(variant, tmp_path):
    # BND_INFLATE_VARIANT is read once per process: 256-bit spans (1), the 256-entry match list (2: lanes are dropped
    # whenever a round holds more matches) and one-warp CTAs reading the CRC tables from global memory (3) run the zoo in a child
    # process against the CPU emulator build, as does the previous default (4: 320-bit spans, 384 matches per round); the shipped shape (0)
    # runs everywhere else
    import subprocess
    import sys
    from pathlib import Path

    root = Path(__file__).resolve().parent.parent
    code = (
        "import ctypes, sys, zlib\n"
        f"sys.path.insert(0, {str(root)!r}); sys.path.insert(0, {str(root / 'tests')!r})\n"
        "from bamnado_b200._native import Native\n"
        "from bamnado_b200.build import build_emul\n"
        "from test_inflate_stress import payload_zoo, bgzf_block\n"
        "nat = Native(ctypes.CDLL(str(build_emul())))\n"
        "zoo = payload_zoo(3)\n"
        "blocks, exp = [], []\n"
        "for name, p in zoo.items():\n"
        "    for lvl, strat in ((6, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_HUFFMAN_ONLY), (1, zlib.Z_FIXED)):\n"
        "        q = p[:40000] if strat == zlib.Z_HUFFMAN_ONLY and name == 'random' else p\n"
        "        try:\n"
        "            blocks.append(bgzf_block(q, lvl, strat)); exp.append(q)\n"
        "        except AssertionError:\n"
        "            pass\n"
        "out, st = nat.bgzf_inflate(b''.join(blocks))\n"
        "assert not st.any(), st\n"
        "assert out == b''.join(exp)\n"
        "print('ok', len(blocks))\n"
    )
    env = dict(**__import__("os").environ, BND_INFLATE_VARIANT=variant)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and r.stdout.startswith("ok"), r.stdout + r.stderr


def _random_payload(rng, max_len=60000):
    """A payload stitched from pieces of very different statistics, so that code lengths, match lengths and distances
    change inside one block (and inside one round of the speculative walk)."""
    out = bytearray()
    n = int(rng.integers(1, max_len))
    while len(out) < n:
        kind = int(rng.integers(0, 6))
        k = int(rng.integers(1, 4000))
        if kind == 0:
            out += bytes(rng.integers(0, 256, k, dtype=np.uint8))                        # incompressible
        elif kind == 1:
            out += bytes([int(rng.integers(0, 256))]) * k                                  # one long run
        elif kind == 2:
            alpha = rng.integers(0, 256, int(rng.integers(2, 20)), dtype=np.uint8)
            out += bytes(alpha[rng.integers(0, len(alpha), k)])                            # small alphabet
        elif kind == 3 and len(out) > 8:
            d = int(rng.integers(1, min(len(out), 32768) + 1))                             # copy from far back
            src = len(out) - d
            out += bytes(out[src:src + min(k, d)])
        elif kind == 4:
            unit = bytes(rng.integers(0, 256, int(rng.integers(1, 300)), dtype=np.uint8))
            out += (unit * (k // len(unit) + 1))[:k]                                       # periodic
        else:
            out += bytes(np.minimum(rng.geometric(0.08, k), 255).astype(np.uint8))         # skewed
    return bytes(out[:n])


def test_spec_inflate_random_structures(backend):
    rng = np.random.default_rng(20260219)
    blocks, expected = [], []
    strategies = [zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_RLE, zlib.Z_HUFFMAN_ONLY, zlib.Z_FIXED]
    while len(blocks) < 120:
        p = _random_payload(rng)
        level = int(rng.integers(1, 10))
        strat = strategies[int(rng.integers(0, len(strategies)))]
        flush = int(rng.choice([0, 0, 0, 513, 5000]))
        try:
            blocks.append(bgzf_block(p, level, strat, mem_level=int(rng.integers(1, 10)), flush_every=flush))
        except AssertionError:
            continue  # did not fit one BGZF block
        expected.append(p)
    out, status = backend.bgzf_inflate(b"".join(blocks))
    assert not np.count_nonzero(status), [int(i) for i in np.nonzero(status)[0]]
    assert out == b"".join(expected)
