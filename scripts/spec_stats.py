"""Walk statistics of the speculative inflate kernel on the CPU emulator (tests/emul), to pick SPAN / NT.

    BND_EMUL_STATS=1 python -m bamnado_b200.build --emul --force
    BND_INFLATE_VARIANT=6 python scripts/spec_stats.py [pairs]
Counters (inflate_spec.cuh, BND_STAT): 0 rounds, 1 fix-up iterations, 2 lanes re-decoded in fix-ups, 3 valid lanes
dropped because the token array was full, 4 tokens.
"""
import ctypes, os, subprocess, sys, zlib
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from bamnado_b200._native import Native  # noqa: E402
from bamnado_b200.build import EMUL_LIB  # noqa: E402

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
work = Path("/tmp/spec_stats")
work.mkdir(exist_ok=True)
bam = work / f"s{pairs}.bam"
if not bam.exists():
    synth = ROOT / "tools" / "_build" / "synth_bam"
    if not synth.exists():
        synth.parent.mkdir(exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-o", str(synth), str(ROOT / "tools" / "synth_bam.cpp"), "-lz", "-pthread"])
    subprocess.check_call([str(synth), "--out", str(bam), "--pairs", str(pairs), "--seed", "20260219"])
nat = Native(ctypes.CDLL(str(EMUL_LIB)))
data = bam.read_bytes()
out, st = nat.bgzf_inflate(data)
assert (st == 0).all(), st
d = zlib.decompressobj(31)
ref = b""
buf = data
while buf:
    d = zlib.decompressobj(31)
    ref += d.decompress(buf)
    buf = d.unused_data
assert out == ref, "inflate mismatch"
f = nat.L.bnd_dbg_counter
f.restype = ctypes.c_ulonglong
c = [f(i) for i in range(19)]
print(f"blocks {len(st)}  inflated {len(out)} B  compressed {len(data)} B")
print(f"rounds {c[0]}  matches/round {c[4] / max(c[0], 1):.1f}")
print(f"fix-up iterations/round {c[1] / max(c[0], 1):.2f}  lanes re-decoded/round {c[2] / max(c[0], 1):.2f}  dropped lanes/round {c[3] / max(c[0], 1):.2f}")
print(f"matches {c[4]}  batches of 32: {c[6]}  waves/batch {c[5] / max(c[6], 1):.2f}  self-overlapping firsts/batch {c[7] / max(c[6], 1):.2f}")
print(f"ready matches/wave {c[9] / max(c[5] - c[7], 1):.1f}  of them longer than 16 B: {c[10] / max(c[9], 1):.3f}  left to the one-by-one loop/batch {c[8] / max(c[6], 1):.2f}")
if c[12]:
    print(f"first fix-up pass: a lane is back on its speculative path after {c[11] / max(c[12], 1) * 100:.1f} % of its span on average "
          f"({c[11] / max(c[0], 1) / 31:.0f} bits); the slowest lane of a round needs {c[13] / max(c[0], 1):.0f} bits; "
          f"{c[14] / max(c[0], 1):.2f} lanes per round never meet it inside the span")
print(f"waves per round: {c[5] / max(c[0], 1):.1f} in batches of 32, {c[15] / max(c[0], 1):.1f} if a whole round were resolved at once")
print(f"exact dependency depth of a round's matches (a match waits only for matches whose output its source overlaps): {c[16] / max(c[0], 1):.1f}")
print(f"pass 1: {c[17] / max(c[0], 1) / 32:.1f} tokens per lane on average, {c[18] / max(c[0], 1):.1f} on the slowest lane of a round "
      f"(a pass lasts as long as its slowest lane: {c[18] / max(c[17] / 32, 1):.2f}x the mean)")
