"""Aggregates a summarize_profile.py listing (all lines) of the speculative inflate kernel by code region."""
import re, sys
listing, srcfile = sys.argv[1], "bamnado_b200/csrc/inflate_spec.cuh"
rows = []
for l in open(listing):
    m = re.match(r"\s+([\d.]+)\s+([\d.]+)\s+([\d.]+)\s+(\S+):(\d+)\s", l)
    if m:
        rows.append((float(m[1]), float(m[2]), float(m[3]), m[4], int(m[5])))
src = open(srcfile).read().split("\n")
def find(s):
    return next(i + 1 for i, l in enumerate(src) if s in l)
marks = [(find("__device__ bool build_lit_two_level"), "build tables"), (find("__device__ __noinline__ uint2 span_decode"), "span_decode"),
         (find("#ifdef BND_EMUL_STATS"), "block header/stored/staging"), (find("// ---- pass 1"), "round control (pass 1, fix-up, place, store call)"),
         (find("// ---- resolve the matches"), "resolve matches"), (find("if (status == INF_OK && outpos != isize)"), "kernel tail")]
def region(f, ln):
    if f == "inflate_spec.cuh":
        name = "prologue"
        for start, n in marks:
            if ln >= start:
                name = n
        return name
    if f == "inflate.cuh":
        if 100 <= ln <= 126: return "canon slow path / entries"
        if 186 <= ln <= 232: return "crc"
        if 233 <= ln <= 330: return "dynamic header"
        return "inflate.cuh other"
    return "intrinsics/common"
agg = {}
for i, s, t, f, ln in rows:
    a = agg.setdefault(region(f, ln), [0, 0])
    a[0] += i
    a[1] += s
for k, v in sorted(agg.items(), key=lambda x: -x[1][0]):
    print(f"{k:55s} inst {v[0]:5.1f}%  samples {v[1]:5.1f}%")
