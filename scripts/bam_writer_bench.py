"""Timing of the BAM writers (split / modify / split-exogenous) on a synthetic paired-end BAM; the output is re-read with the
oracle's zlib reader (record count, every block's CRC) - test infrastructure used as the checker only.

usage: python scripts/bam_writer_bench.py [pairs]"""
import json, os, subprocess, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g
g.build()
from bamnado_b200 import _native
from oracle import oracle_py as O
import ctypes as C

pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
d = "/dev/shm/bnd_writer_bench"
os.makedirs(d, exist_ok=True)
bam = f"{d}/in_{pairs}.bam"
if not os.path.exists(bam):
    subprocess.run([os.path.join(ROOT, "tools/_build/synth_bam"), "--out", bam, "--pairs", str(pairs), "--seed", "20260219", "--unplaced", str(pairs // 100)], check=True, capture_output=True)
in_bytes = os.path.getsize(bam)
N = _native.lib()
os.environ["BND_DEBUG_TIMING"] = "1"


def count(path):
    n, infl, blocks = C.c_uint64(), C.c_uint64(), C.c_uint64()
    O._check(O.lib().orc_count_records(O._b(path), C.byref(n), C.byref(infl), C.byref(blocks)))
    return n.value, infl.value, blocks.value


res = {}
for name, fn in [("split", lambda o: N.bam_split(bam, o, dict(min_mapq=30, proper_pair=True))),
                 ("split_all", lambda o: N.bam_split(bam, o, dict())),
                 ("modify_tn5", lambda o: N.bam_modify(bam, o, dict(min_mapq=20), tn5_shift=True))]:
    out = f"{d}/{name}.bam"
    best = 1e9
    for rep in range(3):
        t = time.perf_counter()
        st = fn(out)
        best = min(best, time.perf_counter() - t)
    n, infl, blocks = count(out)
    assert n == st["n_written"], (name, n, st)
    res[name] = dict(seconds=best, records_in=2 * pairs + pairs // 100, records_out=n, out_bytes=os.path.getsize(out), out_inflated=infl, out_blocks=blocks,
                     reads_per_s=(2 * pairs) / best, in_GBps=in_bytes / best / 1e9, ratio=infl / max(os.path.getsize(out), 1))
    os.unlink(out)
t = time.perf_counter()
st = N.bam_split_exogenous(bam, f"{d}/sp", "chrUn", dict(min_mapq=20))
dt = time.perf_counter() - t
tot = 0
for k in ("endogenous", "exogenous", "both", "unmapped"):
    tot += count(f"{d}/sp.{k}.bam")[0]
    os.unlink(f"{d}/sp.{k}.bam")
assert tot == st["n_total_reads"] - st["n_filtered_reads"], (tot, st)
res["split_exogenous"] = dict(seconds=dt, stats=st, reads_per_s=st["n_total_reads"] / dt)
print(json.dumps(dict(pairs=pairs, in_bytes=in_bytes, results=res)))
