"""`split` at the BASELINE size: the 100 M-pair BAM of bench.py (run bench.py first: the BAM is cached in its work directory),
a filter that keeps about a third of the reads, several stream flushes per pass.  The output is re-read with the oracle's zlib
reader (checker only): record count, CRC-32 of every block.

usage: python scripts/bam_writer_full.py"""
import ctypes as C, glob, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bamnado_b200 import _native
from oracle import oracle_py as O
import bench

bams = sorted(glob.glob(str(bench.workdir() / "*.bam")), key=os.path.getsize)
if not bams:
    raise SystemExit("no cached bench BAM: run bench.py first")
bam = bams[-1]
out = str(bench.workdir() / "split_full.bam")
N = _native.lib()
os.environ["BND_DEBUG_TIMING"] = "1"
filt = dict(min_mapq=60, proper_pair=True, strand="forward", min_length=20, max_length=1000)
best, st = 1e9, None
for rep in range(2):
    t = time.perf_counter()
    st = N.bam_split(bam, out, filt, command_line="bamnado split (bench)")
    best = min(best, time.perf_counter() - t)
n, infl, blocks = C.c_uint64(), C.c_uint64(), C.c_uint64()
O._check(O.lib().orc_count_records(O._b(out), C.byref(n), C.byref(infl), C.byref(blocks)))
assert n.value == st["n_written"], (n.value, st)
size = os.path.getsize(out)
os.unlink(out)
print(json.dumps(dict(bam=os.path.basename(bam), in_bytes=os.path.getsize(bam), records_in=st["n_total_reads"], records_out=st["n_written"], out_bytes=size,
                      out_inflated=infl.value, out_blocks=blocks.value, seconds=best, reads_per_s=st["n_total_reads"] / best,
                      in_GBps=os.path.getsize(bam) / best / 1e9, checked="record count + CRC-32 of every block (oracle zlib reader)")))
