"""to_bigwig several times in one process (fresh handle each): catches state carried over through the device buffer pool.
python scripts/bigwig_twice.py [pairs] [reps]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
from bamnado_b200 import _native

pairs, reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2000000, int(sys.argv[2]) if len(sys.argv) > 2 else 3
class A: pass
a = A(); a.config, a.pairs, a.gpus = 4, pairs, 1
W = bench.Workload(a).ensure()
N = _native.lib()
pile = dict(W.pile, bam_path=str(W.bam), device=0)
for k in range(reps):
    out = bench.outdir() / f"twice_{k}.bw"
    t0 = time.perf_counter()
    with N.pileup(pile, W.filt) as h:
        h.to_bigwig(out)
    print(f"rep {k}: {1e3 * (time.perf_counter() - t0):.1f} ms, {out.stat().st_size} bytes", flush=True)
    names, cid, st, en, val = bench._bigwig_intervals(out)
    print(f"  {len(st)} items re-read, {len(names)} chromosomes", flush=True)
    out.unlink()
