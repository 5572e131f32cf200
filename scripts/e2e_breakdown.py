"""Host-clock breakdown of one bam-coverage job (create handle / to_bedgraph / destroy) on a synthetic BAM."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bamnado_b200 import _native
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
bam = f"/dev/shm/bnd_e2e_{pairs}.bam"
if not os.path.exists(bam):
    subprocess.run([os.path.join(ROOT, "tools/_build/synth_bam"), "--out", bam, "--pairs", str(pairs)], check=True, capture_output=True)
N = _native.lib()
for i in range(6):
    out = f"/dev/shm/bnd_e2e_out_{i}.bedgraph"
    t0 = time.perf_counter()
    h = N.pileup(dict(bam_path=bam, bin_size=50, norm="rpkm"), dict(min_mapq=30, proper_pair=True))
    t1 = time.perf_counter()
    h.to_bedgraph(out)
    t2 = time.perf_counter()
    h.close()
    t3 = time.perf_counter()
    print(f"job {i}: create {1e3*(t1-t0):.1f} ms, to_bedgraph {1e3*(t2-t1):.1f} ms, destroy {1e3*(t3-t2):.1f} ms, total {1e3*(t3-t0):.1f} ms", flush=True)
    os.unlink(out)
