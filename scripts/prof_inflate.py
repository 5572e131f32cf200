"""Profiling target: K1 (BGZF inflate) alone over a 1M-pair synthetic BAM (one launch per call)."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bamnado_b200 import _native
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
bam = "/dev/shm/bnd_prof_1m.bam"
if not os.path.exists(bam):
    subprocess.run([os.path.join(ROOT, "tools/_build/synth_bam"), "--out", bam, "--pairs", "1000000"], check=True, capture_output=True)
data = open(bam, "rb").read()
N = _native.lib()
for i in range(3):
    t = time.time()
    out, st = N.bgzf_inflate(data)
    print(i, len(out), int((st != 0).sum()), round(time.time() - t, 4), flush=True)
