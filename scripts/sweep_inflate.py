"""Times K1 (serialised launches, CUDA events) for one BND_INFLATE_VARIANT on a 2M-pair synthetic BAM."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bamnado_b200 import _native
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
bam = "/dev/shm/bnd_prof_2m.bam"
if not os.path.exists(bam):
    subprocess.run([os.path.join(ROOT, "tools/_build/synth_bam"), "--out", bam, "--pairs", "2000000"], check=True, capture_output=True)
os.environ["BND_SERIALIZE"] = "1"
N = _native.lib()
with N.pileup(dict(bam_path=bam, bin_size=50), dict(min_mapq=30, proper_pair=True)) as h:
    best = None
    for i in range(4):
        h.accumulate()
        t, s = h.timings(), h.shape()
        best = t["inflate_ms"] if best is None else min(best, t["inflate_ms"])
    print("variant", os.environ.get("BND_INFLATE_VARIANT", "0"), "inflate_ms", round(best, 3), "inflated_GBps", round(s["inflated_bytes"] / best / 1e6, 2),
          "index_ms", round(t["index_ms"], 3), "pileup_ms", round(t["pileup_ms"], 3), "records", s["records"], flush=True)
