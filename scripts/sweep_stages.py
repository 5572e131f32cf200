"""Serialised per-stage CUDA-event times of a resident pass (BND_SERIALIZE=1) for PAIRS."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BND_SERIALIZE"] = "1"
from bamnado_b200 import _native
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pairs = os.environ.get("PAIRS", "10000000")
bam = f"/dev/shm/bnd_sweep_{pairs}.bam"
if not os.path.exists(bam):
    subprocess.run([os.path.join(ROOT, "tools/_build/synth_bam"), "--out", bam, "--pairs", pairs], check=True, capture_output=True)
N = _native.lib()
with N.pileup(dict(bam_path=bam, bin_size=50, norm="rpkm"), dict(min_mapq=30, proper_pair=True)) as h:
    h.stage()
    for i in range(3):
        t = time.perf_counter(); h.accumulate_resident(); dt = time.perf_counter() - t
    tm, sh = h.timings(), h.shape()
    print("pairs", pairs, "wall_ms", round(dt * 1e3, 1), {k: round(v, 2) for k, v in tm.items() if k.endswith("_ms") and v}, "segments", sh["launches"] // 5,
          "blocks", sh["blocks"], "records", sh["records"], flush=True)
