"""Per-pass resident time over many back-to-back passes (looks for clock / thermal / allocator drift)."""
import os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bamnado_b200 import _native
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pairs = os.environ.get("PAIRS", "10000000")
bam = f"/dev/shm/bnd_sweep_{pairs}.bam"
if not os.path.exists(bam):
    subprocess.run([os.path.join(ROOT, "tools/_build/synth_bam"), "--out", bam, "--pairs", pairs], check=True, capture_output=True)
N = _native.lib()
with N.pileup(dict(bam_path=bam, bin_size=50, norm="rpkm"), dict(min_mapq=30, proper_pair=True)) as h:
    h.stage()
    ts = []
    for i in range(int(os.environ.get("PASSES", "24"))):
        t = time.perf_counter(); h.accumulate_resident(); ts.append(round((time.perf_counter() - t) * 1e3, 1))
    print("pairs", pairs, "resident ms per pass:", ts, flush=True)
    smi = subprocess.run("nvidia-smi --query-gpu=clocks.sm,clocks.mem,power.draw,temperature.gpu,clocks_event_reasons.active --format=csv,noheader", shell=True, capture_output=True, text=True).stdout
    print(smi.strip())
