"""Wall time of the resident and streaming passes on a 10M-pair BAM for the env knobs given (segment size, variant)."""
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
    res, strm = [], []
    for i in range(4):
        t = time.perf_counter(); h.accumulate_resident(); res.append(time.perf_counter() - t)
    for i in range(4):
        t = time.perf_counter(); h.accumulate(); strm.append(time.perf_counter() - t)
    s = h.shape()
    print("variant", os.environ.get("BND_INFLATE_VARIANT", "0"), "seg", os.environ.get("BND_SEGMENT_BYTES", "dflt"),
          "resident_ms", round(min(res) * 1e3, 2), "stream_ms", round(min(strm) * 1e3, 2),
          "inflated_GBps_resident", round(s["inflated_bytes"] / min(res) / 1e9, 1), "Mreads/s", round(s["records"] / min(res) / 1e6, 1), flush=True)
