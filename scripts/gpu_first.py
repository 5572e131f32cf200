"""First GPU contact: parity on test.bam + a synthetic BAM, then stage timings on a larger synthetic BAM."""
import json, os, subprocess, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from bamnado_b200 import _native
from oracle import oracle_py as o

N = _native.lib()
print("devices", N.device_count(), flush=True)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SYNTH = os.path.join(ROOT, "tools/_build/synth_bam")

def check(bam, pile, filt, label):
    pile = dict(pile, bam_path=bam)
    t = time.time()
    h = N.pileup(pile, filt)
    runs = h.runs()
    t_gpu = time.time() - t
    t = time.time()
    oruns, ostats, of = o.pileup_runs(pile, filt)
    t_cpu = time.time() - t
    ok = len(runs) == len(oruns) and bool((runs == oruns).all()) and h.stats() == ostats and h.norm_factor() == of
    txt = h.bedgraph_text()
    o.pileup_to_bedgraph(pile, filt, "/tmp/o.bdg")
    ok_txt = txt == open("/tmp/o.bdg", "rb").read()
    print(json.dumps(dict(case=label, ok=ok, ok_text=ok_txt, rows=len(runs), gpu_s=round(t_gpu, 3), cpu_s=round(t_cpu, 3),
                          timings=h.timings(), shape=h.shape())), flush=True)
    h.close()
    return ok and ok_txt

allok = True
tb = os.path.join(ROOT, "tests/golden/ref_data/test.bam")
allok &= check(tb, dict(bin_size=100), dict(min_mapq=20, min_length=20, max_length=1000), "cfg1")
allok &= check(tb, dict(bin_size=1000, norm="cpm"), dict(proper_pair=True, max_length=1000), "pin_cpm")
allok &= check(tb, dict(bin_size=1, norm="rpkm"), dict(min_mapq=30), "bin1_collapse")
allok &= check(tb, dict(bin_size=1000, collapse=False, norm="rpkm"), dict(proper_pair=True, max_length=1000), "pin_rpkm_nocollapse")
os.makedirs("/dev/shm/bnd", exist_ok=True)
small = "/dev/shm/bnd/s1m.bam"
subprocess.run([SYNTH, "--out", small, "--pairs", "1000000", "--odd"], check=True, capture_output=True)
allok &= check(small, dict(bin_size=50, norm="rpkm"), dict(min_mapq=30, proper_pair=True), "synth1m_cfg2")
allok &= check(small, dict(bin_size=10, use_fragment=True), dict(max_fragment_length=120, min_mapq=20), "synth1m_frag")
print("PARITY", "OK" if allok else "FAILED", flush=True)

big = "/dev/shm/bnd/s10m.bam"
t = time.time()
r = subprocess.run([SYNTH, "--out", big, "--pairs", "10000000"], check=True, capture_output=True, text=True)
print("gen", round(time.time() - t, 1), r.stdout.strip(), flush=True)
pile = dict(bam_path=big, bin_size=50, norm="rpkm")
filt = dict(min_mapq=30, proper_pair=True)
for it in range(4):
    t = time.time()
    h = N.pileup(pile, filt)
    t1 = time.time()
    h.accumulate()
    t2 = time.time()
    n = h.finalize()
    t3 = time.time()
    txt = h.bedgraph_text()
    t4 = time.time()
    print(json.dumps(dict(iter=it, create_s=round(t1 - t, 4), accumulate_s=round(t2 - t1, 4), finalize_s=round(t3 - t2, 4),
                          text_s=round(t4 - t3, 4), rows=n, text_bytes=len(txt), timings=h.timings(), shape=h.shape())), flush=True)
    if it == 3:
        h.stage()
        for k in range(3):
            t = time.time(); h.accumulate_resident(); dt = time.time() - t
            print(json.dumps(dict(resident_iter=k, s=round(dt, 4), timings=h.timings())), flush=True)
    h.close()
t = time.time()
st, f = o.pileup_to_bedgraph(pile, filt, "/dev/shm/bnd/o.bdg")
print("oracle all-core s", round(time.time() - t, 2), st["n_total"], flush=True)
h = N.pileup(pile, filt); h.to_bedgraph("/dev/shm/bnd/g.bdg"); h.close()
same = subprocess.run(["cmp", "/dev/shm/bnd/o.bdg", "/dev/shm/bnd/g.bdg"]).returncode == 0
print("big bedgraph identical:", same, flush=True)
