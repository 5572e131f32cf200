"""Sweeps the reader-side knobs of the streaming pile-up at full size (fresh process per setting)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pairs = sys.argv[1] if len(sys.argv) > 1 else "100000000"
settings = [{}, {"BND_READERS": "12"}, {"BND_READERS": "16"}, {"BND_READERS": "12", "BND_STAGES": "6"}, {"BND_SETS": "8"}, {}]
for s in settings:
    env = dict(os.environ, BND_DEBUG_TIMING="1", **s)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts/e2e_breakdown.py"), pairs], env=env, capture_output=True, text=True)
    lines = (r.stdout + r.stderr).splitlines()
    tot = [float(l.split("total")[1].split("ms")[0]) for l in lines if l.startswith("job")][2:]
    pu = [float(l.split("pile-up")[1].split("ms")[0]) for l in lines if "to_bedgraph:" in l][2:]
    ps = [l for l in lines if "pass:" in l][-1]
    print(s or "default", "total ms:", [round(t, 1) for t in tot], "pile-up ms:", [round(t, 1) for t in pu], flush=True)
    print("   ", ps.replace("[bnd] ", ""), flush=True)
