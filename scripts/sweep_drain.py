"""Sweeps the writer-side knobs of to_bedgraph (fresh process per setting)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pairs = sys.argv[1] if len(sys.argv) > 1 else "10000000"
settings = [{"BND_WRITERS": str(w)} for w in (8, 16, 24, 32, 48)] * 2
for s in settings:
    env = dict(os.environ, BND_DEBUG_TIMING="1", **s)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts/e2e_breakdown.py"), pairs], env=env, capture_output=True, text=True)
    lines = (r.stdout + r.stderr).splitlines()
    tot = [float(l.split("total")[1].split("ms")[0]) for l in lines if l.startswith("job")][2:]
    dr = [float(l.split("drain")[1].split("ms")[0]) for l in lines if "to_bedgraph:" in l][2:]
    print(s or "default", "total ms:", [round(t, 1) for t in tot], "drain ms:", [round(t, 1) for t in dr], flush=True)
