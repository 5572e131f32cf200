"""First-segment size (pipeline fill) at full size, fresh process per setting."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pairs = sys.argv[1] if len(sys.argv) > 1 else "100000000"
for s in [{}, {"BND_FIRST_SEGMENT_BYTES": str(24 << 20)}, {"BND_FIRST_SEGMENT_BYTES": str(48 << 20)}, {}]:
    env = dict(os.environ, BND_DEBUG_TIMING="1", **s)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts/e2e_breakdown.py"), pairs], env=env, capture_output=True, text=True)
    lines = (r.stdout + r.stderr).splitlines()
    tot = [float(l.split("total")[1].split("ms")[0]) for l in lines if l.startswith("job")][2:]
    pu = [float(l.split("pile-up")[1].split("ms")[0]) for l in lines if "to_bedgraph:" in l][2:]
    print(s or "default", "total ms:", [round(t, 1) for t in tot], "pile-up ms:", [round(t, 1) for t in pu], flush=True)
