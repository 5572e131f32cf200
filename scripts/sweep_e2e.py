"""Sweeps host-side knobs of one bam-coverage job (file -> bedGraph) on a synthetic BAM; each setting in a fresh process."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pairs = sys.argv[1] if len(sys.argv) > 1 else "10000000"
settings = [{}, {"BND_READERS": "12"}, {"BND_READERS": "16"}, {"BND_STAGES": "6"}, {"BND_STAGES": "8", "BND_READERS": "16"},
            {"BND_SEGMENT_BYTES": str(64 << 20)}, {"BND_SEGMENT_BYTES": str(64 << 20), "BND_STAGES": "6", "BND_READERS": "16"},
            {"BND_WRITERS": "12"}, {"BND_WRITERS": "16"}, {"BND_TEXT_CHUNK_BYTES": str(32 << 20)}, {"BND_TEXT_CHUNK_BYTES": str(128 << 20)}]
for s in settings:
    env = dict(os.environ, BND_DEBUG_TIMING="1", **s)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "scripts/e2e_breakdown.py"), pairs], env=env, capture_output=True, text=True)
    lines = [l for l in (r.stdout + r.stderr).splitlines() if l.startswith("job") or "to_bedgraph:" in l or "pass:" in l]
    jobs = [l for l in lines if l.startswith("job")][2:]
    tot = [float(l.split("total")[1].split("ms")[0]) for l in jobs]
    tb = [l for l in lines if "to_bedgraph:" in l][-1]
    ps = [l for l in lines if "pass:" in l][-1]
    print(s or "default", "total ms (jobs 2..5):", [round(t, 1) for t in tot], flush=True)
    print("   ", tb.replace("[bnd] ", ""), flush=True)
    print("   ", ps.replace("[bnd] ", ""), flush=True)
