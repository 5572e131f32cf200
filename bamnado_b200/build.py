"""Build recipe for libbamnado_b200.so (nvcc, sm_100a) — the only artefact the package loads.

    python -m bamnado_b200.build            # product library (needs nvcc; cross-compiles without a GPU)
    python -m bamnado_b200.build --emul     # tests/emul: the same sources against the CPU SIMT emulator (tests only)
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "bamnado_b200" / "csrc"
OUT_DIR = ROOT / "bamnado_b200" / "_lib"
LIB = OUT_DIR / "libbamnado_b200.so"
EMUL_DIR = ROOT / "tests" / "emul"
EMUL_LIB = EMUL_DIR / "_build" / "libbamnado_b200_emul.so"

HOST_SOURCES = ["bamio.cpp", "filter_host.cpp", "engine.cpp", "outputs.cpp", "capi.cpp"]
DEVICE_SOURCES = ["kernels.cu"]
NVCC_ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]


def _newest(paths):
    return max(p.stat().st_mtime for p in paths)


def _sources():
    return [p for p in CSRC.iterdir() if p.suffix in (".cu", ".cuh", ".cpp", ".h")] + [ROOT / "include" / "bamnado_b200.h"]


def nvcc_path() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: bamnado_b200 is CUDA-only (sm_100a) and has no CPU fallback")


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every CUDA/C++ source into bamnado_b200/_lib/libbamnado_b200.so for sm_100a."""
    if not force and LIB.exists() and LIB.stat().st_mtime >= _newest(_sources()):
        return LIB
    OUT_DIR.mkdir(parents=True, exist_ok=True)
    obj_dir = OUT_DIR / "obj"
    obj_dir.mkdir(exist_ok=True)
    nvcc = nvcc_path()
    objs = []
    jobs = []
    for src in DEVICE_SOURCES:
        o = obj_dir / (src + ".o")
        cmd = [nvcc, *NVCC_ARCH, "-lineinfo", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-Xptxas", "-v", "-c", str(CSRC / src), "-o", str(o)]
        jobs.append((cmd, o))
    for src in HOST_SOURCES:
        o = obj_dir / (src + ".o")
        # host files hold no device code; plain g++ with the CUDA include path is enough and much faster
        cmd = ["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-Wall", "-Wextra", "-pthread", "-I", "/usr/local/cuda/include", "-c", str(CSRC / src), "-o", str(o)]
        jobs.append((cmd, o))
    procs = [(subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True), cmd, o) for cmd, o in jobs]
    log = []
    for p, cmd, o in procs:
        out, _ = p.communicate()
        log.append(out)
        if p.returncode != 0:
            raise RuntimeError("build failed: " + " ".join(cmd) + "\n" + out)
        objs.append(str(o))
    (OUT_DIR / "build.log").write_text("\n".join(log))
    link = [nvcc, *NVCC_ARCH, "-shared", "-o", str(LIB), *objs, "-Xcompiler", "-pthread", "-cudart", "static"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed: " + " ".join(link) + "\n" + r.stdout + r.stderr)
    if verbose:
        print("\n".join(log))
    return LIB


def build_emul(force: bool = False) -> Path:
    """tests only: the same sources compiled by g++ against tests/emul/cuda_emul.h (CPU SIMT emulator)."""
    srcs = _sources() + [EMUL_DIR / "cuda_emul.h", EMUL_DIR / "cuda_emul.cpp"]
    if not force and EMUL_LIB.exists() and EMUL_LIB.stat().st_mtime >= _newest(srcs):
        return EMUL_LIB
    EMUL_LIB.parent.mkdir(parents=True, exist_ok=True)
    asan = ["-fsanitize=address", "-fno-omit-frame-pointer"] if os.environ.get("BND_EMUL_ASAN") else []  # debugging aid: LD_PRELOAD libasan.so
    common = ["g++", "-O1", "-g", *asan, "-std=c++17", "-fPIC", "-pthread", "-DBND_EMUL", "-I", str(EMUL_DIR), "-I", str(CSRC), "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-function"]
    objs, procs = [], []
    for src in DEVICE_SOURCES + HOST_SOURCES:
        o = EMUL_LIB.parent / (src + ".o")
        procs.append((subprocess.Popen([*common, "-x", "c++", "-c", str(CSRC / src), "-o", str(o)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True), src))
        objs.append(str(o))
    o = EMUL_LIB.parent / "cuda_emul.o"
    procs.append((subprocess.Popen([*common, "-c", str(EMUL_DIR / "cuda_emul.cpp"), "-o", str(o)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True), "cuda_emul.cpp"))
    objs.append(str(o))
    for p, src in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise RuntimeError(f"emulator build failed for {src}:\n{out}")
    r = subprocess.run(["g++", "-shared", "-pthread", *asan, "-o", str(EMUL_LIB), *objs], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("emulator link failed:\n" + r.stdout + r.stderr)
    return EMUL_LIB


if __name__ == "__main__":
    if "--emul" in sys.argv:
        print(build_emul(force="--force" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
