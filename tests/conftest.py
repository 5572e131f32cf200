"""Shared fixtures.  `-m "not gpu"` runs the oracle / host / emulator tests on CPU; `-m gpu` runs the parity tests
proper through the C ABI of the sm_100a library."""
from __future__ import annotations

import ctypes
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

GOLDEN = ROOT / "tests" / "golden"
TEST_BAM = GOLDEN / "ref_data" / "test.bam"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
    config.addinivalue_line("markers", "emul: runs the kernel sources on the test-only CPU SIMT emulator")


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle_py

    oracle_py.build()
    return oracle_py


@pytest.fixture(scope="session")
def synth_bin():
    out = ROOT / "tools" / "_build" / "synth_bam"
    src = ROOT / "tools" / "synth_bam.cpp"
    if not out.exists() or out.stat().st_mtime < src.stat().st_mtime:
        out.parent.mkdir(parents=True, exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", str(out), str(src), "-lz"], check=True)
    return out


@pytest.fixture(scope="session")
def make_bam(synth_bin, tmp_path_factory):
    """make_bam(name, *synth_args) -> path of a seeded synthetic BAM (+ .bai), cached for the session."""
    base = tmp_path_factory.mktemp("bams")
    cache = {}

    def make(name, *args):
        if name not in cache:
            path = base / f"{name}.bam"
            subprocess.run([str(synth_bin), "--out", str(path), *map(str, args)], check=True, capture_output=True)
            cache[name] = path
        return str(cache[name])

    return make


@pytest.fixture(scope="session")
def native_gpu():
    from bamnado_b200 import _native

    return _native.lib()


@pytest.fixture(scope="session")
def native_emul():
    """The kernel sources compiled against tests/emul (CPU SIMT emulator) — development aid, never the product."""
    from bamnado_b200 import _native
    from bamnado_b200.build import build_emul

    return _native.Native(ctypes.CDLL(str(build_emul())))


@pytest.fixture(params=[pytest.param("emul", marks=pytest.mark.emul), pytest.param("gpu", marks=pytest.mark.gpu)])
def backend(request):
    """Parity tests run twice: on the emulator (CPU suite, small inputs) and on the GPU (the gate)."""
    return request.getfixturevalue("native_" + request.param)


@pytest.fixture
def small_segments(monkeypatch):
    monkeypatch.setenv("BND_SEGMENT_BYTES", "262144")
