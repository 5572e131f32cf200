"""Host-side checks that need no GPU: the C-ABI library loads and exports every symbol of include/*.h, the product
path fails loudly without a CUDA device (no CPU fallback, no route through oracle/), host logic matches the oracle."""
import ctypes
import re
import shutil
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, TEST_BAM


@pytest.fixture(scope="module")
def product_lib():
    from bamnado_b200 import build

    if shutil.which("nvcc") is None and not build.LIB.exists():
        pytest.skip("nvcc not available and no prebuilt library")
    return build.build()


def declared_symbols():
    text = (ROOT / "include" / "bamnado_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bnd_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol(product_lib):
    lib = ctypes.CDLL(str(product_lib))
    names = declared_symbols()
    assert len(names) >= 28
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bamnado_b200.h but not exported"
    from bamnado_b200 import _native

    assert sorted(_native.EXPORTS) == names
    assert lib.bnd_abi_version() == 4


def test_library_is_sm100a_and_self_contained(product_lib):
    out = subprocess.run(["cuobjdump", "-lelf", str(product_lib)], capture_output=True, text=True)
    if out.returncode == 0:
        assert "sm_100a" in out.stdout
    deps = subprocess.run(["ldd", str(product_lib)], capture_output=True, text=True).stdout
    assert "libz" not in deps, "the product must not inflate on the CPU"
    assert "oracle" not in deps


def test_product_sources_never_touch_the_oracle_or_zlib():
    for p in list((ROOT / "bamnado_b200").rglob("*.py")) + list((ROOT / "bamnado_b200" / "csrc").iterdir()):
        if p.is_file() and p.suffix in (".py", ".cpp", ".cu", ".cuh", ".h"):
            t = p.read_text()
            assert "oracle" not in t.replace("C++ oracle", "").lower() or p.name in ("pileup.cuh",), p
            assert "zlib.h" not in t and "inflateInit" not in t, p


@pytest.mark.skipif(subprocess.run(["bash", "-c", "nvidia-smi -L >/dev/null 2>&1"]).returncode == 0, reason="a GPU is visible")
def test_compute_fails_loudly_without_a_gpu(product_lib):
    from bamnado_b200 import _native

    n = _native.Native(ctypes.CDLL(str(product_lib)))
    assert n.device_count() == 0
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        n.pileup(dict(bam_path=str(TEST_BAM), bin_size=100), {})
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        n.bgzf_inflate(open(TEST_BAM, "rb").read()[:20000])
    # pure arithmetic entry points need no device
    assert n.scale_factor("cpm", 1.0, 1000, 10_000_000) == 0.1


def test_python_surface_mirrors_reference():
    import bamnado_b200 as b

    # the reference module's names for this path, plus the count_bins front half of compute_scale_factors (lib.rs:505)
    assert b.__all__ == ["ReadFilter", "get_signal_for_chromosome", "__version__", "count_bins", "BinCounts"]
    assert callable(b.count_bins)
    f = b.ReadFilter()
    assert (f.min_mapq, f.proper_pair, f.min_length, f.max_length, f.strand) == (0, False, 0, 1000, "both")
    with pytest.raises(ValueError):
        b.ReadFilter(strand="sideways")  # bamnado-python/src/lib.rs:109-116
    assert isinstance(b.__version__(), str)
    from bamnado_b200 import cli

    assert callable(cli.main)


def test_scale_factor_matches_reference_vectors(product_lib):
    from bamnado_b200 import _native

    n = _native.Native(ctypes.CDLL(str(product_lib)))
    # signal_normalization.rs:81-103
    assert 1000.0 * n.scale_factor("raw", 1.0, 1000, 10_000_000) == 1000.0
    assert 1000.0 * n.scale_factor("cpm", 1.0, 1000, 10_000_000) == 100.0
    assert 1000.0 * n.scale_factor("rpkm", 1.0, 1000, 10_000_000) == 100.0
    assert n.scale_factor("rpkm", 2.0, 50, 0) == 0.0


def test_emulated_host_logic_bam_stats(native_emul, oracle, make_bam):
    # header parse (device-inflated), BAI metadata, chunk geometry: bam_utils.rs:311-436
    for bam, b in ((str(TEST_BAM), 100), (str(TEST_BAM), 7), (make_bam("pe", "--pairs", 20000, "--odd"), 50)):
        assert native_emul.bam_stats(bam, b) == oracle.bam_stats(bam, b)
