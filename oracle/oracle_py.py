"""ctypes wrapper around the CPU oracle (oracle/bamnado_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package (bamnado_b200/) never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_build" / "libbamnado_oracle.so"

COUNTER_NAMES = [
    "n_total", "n_failed_proper_pair", "n_failed_mapq", "n_failed_length", "n_failed_blacklist",
    "n_failed_barcode", "n_not_in_read_group", "n_incorrect_strand", "n_failed_tag_filter",
    "n_failed_fragment_length", "n_failed_duplicate", "n_failed_secondary", "n_failed_supplementary",
]
N_COUNTERS = len(COUNTER_NAMES)
NORM = {"raw": 0, "rpkm": 1, "cpm": 2}
STRAND = {"both": 0, "forward": 1, "fwd": 1, "+": 1, "reverse": 2, "rev": 2, "-": 2}


class FilterCfg(C.Structure):
    _fields_ = [
        ("strand", C.c_int32), ("proper_pair", C.c_int32), ("min_mapq", C.c_uint32),
        ("min_length", C.c_uint32), ("max_length", C.c_uint32),
        ("has_min_fragment_length", C.c_int32), ("min_fragment_length", C.c_uint32),
        ("has_max_fragment_length", C.c_int32), ("max_fragment_length", C.c_uint32),
        ("exclude_duplicates", C.c_int32), ("exclude_secondary", C.c_int32), ("exclude_supplementary", C.c_int32),
        ("read_group", C.c_char_p), ("filter_tag", C.c_char_p), ("filter_tag_value", C.c_char_p),
        ("blacklist_bed", C.c_char_p), ("blacklist_merge_overlaps", C.c_int32),
        ("barcode_csv", C.c_char_p), ("barcode_csv_column", C.c_int32),
        ("has_barcode_list", C.c_int32), ("n_barcodes", C.c_uint64), ("barcodes", C.POINTER(C.c_char_p)),
    ]


class PileupCfg(C.Structure):
    _fields_ = [
        ("bam_path", C.c_char_p), ("bin_size", C.c_uint64), ("norm_method", C.c_int32),
        ("scale_factor", C.c_float), ("use_fragment", C.c_int32), ("collapse", C.c_int32),
        ("ignore_scaffold", C.c_int32), ("has_shift", C.c_int32), ("shift", C.c_int32 * 4),
        ("has_truncate", C.c_int32), ("trunc_has5", C.c_int32), ("trunc5", C.c_uint64),
        ("trunc_has3", C.c_int32), ("trunc3", C.c_uint64), ("n_threads", C.c_int32), ("faithful_io", C.c_int32),
    ]


class Run(C.Structure):
    _fields_ = [("ref_id", C.c_int32), ("val", C.c_uint32), ("start", C.c_uint64), ("stop", C.c_uint64)]


RUN_DTYPE = np.dtype([("ref_id", "<i4"), ("val", "<u4"), ("start", "<u8"), ("stop", "<u8")])


def build(force: bool = False) -> Path:
    """Compile the oracle with its Makefile (gcc + zlib)."""
    src = HERE / "bamnado_oracle.cpp"
    if force or not LIB_PATH.exists() or LIB_PATH.stat().st_mtime < max(p.stat().st_mtime for p in (src, HERE / "bamnado_oracle.h", HERE / "bam_writers.inl")):
        subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB_PATH))
        L.orc_last_error.restype = C.c_char_p
        L.orc_scale_factor.restype = C.c_double
        L.orc_scale_factor.argtypes = [C.c_int32, C.c_float, C.c_uint64, C.c_uint64]
        L.orc_collapse_equal_bins.restype = C.c_uint64
        L.orc_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


class OracleError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(msg)
        self.code = code


def _check(rc):
    if rc != 0:
        raise OracleError(rc, lib().orc_last_error().decode())


def _b(s):
    return None if s is None else os.fsencode(s) if isinstance(s, (str, Path)) else s


def make_filter_cfg(strand="both", proper_pair=False, min_mapq=0, min_length=0, max_length=0xFFFFFFFF,
                    min_fragment_length=None, max_fragment_length=None, exclude_duplicates=False,
                    exclude_secondary=False, exclude_supplementary=False, read_group=None, filter_tag=None,
                    filter_tag_value=None, blacklist_bed=None, blacklist_merge_overlaps=False, barcode_csv=None,
                    barcode_csv_column=-1, barcodes=None, cls=FilterCfg):
    f = cls()
    f.strand = STRAND[strand]
    f.proper_pair = int(proper_pair)
    f.min_mapq, f.min_length, f.max_length = min_mapq, min_length, max_length
    f.has_min_fragment_length = int(min_fragment_length is not None)
    f.min_fragment_length = min_fragment_length or 0
    f.has_max_fragment_length = int(max_fragment_length is not None)
    f.max_fragment_length = max_fragment_length or 0
    f.exclude_duplicates, f.exclude_secondary, f.exclude_supplementary = int(exclude_duplicates), int(exclude_secondary), int(exclude_supplementary)
    f.read_group, f.filter_tag, f.filter_tag_value = _b(read_group), _b(filter_tag), _b(filter_tag_value)
    f.blacklist_bed = _b(blacklist_bed)
    f.blacklist_merge_overlaps = int(blacklist_merge_overlaps)
    f.barcode_csv = _b(barcode_csv)
    f.barcode_csv_column = barcode_csv_column
    if barcodes is not None:
        arr = (C.c_char_p * max(len(barcodes), 1))(*[_b(x) for x in barcodes])
        f._keep = arr
        f.has_barcode_list = 1
        f.n_barcodes = len(barcodes)
        f.barcodes = C.cast(arr, C.POINTER(C.c_char_p))
    return f


def make_pileup_cfg(bam_path, bin_size=50, norm="raw", scale_factor=1.0, use_fragment=False, collapse=True,
                    ignore_scaffold=False, shift=None, truncate=None, n_threads=0, faithful_io=False, cls=PileupCfg):
    p = cls()
    p.bam_path = _b(bam_path)
    p.bin_size = bin_size
    p.norm_method = NORM[norm]
    p.scale_factor = scale_factor
    p.use_fragment, p.collapse, p.ignore_scaffold = int(use_fragment), int(collapse), int(ignore_scaffold)
    if shift is not None:
        p.has_shift = 1
        for i in range(4):
            p.shift[i] = shift[i]
    if truncate is not None:
        p.has_truncate = 1
        p.trunc_has5, p.trunc5 = int(truncate[0] is not None), truncate[0] or 0
        p.trunc_has3, p.trunc3 = int(truncate[1] is not None), truncate[1] or 0
    p.n_threads = n_threads
    if hasattr(p, "faithful_io"):
        p.faithful_io = int(faithful_io)
    return p


def _runs_to_np(ptr, n):
    if n == 0:
        arr = np.zeros(0, dtype=RUN_DTYPE)
    else:
        arr = np.frombuffer(C.string_at(ptr, n * C.sizeof(Run)), dtype=RUN_DTYPE).copy()
    lib().orc_free(ptr)
    return arr


def stats_dict(stats):
    return {k: int(v) for k, v in zip(COUNTER_NAMES, stats)}


def pileup_runs(pileup: dict, filt: dict):
    """-> (runs structured array sorted by (ref_id,start), stats dict, normalisation factor)"""
    p, f = make_pileup_cfg(**pileup), make_filter_cfg(**filt)
    runs = C.POINTER(Run)()
    n = C.c_uint64()
    stats = (C.c_uint64 * N_COUNTERS)()
    factor = C.c_double()
    _check(lib().orc_pileup_runs(C.byref(p), C.byref(f), C.byref(runs), C.byref(n), stats, C.byref(factor)))
    return _runs_to_np(runs, n.value), stats_dict(stats), factor.value


def pileup_chromosome(pileup: dict, filt: dict, chrom: str):
    p, f = make_pileup_cfg(**pileup), make_filter_cfg(**filt)
    runs = C.POINTER(Run)()
    n = C.c_uint64()
    stats = (C.c_uint64 * N_COUNTERS)()
    _check(lib().orc_pileup_chromosome(C.byref(p), C.byref(f), _b(chrom), C.byref(runs), C.byref(n), stats))
    return _runs_to_np(runs, n.value), stats_dict(stats)


def signal_for_chromosome(pileup: dict, filt: dict, chrom: str) -> np.ndarray:
    p, f = make_pileup_cfg(**pileup), make_filter_cfg(**filt)
    out = C.POINTER(C.c_float)()
    n = C.c_uint64()
    rc = lib().orc_signal_for_chromosome(C.byref(p), C.byref(f), _b(chrom), C.byref(out), C.byref(n))
    if rc == 2:
        raise KeyError(lib().orc_last_error().decode())
    _check(rc)
    arr = np.ctypeslib.as_array(out, shape=(n.value,)).copy() if n.value else np.zeros(0, np.float32)
    lib().orc_free(out)
    return arr


def pileup_to_bedgraph(pileup: dict, filt: dict, out_path):
    p, f = make_pileup_cfg(**pileup), make_filter_cfg(**filt)
    stats = (C.c_uint64 * N_COUNTERS)()
    factor = C.c_double()
    _check(lib().orc_pileup_to_bedgraph(C.byref(p), C.byref(f), _b(out_path), stats, C.byref(factor)))
    return stats_dict(stats), factor.value


def pileup_chroms_to_bedgraph(pileup: dict, filt: dict, chroms: list, out_path):
    """bench helper: pileup + bedGraph of a few chromosomes only -> (stats dict, factor)"""
    p, f = make_pileup_cfg(**pileup), make_filter_cfg(**filt)
    arr = (C.c_char_p * len(chroms))(*[_b(c) for c in chroms])
    stats = (C.c_uint64 * N_COUNTERS)()
    factor = C.c_double()
    _check(lib().orc_pileup_chroms_to_bedgraph(C.byref(p), C.byref(f), arr, len(chroms), _b(out_path), stats, C.byref(factor)))
    return stats_dict(stats), factor.value


def multi_to_tsv(pileups: list, filters: list, out_path):
    n = len(pileups)
    ps = (PileupCfg * n)(*[make_pileup_cfg(**p) for p in pileups])
    fs = (FilterCfg * n)(*[make_filter_cfg(**f) for f in filters])
    _check(lib().orc_multi_to_tsv(ps, fs, n, _b(out_path)))


def collapse_bedgraph(in_path, out_path):
    _check(lib().orc_collapse_bedgraph(_b(in_path), _b(out_path)))


def scale_factor(method: str, base_scale: float, bin_size: int, n_reads: int) -> float:
    return lib().orc_scale_factor(NORM[method], base_scale, bin_size, n_reads)


def bam_stats(bam_path, bin_size):
    vals = [C.c_uint64() for _ in range(5)]
    nrefs = C.c_int32()
    _check(lib().orc_bam_stats(_b(bam_path), C.c_uint64(bin_size), *[C.byref(v) for v in vals], C.byref(nrefs)))
    keys = ["n_mapped", "n_unmapped", "genome_len", "chunk_len", "n_chunks"]
    d = {k: v.value for k, v in zip(keys, vals)}
    d["n_refs"] = nrefs.value
    return d


def format_f64(v: float) -> str:
    buf = C.create_string_buffer(64)
    n = lib().orc_format_f64(C.c_double(v), buf, 64)
    return buf.raw[:n].decode()


def collapse_equal_bins(chrom, start, end, scores):
    chrom = np.ascontiguousarray(chrom, dtype=np.int32).copy()
    start = np.ascontiguousarray(start, dtype=np.uint64).copy()
    end = np.ascontiguousarray(end, dtype=np.uint64).copy()
    scores = np.ascontiguousarray(scores, dtype=np.uint32).copy()
    if scores.ndim == 1:
        scores = scores[:, None].copy()
    n, k = scores.shape
    m = lib().orc_collapse_equal_bins(chrom.ctypes.data_as(C.c_void_p), start.ctypes.data_as(C.c_void_p),
                                      end.ctypes.data_as(C.c_void_p), scores.ctypes.data_as(C.c_void_p),
                                      C.c_uint64(n), C.c_uint32(k))
    return chrom[:m], start[:m], end[:m], scores.reshape(-1)[: m * k].reshape(m, k)


def count_records(bam_path):
    a, b, c = C.c_uint64(), C.c_uint64(), C.c_uint64()
    _check(lib().orc_count_records(_b(bam_path), C.byref(a), C.byref(b), C.byref(c)))
    return {"n_records": a.value, "inflated_bytes": b.value, "n_blocks": c.value}


def bam_header_refs(bam_path):
    """(names, lengths) of the @SQ dictionary in the binary BAM header (SAM spec 4.2), read through Python's gzip."""
    import gzip
    import struct

    with gzip.open(bam_path, "rb") as f:
        if f.read(4) != b"BAM\1":
            raise OracleError(1, "not a BAM file")
        (l_text,) = struct.unpack("<i", f.read(4))
        f.read(l_text)
        (n_ref,) = struct.unpack("<i", f.read(4))
        names, lens = [], []
        for _ in range(n_ref):
            (l_name,) = struct.unpack("<i", f.read(4))
            names.append(f.read(l_name)[:-1].decode())
            lens.append(struct.unpack("<i", f.read(4))[0])
    return names, lens


def count_bins(bam_paths, bin_size, filters, use_fragment=False, ignore_scaffold=False):
    """Restatement of bin_counts::count_bins (bamnado/src/bin_counts.rs:100-289) on top of the oracle's pile-up.

    Grid (:146-170): chromosomes of the first BAM's stats sorted by name (scaffolds dropped on request), bins
    [start, min(start + bin - 1, len)] while start <= len.  Per sample (:178-262): the chunk list without scaffold
    chunks, `bin_intervals(.., collapse=false)` per chunk, every interval's count written to row
    offset[chrom] + (start - 1) // bin.  library size = n_reads_after_filtering of that sample's filter (:264).
    Returns (counts[n_bins, n_samples] u64, [(chrom, len, first_row)], library_sizes).
    """
    import re

    scaffold = re.compile(r"^chrUn|_alt$|_random$")  # bam_utils.rs:38-43
    sizes0 = dict(zip(*bam_header_refs(bam_paths[0])))
    names = sorted(n for n in sizes0 if not (ignore_scaffold and scaffold.search(n)))
    grid, rows = [], 0
    for n in names:
        grid.append((n, sizes0[n], rows))
        rows += -(-sizes0[n] // bin_size)
    if rows == 0:
        raise OracleError(1, "No bins produced; check --bin-size and --ignore-scaffolds")
    offset = {n: r for n, _, r in grid}
    counts = np.zeros((rows, len(bam_paths)), dtype=np.uint64)
    libs = []
    for s, (bam, filt) in enumerate(zip(bam_paths, filters)):
        ref_names, ref_lens = bam_header_refs(bam)
        if dict(zip(ref_names, ref_lens)) != sizes0:
            raise OracleError(1, "Chromosome name/length mismatch between BAM files")
        keep = [n for n in ref_names if not (ignore_scaffold and scaffold.search(n))]
        lib = 0
        # the reference's chunk list minus the scaffold chunks == the pile-up restricted to the kept chromosomes; the
        # filter counters only see reads of visited chunks, so they are summed over those chromosomes
        for chrom in keep:
            runs, sd = pileup_chromosome(dict(bam_path=bam, bin_size=bin_size, norm="raw", use_fragment=use_fragment, collapse=False), filt, chrom)
            lib += sd["n_total"] - sum(v for k, v in sd.items() if k != "n_total")
            if len(runs):
                idx = offset[chrom] + (runs["start"].astype(np.int64) - 1) // bin_size
                counts[idx, s] = runs["val"]
        libs.append(int(lib))
    return counts, grid, libs


# ---- BAM-writing tools (bam_writers.inl) -----------------------------------------------------------------------------------
SPLIT_STATS = ["n_total_reads", "n_unmapped_reads", "n_qcfail_reads", "n_duplicate_reads", "n_secondary_reads", "n_filtered_reads", "n_low_maq",
               "n_both_genomes", "n_exogenous", "n_endogenous"]


def bam_split(bam, out_path, filt=None, command_line=""):
    n = C.c_uint64(0)
    f = make_filter_cfg(**(filt or {}))
    _check(lib().orc_bam_split(_b(bam), C.byref(f), _b(out_path), _b(command_line), C.byref(n)))
    return int(n.value)


def bam_modify(bam, out_path, filt=None, tn5_shift=False, command_line=""):
    n = C.c_uint64(0)
    f = make_filter_cfg(**(filt or {}))
    _check(lib().orc_bam_modify(_b(bam), C.byref(f), int(bool(tn5_shift)), _b(out_path), _b(command_line), C.byref(n)))
    return int(n.value)


def bam_split_exogenous(bam, paths, exogenous_prefix, filt=None, command_line=""):
    """paths: [endogenous, exogenous, both, unmapped] -> SplitStats counters"""
    f = make_filter_cfg(**(filt or {}))
    arr = (C.c_char_p * 4)(*[_b(p) for p in paths])
    st = (C.c_uint64 * 10)()
    _check(lib().orc_bam_split_exogenous(_b(bam), C.byref(f), _b(exogenous_prefix), arr, _b(command_line), st))
    return {k: int(st[i]) for i, k in enumerate(SPLIT_STATS)}
