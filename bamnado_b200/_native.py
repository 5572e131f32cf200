"""ctypes binding of libbamnado_b200.so (the C ABI in include/bamnado_b200.h).

The library is CUDA-only: there is no CPU fallback.  `lib()` loads the in-tree sm_100a build and raises if it is
missing; every compute call raises RuntimeError on a machine without a CUDA device.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_lib" / "libbamnado_b200.so"

COUNTER_NAMES = [
    "n_total", "n_failed_proper_pair", "n_failed_mapq", "n_failed_length", "n_failed_blacklist",
    "n_failed_barcode", "n_not_in_read_group", "n_incorrect_strand", "n_failed_tag_filter",
    "n_failed_fragment_length", "n_failed_duplicate", "n_failed_secondary", "n_failed_supplementary",
]
N_COUNTERS = len(COUNTER_NAMES)
NORM = {"raw": 0, "rpkm": 1, "cpm": 2}
STRAND = {"both": 0, "forward": 1, "reverse": 2}
ERR_RUNTIME, ERR_NOT_FOUND, ERR_INVALID = 1, 2, 3

EXPORTS = [
    "bnd_last_error", "bnd_last_error_class", "bnd_free", "bnd_abi_version", "bnd_device_count", "bnd_launch_count", "bnd_pileup_create",
    "bnd_pileup_destroy", "bnd_pileup_to_bedgraph", "bnd_pileup_to_bigwig", "bnd_pileup_chromosome", "bnd_pileup_runs",
    "bnd_pileup_norm_factor", "bnd_pileup_stats", "bnd_pileup_accumulate", "bnd_pileup_stage",
    "bnd_pileup_accumulate_resident", "bnd_pileup_set_total_stats", "bnd_pileup_finalize", "bnd_pileup_bedgraph_text",
    "bnd_pileup_timings", "bnd_pileup_shape", "bnd_pileup_preallocate_output", "bnd_pileup_preallocated_bytes", "bnd_pileup_format_bedgraph", "bnd_pileup_write_bedgraph_pieces", "bnd_signal_for_chromosome", "bnd_multi_to_tsv", "bnd_bin_counts", "bnd_multi_prepare",
    "bnd_multi_prepare_resident", "bnd_multi_change_bits", "bnd_multi_or_bits", "bnd_multi_compact_columns", "bnd_multi_format_tsv", "bnd_multi_write_tsv_pieces", "bnd_pileup_refs", "bnd_collapse_bedgraph", "bnd_bigwig_compare", "bnd_bigwig_aggregate", "bnd_bigwig_from_intervals", "bnd_bam_split", "bnd_bam_modify", "bnd_bam_split_exogenous",
    "bnd_scale_factor", "bnd_bam_stats", "bnd_cli_main", "bnd_bgzf_inflate", "bnd_bgzf_compress",
]


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
        ("trunc_has3", C.c_int32), ("trunc3", C.c_uint64), ("n_threads", C.c_int32), ("device", C.c_int32),
        ("shard_rank", C.c_int32), ("shard_world", C.c_int32),
    ]


class Run(C.Structure):
    _fields_ = [("ref_id", C.c_int32), ("val", C.c_uint32), ("start", C.c_uint64), ("stop", C.c_uint64)]


RUN_DTYPE = np.dtype([("ref_id", "<i4"), ("val", "<u4"), ("start", "<u8"), ("stop", "<u8")])


def _b(s):
    return None if s is None else os.fsencode(s) if isinstance(s, (str, Path)) else s


def make_filter_cfg(strand="both", proper_pair=False, min_mapq=0, min_length=0, max_length=0xFFFFFFFF,
                    min_fragment_length=None, max_fragment_length=None, exclude_duplicates=False,
                    exclude_secondary=False, exclude_supplementary=False, read_group=None, filter_tag=None,
                    filter_tag_value=None, blacklist_bed=None, blacklist_merge_overlaps=False, barcode_csv=None,
                    barcode_csv_column=-1, barcodes=None) -> FilterCfg:
    if strand not in STRAND:
        raise ValueError(f"Invalid strand: {strand}")
    f = FilterCfg()
    f.strand = STRAND[strand]
    f.proper_pair = int(bool(proper_pair))
    f.min_mapq, f.min_length, f.max_length = int(min_mapq), int(min_length), int(max_length)
    f.has_min_fragment_length = int(min_fragment_length is not None)
    f.min_fragment_length = int(min_fragment_length or 0)
    f.has_max_fragment_length = int(max_fragment_length is not None)
    f.max_fragment_length = int(max_fragment_length or 0)
    f.exclude_duplicates, f.exclude_secondary, f.exclude_supplementary = int(bool(exclude_duplicates)), int(bool(exclude_secondary)), int(bool(exclude_supplementary))
    f.read_group, f.filter_tag, f.filter_tag_value = _b(read_group), _b(filter_tag), _b(filter_tag_value)
    f.blacklist_bed = _b(blacklist_bed)
    f.blacklist_merge_overlaps = int(bool(blacklist_merge_overlaps))
    f.barcode_csv = _b(barcode_csv)
    f.barcode_csv_column = int(barcode_csv_column)
    if barcodes is not None:
        arr = (C.c_char_p * max(len(barcodes), 1))(*[_b(x) for x in barcodes])
        f._keep = arr
        f.has_barcode_list = 1
        f.n_barcodes = len(barcodes)
        f.barcodes = C.cast(arr, C.POINTER(C.c_char_p))
    return f


def make_pileup_cfg(bam_path, bin_size=50, norm="raw", scale_factor=1.0, use_fragment=False, collapse=True,
                    ignore_scaffold=False, shift=None, truncate=None, n_threads=0, device=-1, shard_rank=0,
                    shard_world=1) -> PileupCfg:
    p = PileupCfg()
    p.bam_path = _b(bam_path)
    p.bin_size = int(bin_size)
    p.norm_method = NORM[norm]
    p.scale_factor = float(scale_factor)
    p.use_fragment, p.collapse, p.ignore_scaffold = int(bool(use_fragment)), int(bool(collapse)), int(bool(ignore_scaffold))
    if shift is not None:
        p.has_shift = 1
        for i in range(4):
            p.shift[i] = int(shift[i])
    if truncate is not None:
        p.has_truncate = 1
        p.trunc_has5, p.trunc5 = int(truncate[0] is not None), int(truncate[0] or 0)
        p.trunc_has3, p.trunc3 = int(truncate[1] is not None), int(truncate[1] or 0)
    p.n_threads, p.device, p.shard_rank, p.shard_world = int(n_threads), int(device), int(shard_rank), int(shard_world)
    return p


class Native:
    """The C ABI bound to one loaded shared library."""

    def __init__(self, cdll: C.CDLL):
        L = self.L = cdll
        missing = [s for s in EXPORTS if not hasattr(L, s)]
        if missing:
            raise RuntimeError(f"libbamnado_b200 is missing symbols: {missing}")
        L.bnd_last_error.restype = C.c_char_p
        L.bnd_free.argtypes = [C.c_void_p]
        L.bnd_launch_count.restype = C.c_uint64
        L.bnd_pileup_create.restype = C.c_void_p
        L.bnd_pileup_create.argtypes = [C.POINTER(PileupCfg), C.POINTER(FilterCfg)]
        L.bnd_pileup_destroy.argtypes = [C.c_void_p]
        for name in ("bnd_pileup_to_bedgraph", "bnd_pileup_to_bigwig"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_char_p]
        L.bnd_pileup_chromosome.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.POINTER(Run)), C.POINTER(C.c_uint64)]
        L.bnd_pileup_runs.argtypes = [C.c_void_p, C.POINTER(C.POINTER(Run)), C.POINTER(C.c_uint64)]
        L.bnd_pileup_norm_factor.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.bnd_pileup_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        for name in ("bnd_pileup_accumulate", "bnd_pileup_stage", "bnd_pileup_accumulate_resident"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.bnd_pileup_set_total_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.bnd_pileup_finalize.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.bnd_pileup_bedgraph_text.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.bnd_pileup_timings.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.bnd_pileup_shape.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.bnd_pileup_preallocate_output.argtypes = [C.c_void_p, C.c_char_p]
        L.bnd_pileup_format_bedgraph.argtypes = [C.c_void_p, C.POINTER(C.c_uint64), C.c_uint64]
        L.bnd_pileup_preallocated_bytes.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.bnd_pileup_write_bedgraph_pieces.argtypes = [C.c_void_p, C.c_char_p, C.POINTER(C.c_uint64), C.c_int32, C.c_int32, C.c_uint64]
        L.bnd_signal_for_chromosome.argtypes = [C.POINTER(PileupCfg), C.POINTER(FilterCfg), C.c_char_p,
                                                C.POINTER(C.POINTER(C.c_float)), C.POINTER(C.c_uint64)]
        L.bnd_multi_to_tsv.argtypes = [C.POINTER(PileupCfg), C.POINTER(FilterCfg), C.c_int32, C.c_char_p]
        L.bnd_multi_prepare.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.bnd_multi_prepare_resident.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        L.bnd_multi_change_bits.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_void_p]
        L.bnd_multi_or_bits.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64]
        L.bnd_multi_compact_columns.argtypes = [C.POINTER(C.c_void_p), C.c_int32, C.c_void_p, C.POINTER(C.c_uint64), C.POINTER(C.c_void_p)]
        L.bnd_multi_format_tsv.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int32, C.POINTER(C.c_int32), C.c_uint64, C.c_uint64, C.c_int32, C.POINTER(C.c_uint64), C.c_uint64]
        L.bnd_multi_write_tsv_pieces.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_uint64), C.c_int32, C.c_int32]
        L.bnd_pileup_refs.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.POINTER(C.c_uint64)), C.POINTER(C.c_uint64)]
        L.bnd_collapse_bedgraph.argtypes = [C.c_char_p, C.c_char_p]
        dp = C.POINTER(C.c_double)
        L.bnd_bigwig_compare.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32, C.c_uint32, dp, dp, dp]
        L.bnd_bigwig_aggregate.argtypes = [C.POINTER(C.c_char_p), C.c_int32, C.c_char_p, C.c_int32, C.c_uint32, dp, dp]
        L.bnd_bigwig_from_intervals.argtypes = [C.POINTER(C.c_char_p), C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_char_p]
        u64p = C.POINTER(C.c_uint64)
        L.bnd_bgzf_compress.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_void_p), u64p]
        L.bnd_bam_split.argtypes = [C.c_char_p, C.POINTER(FilterCfg), C.c_char_p, C.c_char_p, C.c_int32, u64p]
        L.bnd_bam_modify.argtypes = [C.c_char_p, C.POINTER(FilterCfg), C.c_int32, C.c_char_p, C.c_char_p, C.c_int32, u64p]
        L.bnd_bam_split_exogenous.argtypes = [C.c_char_p, C.POINTER(FilterCfg), C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32, u64p]
        L.bnd_scale_factor.restype = C.c_double
        L.bnd_scale_factor.argtypes = [C.c_int32, C.c_float, C.c_uint64, C.c_uint64]
        L.bnd_bam_stats.argtypes = [C.c_char_p, C.c_uint64, C.POINTER(C.c_uint64)]
        L.bnd_cli_main.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        L.bnd_bgzf_inflate.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64), C.c_void_p,
                                       C.c_uint64, C.POINTER(C.c_uint64)]

    # ---- errors -------------------------------------------------------------------------------------------------
    def check(self, rc: int):
        if rc == 0:
            return
        msg = (self.L.bnd_last_error() or b"").decode(errors="replace")
        cls = self.L.bnd_last_error_class()
        if cls == ERR_NOT_FOUND:
            raise KeyError(msg)
        if cls == ERR_INVALID:
            raise ValueError(msg)
        raise RuntimeError(msg)

    # ---- plain functions ----------------------------------------------------------------------------------------
    def device_count(self) -> int:
        return self.L.bnd_device_count()

    def launch_count(self) -> int:
        return int(self.L.bnd_launch_count())

    def scale_factor(self, method: str, base_scale: float, bin_size: int, n_reads: int) -> float:
        return self.L.bnd_scale_factor(NORM[method], base_scale, bin_size, n_reads)

    def bam_stats(self, bam_path, bin_size: int) -> dict:
        out = (C.c_uint64 * 6)()
        self.check(self.L.bnd_bam_stats(_b(bam_path), bin_size, out))
        return dict(zip(["n_mapped", "n_unmapped", "genome_len", "chunk_len", "n_chunks", "n_refs"], [int(v) for v in out]))

    def bgzf_compress(self, data: bytes) -> bytes:
        """The device BGZF compressor on its own: BGZF blocks + EOF marker."""
        out, n = C.c_void_p(), C.c_uint64()
        self.check(self.L.bnd_bgzf_compress(data, len(data), C.byref(out), C.byref(n)))
        try:
            return C.string_at(out, n.value)
        finally:
            self.L.bnd_free(out)

    def bgzf_inflate(self, data: bytes):
        """K1 on its own: returns (inflated bytes, per-block status array)."""
        n = len(data)
        src = np.frombuffer(data, dtype=np.uint8)
        cap, nblk, off = 0, 0, 0
        while off + 18 <= n:  # BSIZE / ISIZE walk (standard BGZF layout), only to size the output buffers
            bsize = int(src[off + 16]) + (int(src[off + 17]) << 8) + 1
            if off + bsize > n:
                break
            cap += int.from_bytes(src[off + bsize - 4: off + bsize].tobytes(), "little")
            nblk += 1
            off += bsize
        out = np.empty(max(cap, 1), dtype=np.uint8)
        st = np.empty(max(nblk, 1), dtype=np.int32)
        out_len, n_blocks = C.c_uint64(), C.c_uint64()
        self.check(self.L.bnd_bgzf_inflate(src.ctypes.data, n, out.ctypes.data, cap, C.byref(out_len), st.ctypes.data, len(st), C.byref(n_blocks)))
        return out[: out_len.value].tobytes(), st[: n_blocks.value].copy()

    def signal_for_chromosome(self, pileup: dict, filt: dict, chrom: str) -> np.ndarray:
        p, f = make_pileup_cfg(**pileup), make_filter_cfg(**filt)
        out = C.POINTER(C.c_float)()
        n = C.c_uint64()
        self.check(self.L.bnd_signal_for_chromosome(C.byref(p), C.byref(f), _b(chrom), C.byref(out), C.byref(n)))
        arr = np.ctypeslib.as_array(out, shape=(n.value,)).copy() if n.value else np.zeros(0, np.float32)
        self.L.bnd_free(out)
        return arr

    def multi_change_bits(self, handles: list, dev_ptr: int):
        """dev_ptr: ceil(n_bins / 32) u32 words of device memory; bit i = bin i opens a new row for one of these BAMs."""
        hs = (C.c_void_p * len(handles))(*[h.h for h in handles])
        self.check(self.L.bnd_multi_change_bits(hs, len(handles), C.c_void_p(dev_ptr)))

    def multi_compact_columns(self, handles: list, dev_bits_ptr: int):
        """-> (n_rows, device pointer of u32[len(handles)][n_rows]: the BAMs' counts at the shared row heads; owned by handles[0])."""
        hs = (C.c_void_p * len(handles))(*[h.h for h in handles])
        n, vals = C.c_uint64(), C.c_void_p()
        self.check(self.L.bnd_multi_compact_columns(hs, len(handles), C.c_void_p(dev_bits_ptr), C.byref(n), C.byref(vals)))
        return n.value, vals.value or 0

    def multi_or_bits(self, device: int, dst_ptr: int, src_ptr: int, n_src: int, n_words: int):
        """dst[w] = OR over n_src masks of n_words u32 laid out back to back at src (device pointers; dst may alias mask 0)."""
        self.check(self.L.bnd_multi_or_bits(device, C.c_void_p(dst_ptr), C.c_void_p(src_ptr), n_src, n_words))

    def multi_to_tsv(self, pileups: list, filters: list, out_path):
        n = len(pileups)
        ps = (PileupCfg * n)(*[make_pileup_cfg(**p) for p in pileups])
        fs = (FilterCfg * n)(*[make_filter_cfg(**f) for f in filters])
        self.check(self.L.bnd_multi_to_tsv(ps, fs, n, _b(out_path)))

    def bin_counts(self, pileups: list, filters: list):
        """bin_counts::count_bins: dict(counts (n_bins, n_samples) u64, chroms, chrom_len, chrom_first_row, library_sizes)."""
        n = len(pileups)
        if n != len(filters):
            raise ValueError(f"Number of BAM paths ({n}) does not match number of filters ({len(filters)})")
        ps = (PileupCfg * max(n, 1))(*[make_pileup_cfg(**p) for p in pileups])
        fs = (FilterCfg * max(n, 1))(*[make_filter_cfg(**f) for f in filters])
        counts, names = C.POINTER(C.c_uint64)(), C.c_char_p()
        clen, crow = C.POINTER(C.c_uint64)(), C.POINTER(C.c_uint64)()
        n_bins, n_chroms = C.c_uint64(), C.c_uint64()
        lib = (C.c_uint64 * max(n, 1))()
        names_p = C.c_void_p()
        self.check(self.L.bnd_bin_counts(ps, fs, n, C.byref(counts), C.byref(n_bins), C.byref(names_p), C.byref(clen), C.byref(crow),
                                         C.byref(n_chroms), lib))
        nb, nc = n_bins.value, n_chroms.value
        out = dict(
            counts=np.ctypeslib.as_array(counts, shape=(nb * n,)).reshape(nb, n).copy() if nb else np.zeros((0, n), np.uint64),
            chroms=C.string_at(names_p).decode().split("\n") if nc else [],
            chrom_len=np.ctypeslib.as_array(clen, shape=(nc,)).copy() if nc else np.zeros(0, np.uint64),
            chrom_first_row=np.ctypeslib.as_array(crow, shape=(nc + 1,)).copy(),
            library_sizes=[int(lib[i]) for i in range(n)],
        )
        for ptr in (counts, names_p, clen, crow):
            self.L.bnd_free(C.cast(ptr, C.c_void_p))
        return out

    def collapse_bedgraph(self, in_path, out_path):
        self.check(self.L.bnd_collapse_bedgraph(_b(in_path), _b(out_path)))

    COMPARISONS = {"subtraction": 0, "ratio": 1, "log-ratio": 2}
    AGGREGATIONS = {"sum": 0, "mean": 1, "max": 2, "min": 3, "median": 4}

    def bigwig_compare(self, bw1, bw2, out_path, comparison: str, bin_size: int = 50, pseudocount=None, scale_factor_bw1=None, scale_factor_bw2=None):
        opt = lambda v: None if v is None else C.byref(C.c_double(v))  # noqa: E731
        self.check(self.L.bnd_bigwig_compare(_b(bw1), _b(bw2), _b(out_path), self.COMPARISONS[comparison], bin_size, opt(pseudocount), opt(scale_factor_bw1),
                                             opt(scale_factor_bw2)))

    def bigwig_aggregate(self, bigwigs: list, out_path, method: str, bin_size: int = 50, pseudocount=None, scale_factors=None):
        paths = (C.c_char_p * len(bigwigs))(*[_b(p) for p in bigwigs])
        sf = (C.c_double * len(scale_factors))(*scale_factors) if scale_factors is not None else None
        pc = None if pseudocount is None else C.byref(C.c_double(pseudocount))
        self.check(self.L.bnd_bigwig_aggregate(paths, len(bigwigs), _b(out_path), self.AGGREGATIONS[method], bin_size, pc, sf))

    def bigwig_from_intervals(self, chroms: list, chrom_idx, start, end, value, out_path):
        """chroms: [(name, length)]; the four arrays describe intervals sorted by (chromosome index, start)."""
        names = (C.c_char_p * len(chroms))(*[c[0].encode() for c in chroms])
        lens = np.ascontiguousarray([c[1] for c in chroms], dtype=np.uint32)
        ci, st, en = (np.ascontiguousarray(x, dtype=np.uint32) for x in (chrom_idx, start, end))
        va = np.ascontiguousarray(value, dtype=np.float32)
        self.check(self.L.bnd_bigwig_from_intervals(names, lens.ctypes.data, len(chroms), ci.ctypes.data, st.ctypes.data, en.ctypes.data, va.ctypes.data, len(ci), _b(out_path)))

    BAM_WRITE_COUNTERS = ["n_total_reads", "n_unmapped_reads", "n_qcfail_reads", "n_duplicate_reads", "n_secondary_reads", "n_filtered_reads", "n_low_maq",
                          "n_both_genomes", "n_exogenous", "n_endogenous", "n_written", "n_dropped_by_shift", "n_errors"]

    def _bam_counters(self, arr):
        return {k: int(arr[i]) for i, k in enumerate(self.BAM_WRITE_COUNTERS)}

    def bam_split(self, bam, out_path, filt: dict | None = None, command_line: str = "", device: int = -1):
        """`bamnado split` (BamFilterer): the records the filter keeps, written as one BAM."""
        f, cnt = make_filter_cfg(**(filt or {})), (C.c_uint64 * 16)()
        self.check(self.L.bnd_bam_split(_b(bam), C.byref(f), _b(out_path), _b(command_line), device, cnt))
        return self._bam_counters(cnt)

    def bam_modify(self, bam, out_path, filt: dict | None = None, tn5_shift: bool = False, command_line: str = "", device: int = -1):
        """`bamnado modify` (BamModifier): filter, optionally the Tn5 shift of every properly paired record."""
        f, cnt = make_filter_cfg(**(filt or {})), (C.c_uint64 * 16)()
        self.check(self.L.bnd_bam_modify(_b(bam), C.byref(f), int(bool(tn5_shift)), _b(out_path), _b(command_line), device, cnt))
        return self._bam_counters(cnt)

    def bam_split_exogenous(self, bam, out_prefix, exogenous_prefix: str, filt: dict | None = None, command_line: str = "", device: int = -1):
        """`bamnado split-exogenous` (BamSplitter): <prefix>.{endogenous,exogenous,both,unmapped}.bam + SplitStats counters."""
        f, cnt = make_filter_cfg(**(filt or {})), (C.c_uint64 * 16)()
        self.check(self.L.bnd_bam_split_exogenous(_b(bam), C.byref(f), _b(exogenous_prefix), _b(out_prefix), _b(command_line), device, cnt))
        return self._bam_counters(cnt)

    def cli_main(self, argv: list) -> int:
        arr = (C.c_char_p * len(argv))(*[_b(a) for a in argv])
        return self.L.bnd_cli_main(len(argv), arr)

    def pileup(self, pileup: dict, filt: dict | None = None) -> "PileupHandle":
        return PileupHandle(self, pileup, filt or {})


class PileupHandle:
    """Owner of one bnd_pileup (BamPileup in the reference)."""

    def __init__(self, native: Native, pileup: dict, filt: dict):
        self.n = native
        self._cfg, self._filt = make_pileup_cfg(**pileup), make_filter_cfg(**filt)
        self.h = native.L.bnd_pileup_create(C.byref(self._cfg), C.byref(self._filt))
        if not self.h:
            native.check(native.L.bnd_last_error_class() or ERR_RUNTIME)

    def close(self):
        if self.h:
            self.n.L.bnd_pileup_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _rows(self, ptr, n):
        if n:
            raw = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_uint8)), shape=(n * C.sizeof(Run),))
            arr = raw.view(RUN_DTYPE).copy()
        else:
            arr = np.zeros(0, dtype=RUN_DTYPE)
        self.n.L.bnd_free(ptr)
        return arr

    def runs(self) -> np.ndarray:
        ptr, n = C.POINTER(Run)(), C.c_uint64()
        self.n.check(self.n.L.bnd_pileup_runs(self.h, C.byref(ptr), C.byref(n)))
        return self._rows(ptr, n.value)

    def chromosome(self, chrom: str) -> np.ndarray:
        ptr, n = C.POINTER(Run)(), C.c_uint64()
        self.n.check(self.n.L.bnd_pileup_chromosome(self.h, _b(chrom), C.byref(ptr), C.byref(n)))
        return self._rows(ptr, n.value)

    def stats(self) -> dict:
        s = (C.c_uint64 * N_COUNTERS)()
        self.n.check(self.n.L.bnd_pileup_stats(self.h, s))
        return {k: int(v) for k, v in zip(COUNTER_NAMES, s)}

    def set_total_stats(self, values):
        s = (C.c_uint64 * N_COUNTERS)(*[int(v) for v in values])
        self.n.check(self.n.L.bnd_pileup_set_total_stats(self.h, s))

    def norm_factor(self) -> float:
        f = C.c_double()
        self.n.check(self.n.L.bnd_pileup_norm_factor(self.h, C.byref(f)))
        return f.value

    def to_bedgraph(self, path):
        self.n.check(self.n.L.bnd_pileup_to_bedgraph(self.h, _b(path)))

    def to_bigwig(self, path):
        self.n.check(self.n.L.bnd_pileup_to_bigwig(self.h, _b(path)))

    def accumulate(self):
        self.n.check(self.n.L.bnd_pileup_accumulate(self.h))

    def stage(self):
        self.n.check(self.n.L.bnd_pileup_stage(self.h))

    def accumulate_resident(self):
        self.n.check(self.n.L.bnd_pileup_accumulate_resident(self.h))

    def finalize(self) -> int:
        n = C.c_uint64()
        self.n.check(self.n.L.bnd_pileup_finalize(self.h, C.byref(n)))
        return n.value

    def bedgraph_text(self) -> bytes:
        ptr, n = C.c_void_p(), C.c_uint64()
        self.n.check(self.n.L.bnd_pileup_bedgraph_text(self.h, C.byref(ptr), C.byref(n)))
        data = C.string_at(ptr, n.value)
        self.n.L.bnd_free(ptr)
        return data

    def refs(self):
        """-> (names, lengths) of the BAM's reference dictionary, header order"""
        names, lens, n = C.c_void_p(), C.POINTER(C.c_uint64)(), C.c_uint64()
        self.n.check(self.n.L.bnd_pileup_refs(self.h, C.byref(names), C.byref(lens), C.byref(n)))
        text = C.string_at(names).decode()
        out = (text.split("\n") if n.value else [], [int(lens[i]) for i in range(n.value)])
        self.n.L.bnd_free(names)
        self.n.L.bnd_free(lens)
        return out

    # ---- one BAM per rank (multi-bam-coverage across GPUs) ----
    def multi_prepare(self, resident: bool = False) -> int:
        n = C.c_uint64()
        f = self.n.L.bnd_multi_prepare_resident if resident else self.n.L.bnd_multi_prepare
        self.n.check(f(self.h, C.byref(n)))
        return n.value

    def multi_format_tsv(self, dev_cols_ptr: int, col_stride: int, n_cols: int, row_lo: int, row_hi: int, col_order=None, sizes_only=False) -> list:
        """BAM c's count of row r is dev_cols[col_order[c] * col_stride + (r - row_lo)] (col_order None = identity)."""
        names, _ = self.refs()
        sizes = (C.c_uint64 * max(len(names), 1))()
        order = (C.c_int32 * n_cols)(*col_order) if col_order is not None else None
        self.n.check(self.n.L.bnd_multi_format_tsv(self.h, C.c_void_p(dev_cols_ptr), col_stride, n_cols, order, row_lo, row_hi, int(sizes_only), sizes, len(names)))
        return [int(sizes[i]) for i in range(len(names))]

    def multi_write_tsv_pieces(self, path, header: str, tables: list, rank: int):
        world, n_ref = len(tables), len(tables[0])
        flat = (C.c_uint64 * max(world * n_ref, 1))(*[v for t in tables for v in t])
        self.n.check(self.n.L.bnd_multi_write_tsv_pieces(self.h, _b(path), header.encode(), flat, world, rank))

    def timings(self) -> dict:
        ms = (C.c_double * 8)()
        self.n.check(self.n.L.bnd_pileup_timings(self.h, ms))
        return dict(zip(["inflate_ms", "index_ms", "pileup_ms", "finalize_ms", "h2d_ms", "text_ms", "wall_ms", "read_ms"], [float(v) for v in ms]))

    def shape(self) -> dict:
        s = (C.c_uint64 * 16)()
        self.n.check(self.n.L.bnd_pileup_shape(self.h, s))
        return dict(zip(["records", "comp_bytes", "inflated_bytes", "blocks", "bins", "launches", "chunk_len", "n_chunks", "segments"], [int(v) for v in s]))

    # ---- sharded jobs, one bedGraph ----
    def format_bedgraph(self) -> list:
        """Formats this shard's rows on the device -> text bytes per reference (header order)."""
        names, _ = self.refs()
        sizes = (C.c_uint64 * max(len(names), 1))()
        self.n.check(self.n.L.bnd_pileup_format_bedgraph(self.h, sizes, len(names)))
        return [int(sizes[i]) for i in range(len(names))]

    def preallocate_output(self, path):
        """One rank of a sharded job, before its pile-up: the merged file's estimated pages are allocated in the background."""
        self.n.check(self.n.L.bnd_pileup_preallocate_output(self.h, _b(path)))

    def preallocated_bytes(self) -> int:
        """Stops this rank's background allocation of the merged output (if any) -> leading bytes of the file that have pages."""
        n = C.c_uint64()
        self.n.check(self.n.L.bnd_pileup_preallocated_bytes(self.h, C.byref(n)))
        return n.value

    def write_bedgraph_pieces(self, path, tables: list, rank: int, preallocated: int = 0):
        """tables[k] = format_bedgraph() of rank k; fills this rank's pieces of the one output file.  `preallocated`: leading
        bytes of the file that already have pages (max over ranks of preallocated_bytes())."""
        world, n_ref = len(tables), len(tables[0])
        flat = (C.c_uint64 * max(world * n_ref, 1))(*[v for t in tables for v in t])
        self.n.check(self.n.L.bnd_pileup_write_bedgraph_pieces(self.h, _b(path), flat, world, rank, preallocated))


_native = None


def lib() -> Native:
    """The product library (in-tree sm_100a build).  Raises if it has not been built: there is no fallback."""
    global _native
    if _native is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -m bamnado_b200.build` (nvcc, sm_100a). "
                               "bamnado_b200 is CUDA-only and has no CPU fallback.")
        _native = Native(C.CDLL(str(LIB_PATH)))
    return _native
