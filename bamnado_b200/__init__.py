"""bamnado_b200 — B200-native (sm_100a) implementation of BamNado's coverage hot path.

Mirrors the reference's Python front door for this path (bamnado-python/python/bamnado/__init__.py):
`ReadFilter`, `get_signal_for_chromosome`, `__version__`, plus `cli.main`.  CUDA-only: no CPU fallback.
"""
from __future__ import annotations

import numpy as np

from . import _native

__all__ = ["ReadFilter", "get_signal_for_chromosome", "__version__", "count_bins", "BinCounts"]


def __version__() -> str:
    """Version of the reference surface this package mirrors (bamnado 0.9.0) plus the backend tag."""
    return "0.9.0+b200"


class ReadFilter:
    """Configuration for filtering BAM reads (bamnado-python/src/lib.rs:48-163: same fields and defaults)."""

    def __init__(self, min_mapq: int = 0, proper_pair: bool = False, min_length: int = 0, max_length: int = 1000,
                 strand: str = "both", min_fragment_length: int | None = None, max_fragment_length: int | None = None,
                 blacklist_bed: str | None = None, whitelisted_barcodes: list[str] | None = None,
                 read_group: str | None = None, filter_tag: str | None = None, filter_tag_value: str | None = None) -> None:
        if strand not in ("both", "forward", "reverse"):
            raise ValueError(f"Invalid strand: {strand}")  # lib.rs:109-116
        self.min_mapq = min_mapq
        self.proper_pair = proper_pair
        self.min_length = min_length
        self.max_length = max_length
        self.strand = strand
        self.min_fragment_length = min_fragment_length
        self.max_fragment_length = max_fragment_length
        self.blacklist_bed = blacklist_bed
        self.whitelisted_barcodes = whitelisted_barcodes
        self.read_group = read_group
        self.filter_tag = filter_tag
        self.filter_tag_value = filter_tag_value

    def _cfg(self) -> dict:
        # build_bam_read_filter (lib.rs:167-219): duplicates / secondary / supplementary exclusion is not exposed
        return dict(strand=self.strand, proper_pair=self.proper_pair, min_mapq=self.min_mapq, min_length=self.min_length,
                    max_length=self.max_length, min_fragment_length=self.min_fragment_length,
                    max_fragment_length=self.max_fragment_length, blacklist_bed=self.blacklist_bed,
                    blacklist_merge_overlaps=False, barcodes=self.whitelisted_barcodes, read_group=self.read_group,
                    filter_tag=self.filter_tag, filter_tag_value=self.filter_tag_value)

    def __repr__(self) -> str:
        return (f"ReadFilter(min_mapq={self.min_mapq}, proper_pair={self.proper_pair}, min_length={self.min_length}, "
                f"max_length={self.max_length}, strand='{self.strand}')")


# BamReadFilter::default() (read_filter.rs:403-420): what `read_filter=None` means
_DEFAULT_FILTER = dict(strand="both", proper_pair=True, min_mapq=0, min_length=0, max_length=1000)


def get_signal_for_chromosome(bam_path: str, chromosome_name: str, bin_size: int, scale_factor: float, use_fragment: bool,
                              ignore_scaffold_chromosomes: bool, read_filter: ReadFilter | None = None) -> np.ndarray:
    """Dense float32 signal of one chromosome (bamnado-python/src/lib.rs:248-304).

    Returns an array of length chrom_size; unknown chromosome -> KeyError; other failures -> RuntimeError.
    """
    filt = _DEFAULT_FILTER if read_filter is None else read_filter._cfg()
    pile = dict(bam_path=bam_path, bin_size=bin_size, norm="raw", scale_factor=scale_factor, use_fragment=use_fragment,
                collapse=True, ignore_scaffold=ignore_scaffold_chromosomes)
    return _native.lib().signal_for_chromosome(pile, filt, chromosome_name)


class BinCounts:
    """Sample x bin read-count matrix (`BinCounts`, bamnado/src/bin_counts.rs:36-48).

    `counts` has shape (n_bins, n_samples), dtype uint64; `bins` yields (chrom, start, end) with 1-based inclusive
    coordinates in row order; `library_sizes` are the post-filter library sizes per sample.
    """

    def __init__(self, sample_names, bin_size, raw):
        self.sample_names = list(sample_names)
        self.bin_size = int(bin_size)
        self.counts = raw["counts"]
        self.library_sizes = raw["library_sizes"]
        self.chroms = raw["chroms"]
        self.chrom_len = raw["chrom_len"]
        self.chrom_first_row = raw["chrom_first_row"]

    @property
    def n_bins(self) -> int:
        return int(self.counts.shape[0])

    def bin_arrays(self):
        """(chrom index per row, start, end) as numpy arrays."""
        rows = np.diff(self.chrom_first_row).astype(np.int64)
        chrom_idx = np.repeat(np.arange(len(self.chroms)), rows)
        i = np.arange(self.n_bins, dtype=np.uint64) - np.repeat(self.chrom_first_row[:-1], rows)
        start = 1 + i * np.uint64(self.bin_size)
        end = np.minimum(start + np.uint64(self.bin_size - 1), np.repeat(self.chrom_len, rows))
        return chrom_idx, start, end

    @property
    def bins(self):
        chrom_idx, start, end = self.bin_arrays()
        for c, s, e in zip(chrom_idx, start, end):
            yield self.chroms[int(c)], int(s), int(e)


def count_bins(bam_paths, sample_names, bin_size: int, filters=None, use_fragment: bool = False,
               ignore_scaffold_chromosomes: bool = False) -> BinCounts:
    """Bin-count reads (or fragments) of several BAMs on one shared grid (`count_bins`, bamnado/src/bin_counts.rs:100-289).

    Same argument checks as the reference: at least one BAM, as many sample names and filters as BAMs, identical
    chromosome name -> length maps (RuntimeError otherwise).
    """
    bam_paths = [str(p) for p in bam_paths]
    if not bam_paths:
        raise RuntimeError("At least one BAM file is required for bin counting")
    if len(bam_paths) != len(sample_names):
        raise RuntimeError(f"Number of BAM paths ({len(bam_paths)}) does not match number of sample names ({len(sample_names)})")
    if filters is None:
        filters = [None] * len(bam_paths)
    if len(bam_paths) != len(filters):
        raise RuntimeError(f"Number of BAM paths ({len(bam_paths)}) does not match number of filters ({len(filters)})")
    piles = [dict(bam_path=p, bin_size=bin_size, norm="raw", use_fragment=use_fragment, collapse=False,
                  ignore_scaffold=ignore_scaffold_chromosomes) for p in bam_paths]
    filts = [dict(_DEFAULT_FILTER) if f is None else (f._cfg() if isinstance(f, ReadFilter) else dict(f)) for f in filters]
    return BinCounts(sample_names, bin_size, _native.lib().bin_counts(piles, filts))
