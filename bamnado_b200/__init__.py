"""bamnado_b200 — B200-native (sm_100a) implementation of BamNado's coverage hot path.

Mirrors the reference's Python front door for this path (bamnado-python/python/bamnado/__init__.py):
`ReadFilter`, `get_signal_for_chromosome`, `__version__`, plus `cli.main`.  CUDA-only: no CPU fallback.
"""
from __future__ import annotations

import numpy as np

from . import _native

__all__ = ["ReadFilter", "get_signal_for_chromosome", "__version__"]


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
