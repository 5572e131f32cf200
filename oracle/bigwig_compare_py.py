"""CPU restatement of bamnado/src/bigwig_compare.rs for the parity tests (test infrastructure, never the product).

Pure Python on (start, end, value) interval lists: the same sweeps, the same order of f64 additions, f32 casts where the
reference casts.  Small inputs only.  Citations are into /root/reference/bamnado/src/bigwig_compare.rs.
"""
from __future__ import annotations

import math

import numpy as np

f32 = np.float32


def n_bins(chrom_len: int, bin_size: int) -> int:  # :70-76
    return 0 if bin_size == 0 else -(-chrom_len // bin_size)


class BinAccumulator:  # :83-142
    def __init__(self, chrom_len, bin_size):
        self.chrom_len, self.bin_size = chrom_len, bin_size
        self.sums = [0.0] * n_bins(chrom_len, bin_size)

    def add_segment(self, start, end, value: float):
        if self.bin_size == 0 or end <= start:
            return
        start, end = min(start, self.chrom_len), min(end, self.chrom_len)
        if end <= start:
            return
        pos, b = start, self.bin_size
        while pos < end:
            bi = pos // b
            overlap_end = min(end, min((bi + 1) * b, self.chrom_len))
            self.sums[bi] += value * float(overlap_end - pos)
            pos = overlap_end

    def into_means(self):
        out = []
        for i, s in enumerate(self.sums):
            start = i * self.bin_size
            if start >= self.chrom_len:
                break
            ln = float(min(start + self.bin_size, self.chrom_len) - start)
            out.append(f32(s / ln) if ln > 0 else f32(0.0))
        return out


class Cursor:  # StepSignalCursor :152-178
    def __init__(self, ivs):
        self.ivs, self.idx = ivs, 0

    def fetch(self, pos, chrom_len):
        while self.idx < len(self.ivs) and self.ivs[self.idx][1] <= pos:
            self.idx += 1
        if self.idx < len(self.ivs):
            s, e, v = self.ivs[self.idx]
            if s <= pos < e:
                return float(v), min(e, chrom_len)
            return 0.0, min(s, chrom_len)
        return 0.0, chrom_len


def load(ivs):  # load_intervals_for_chrom :46-67: empty intervals dropped, defensive sort
    return sorted([(s, e, f32(v)) for s, e, v in ivs if e > s], key=lambda t: (t[0], t[1]))


def single_means(ivs, chrom_len, bin_size):  # binned_means_single_bw :185-206
    acc = BinAccumulator(chrom_len, bin_size)
    for s, e, v in load(ivs):
        if v == 0.0:
            continue
        acc.add_segment(s, e, float(v))
    return acc.into_means()


def compare(a, b, chrom_len, bin_size, comparison, pseudocount=None, scale1=None, scale2=None):  # binned_compare_two_bw :226-268 with the defaults of :640-650
    pc = 1e-12 if pseudocount is None else pseudocount
    s1f, s2f = 1.0 if scale1 is None else scale1, 1.0 if scale2 is None else scale2
    c1, c2, acc = Cursor(load(a)), Cursor(load(b)), BinAccumulator(chrom_len, bin_size)
    pos = 0
    while pos < chrom_len:
        v1, n1 = c1.fetch(pos, chrom_len)
        v2, n2 = c2.fetch(pos, chrom_len)
        nxt = min(n1, n2, chrom_len)
        if nxt <= pos:
            nxt = min(pos + 1, chrom_len)
        s1, s2 = v1 * s1f, v2 * s2f
        if comparison == "subtraction":
            val = s1 - s2
        elif comparison == "ratio":
            d = s2 + pc
            val = s1 / d if d != 0.0 else (math.nan if s1 == 0.0 else math.copysign(math.inf, s1) * math.copysign(1.0, d))
        else:
            q = (s1 + pc) / (s2 + pc) if (s2 + pc) != 0.0 else math.inf
            val = math.log(q) if q > 0 else (-math.inf if q == 0 else math.nan)
        acc.add_segment(pos, nxt, val)
        pos = nxt
    return acc.into_means()


def aggregate(files, chrom_len, bin_size, method, pseudocount=None, scale_factors=None):  # aggregate_bigwigs_with_threads :724-790
    pc = 0.0 if pseudocount is None else pseudocount
    sfs = [1.0] * len(files) if scale_factors is None else list(scale_factors)
    nb = n_bins(chrom_len, bin_size)
    if method in ("sum", "mean"):  # binned_sum_or_mean :676-698
        acc = [0.0] * nb
        for k, ivs in enumerate(files):
            for i, x in enumerate(single_means(ivs, chrom_len, bin_size)):
                acc[i] += float(x) * sfs[k] + pc
        den = float(len(files)) if method == "mean" else 1.0
        return [f32(s / den) for s in acc]
    if method in ("max", "min"):  # binned_extrema_multi_bw :276-328
        cur = [Cursor(load(ivs)) for ivs in files]
        acc = BinAccumulator(chrom_len, bin_size)
        pos = 0
        while pos < chrom_len:
            nxt, ext = chrom_len, (-math.inf if method == "max" else math.inf)
            for k, c in enumerate(cur):
                v, n = c.fetch(pos, chrom_len)
                nxt = min(nxt, n)
                vv = v * sfs[k] + pc
                ext = max(ext, vv) if method == "max" else min(ext, vv)
            if nxt <= pos:
                nxt = min(pos + 1, chrom_len)
            acc.add_segment(pos, nxt, ext)
            pos = nxt
        return acc.into_means()
    # median :334-470
    scaled = [[(s, e, f32(float(v) * sfs[k] + pc)) for s, e, v in load(ivs)] for k, ivs in enumerate(files)]
    gap = f32(pc)
    out = []
    for bi in range(nb):
        start = bi * bin_size
        if start >= chrom_len:
            break
        end = min(start + bin_size, chrom_len)
        pairs = []
        for ivs in scaled:
            pairs += _collect(ivs, start, end, gap)
        out.append(_weighted_median(pairs))
    return out


def _collect(ivs, bs, be, gap):  # collect_weighted_values_for_bw_in_bin :361-418
    out, pos = [], bs
    i = 0
    while i < len(ivs) and ivs[i][1] <= bs:
        i += 1
    while pos < be:
        if i >= len(ivs):
            out.append((gap, be - pos))
            break
        s, e, v = ivs[i]
        if s > pos:
            g = min(s, be)
            if g > pos:
                out.append((gap, g - pos))
                pos = g
                continue
        if e <= pos:
            i += 1
            continue
        if s <= pos < e:
            se = min(e, be)
            if se > pos:
                out.append((v, se - pos))
                pos = se
                if pos >= e:
                    i += 1
                continue
        if s >= be:
            break
        i += 1
    return out


def _weighted_median(pairs):  # :334-358
    if not pairs:
        return f32(0.0)
    pairs = sorted(pairs, key=lambda t: t[0])
    total = sum(w for _, w in pairs)
    if total == 0:
        return f32(0.0)
    half = -(-total // 2)
    cum = 0
    for v, w in pairs:
        cum += w
        if cum >= half:
            return v
    return f32(0.0)


def merge_bins(bins, chrom_len, bin_size):  # push_bigwig_values_for_bins :520-566 -> [(start, end, value)]
    out, i = [], 0
    while i < len(bins):
        start = i * bin_size
        if start >= chrom_len:
            break
        end, v, j = min(start + bin_size, chrom_len), bins[i], i + 1
        while j < len(bins) and j * bin_size < chrom_len and not (bins[j] != v):
            end = min(j * bin_size + bin_size, chrom_len)
            j += 1
        out.append((start, end, v))
        i = j
    return out
