#!/usr/bin/env python
"""bench.py — throughput of the B200 coverage path on BASELINE.json's configurations.

Default workload (config.workload): configs[1] — synthetic coordinate-sorted paired-end hg38 BAM (tools/synth_bam.cpp,
seed 20260219, zlib level 6), `bam-coverage --bin-size 50 --normalize rpkm --min-mapq 30 --proper-pairs`, bedGraph output.
`--config 3|4|5` selects the other BASELINE configurations (ATAC fragments + blacklist at bin 10; barcode allowlist +
duplicates + CPM to BigWig; multi-bam-coverage at bin 1 + collapse-bedgraph) with the same line schema.
A "step" is one whole coverage job over that input.

  value   reads/s with the compressed BAM already resident in HBM: K1 inflate -> K2 record index -> K3/K4 filter +
          scatter -> K5 scan/collapse into rows (host clock around blocking C-ABI calls, max over ranks)
  e2e     the same job through the C ABI with HOST buffers: BAM file (page cache) -> pinned -> H2D -> kernels ->
          output text/sections on device -> D2H -> ONE output file closed (N > 1: every rank writes its pieces of the
          same file at offsets agreed by an all-gather of per-contig byte counts)
  roofline  dominant kernel (K1 inflate): algorithmic bytes = compressed bytes read + inflated bytes written per launch,
          over the launch duration measured with CUDA events on the launching stream in a serialised pass
  cpu_baseline  the CPU oracle (port of the reference algorithm) on a bounded sample of the same BAM, all host cores
  parity  the GPU job's output rows over the sample chromosomes against the oracle rows the CPU leg computed, in this run,
          at this size: coordinates and raw counts exact, normalised scores within 1e-6 relative, filter counters exact

N > 1 (torchrun): the genomic chunks are sharded across ranks by BAI byte ranges (disjoint bins, no data-path
collective); the 13 filter counters are summed with one NCCL all-reduce before normalisation. Total work is fixed,
so scaling is "strong".
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SEED = 20260219
CLI_DEFAULTS = dict(min_mapq=20, min_length=20, max_length=1000)  # src/cli.rs FilterOptions defaults
SAMPLE_CHROMS = ["chr1", "chr2", "chr3"]
REL_TOL = 1e-6  # north_star: CPM/RPKM values within 1e-6 relative


def workdir() -> Path:
    base = Path(os.environ.get("BND_BENCH_DIR", "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp")) / "bnd_bench"
    base.mkdir(parents=True, exist_ok=True)
    return base


def outdir() -> Path:
    """Output directory (BND_BENCH_OUT_DIR; default: the tmpfs work dir, the same for both arms — measured faster than the
    VM's ext4 disk for the mmap writer on the B200 hosts)."""
    base = Path(os.environ["BND_BENCH_OUT_DIR"]) / "bnd_bench_out" if os.environ.get("BND_BENCH_OUT_DIR") else workdir()
    base.mkdir(parents=True, exist_ok=True)
    return base


def synth_tool() -> Path:
    out = ROOT / "tools" / "_build" / "synth_bam"
    src = ROOT / "tools" / "synth_bam.cpp"
    if not out.exists() or out.stat().st_mtime < src.stat().st_mtime:
        out.parent.mkdir(parents=True, exist_ok=True)
        subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-o", str(out), str(src), "-lz"], check=True)
    return out


def ensure_bam(tag: str, synth_args: list) -> tuple[Path, dict]:
    bam = workdir() / f"{tag}.bam"
    meta = bam.with_suffix(".json")
    if not (bam.exists() and meta.exists() and Path(str(bam) + ".bai").exists()):
        tmp = bam.with_suffix(".tmp.bam")
        r = subprocess.run([str(synth_tool()), "--out", str(tmp), *map(str, synth_args)], check=True, capture_output=True, text=True)
        os.replace(str(tmp) + ".bai", str(bam) + ".bai")
        os.replace(tmp, bam)
        meta.write_text(r.stdout.strip().splitlines()[-1])
    return bam, json.loads(meta.read_text())


# ---- synthetic side inputs of configs 3 and 4 ---------------------------------------------------------------------
HG38_PRIMARY = [("chr1", 248956422), ("chr2", 242193529), ("chr3", 198295559), ("chr4", 190214555), ("chr5", 181538259), ("chr6", 170805979),
                ("chr7", 159345973), ("chr8", 145138636), ("chr9", 138394717), ("chr10", 133797422), ("chr11", 135086622), ("chr12", 133275309),
                ("chr13", 114364328), ("chr14", 107043718), ("chr15", 101991189), ("chr16", 90338345), ("chr17", 83257441), ("chr18", 80373285),
                ("chr19", 58617616), ("chr20", 64444167), ("chr21", 46709983), ("chr22", 50818468), ("chrX", 156040895), ("chrY", 57227415)]


def make_blacklist(path: Path):
    """1 000 BED intervals, 4 columns, 200-20 000 bp, seeded; every tenth overlaps its predecessor (merge_overlaps has work)."""
    import numpy as np

    rng = np.random.default_rng(SEED)
    lines, prev = [], None
    for i in range(1000):
        name, ln = HG38_PRIMARY[int(rng.integers(0, len(HG38_PRIMARY)))]
        w = int(rng.integers(200, 20001))
        if i % 10 == 9 and prev is not None:
            name, s = prev[0], prev[1] + int(rng.integers(0, max(1, prev[2] - prev[1])))
        else:
            s = int(rng.integers(0, ln - w))
        lines.append(f"{name}\t{s}\t{s + w}\tbl{i}\n")
        prev = (name, s, s + w)
    path.write_text("".join(lines))


def _splitmix(x):
    m = (1 << 64) - 1
    x = (x + 0x9E3779B97F4A7C15) & m
    z = x
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & m
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & m
    return z ^ (z >> 31)


def barcode_text(seed: int, index: int) -> str:
    """The 16-mer tools/synth_bam.cpp writes for barcode `index` (mix(seed ^ 0x6263, index), 2 bits per base)."""
    m = (1 << 64) - 1
    a, b = seed ^ 0x6263, index
    h = _splitmix(a ^ ((b + 0x9E3779B97F4A7C15 + ((a << 6) & m) + (a >> 2)) & m))
    return "".join("ACGT"[(h >> (2 * k)) & 3] for k in range(16)) + "-1"


def make_allowlist(path: Path, n_barcodes=50000):
    """CSV with header `barcode`: 10 000 barcodes, 8 000 of them among the 50 000 in the data (seeded sample), 2 000 absent."""
    import numpy as np

    rng = np.random.default_rng(SEED)
    present = rng.choice(n_barcodes, 8000, replace=False)
    rows = [barcode_text(SEED, int(i)) for i in present] + [barcode_text(SEED, n_barcodes + 7 + i) for i in range(2000)]
    rng.shuffle(rows)
    path.write_text("barcode\n" + "".join(r + "\n" for r in rows))


class Workload:
    """One BASELINE configuration: input BAM(s), pile-up and filter settings, output kind."""

    def __init__(self, args):
        c, pairs = args.config, args.pairs
        self.config, self.pairs, self.out_kind = c, pairs, "bedgraph"
        self.multi = False
        if c == 2:
            self.tag, self.synth = f"cfg2_pe_{pairs}_seed{SEED}", ["--pairs", pairs, "--seed", SEED]
            self.pile = dict(bin_size=50, norm="rpkm", collapse=True)
            self.filt = dict(CLI_DEFAULTS, min_mapq=30, proper_pair=True)
            self.flags = "bam-coverage --bin-size 50 --normalize rpkm --min-mapq 30 --proper-pairs, bedGraph"
            self.name = "configs[1]"
        elif c == 3:
            self.tag, self.synth = f"cfg3_atac_{pairs}_seed{SEED}", ["--pairs", pairs, "--seed", SEED, "--mode", "atac"]
            bed = workdir() / "cfg3_blacklist.bed"
            if not bed.exists():
                make_blacklist(bed)
            self.pile = dict(bin_size=10, norm="raw", collapse=True, use_fragment=True)
            self.filt = dict(CLI_DEFAULTS, max_fragment_length=120, blacklist_bed=str(bed), blacklist_merge_overlaps=True)
            self.flags = "bam-coverage --fragment-counts --max-fragment-len 120 --blacklist bl.bed (1 000 intervals) --bin-size 10, bedGraph"
            self.name = "configs[2]"
        elif c == 4:
            self.tag, self.synth = f"cfg4_cb_{pairs}_seed{SEED}", ["--pairs", pairs, "--seed", SEED, "--barcodes"]
            csv = workdir() / "cfg4_allowlist.csv"
            if not csv.exists():
                make_allowlist(csv)
            self.pile = dict(bin_size=50, norm="cpm", collapse=True)
            self.filt = dict(CLI_DEFAULTS, barcode_csv=str(csv), exclude_duplicates=True)
            self.flags = "bam-coverage --barcode-allowlist bc.csv (10 000 barcodes) --ignore-duplicates --normalize cpm, BigWig"
            self.out_kind, self.name = "bigwig", "configs[3]"
        else:
            raise SystemExit(f"--config {c}: use 2, 3, 4 (single BAM) or 5 (bench_multi.py drives multi-bam-coverage)")

    def ensure(self):
        self.bam, self.meta = ensure_bam(self.tag, self.synth)
        return self

    def describe(self, args):
        m = self.meta
        return {"workload": f"{self.name}: synthetic {self.pairs}-pair hg38 BAM (seed {SEED}, zlib-6, {m['records']} records, {m['file_bytes']} B "
                            f"compressed, {m['inflated_bytes']} B inflated), {self.flags}",
                "pairs": self.pairs, "bin_size": self.pile["bin_size"], "parallelism": f"chunk-sharded x{args.gpus}" if args.gpus > 1 else "single GPU",
                "l2": "inputs (compressed + inflated stream) are far larger than the 126 MB L2; no explicit flush"}


class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons during the timed region."""

    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx = float(r[1])
                for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        busy = [x for x in sm if mx and x > 0.5 * mx] or sm
        return {"sm_mhz": statistics.median(busy) if busy else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def oracle_sample(W: Workload, keep: bool = False, threads: int = 0):
    """CPU oracle on a bounded sample (a few chromosomes of the same BAM, pileup + bedGraph written)
    -> (reads, seconds, stats, factor, path).  The oracle has no BigWig writer: config 4's CPU leg writes the same rows as text."""
    from oracle import oracle_py as o

    pile = dict(W.pile, bam_path=str(W.bam), n_threads=threads)
    out = outdir() / f"oracle_sample_cfg{W.config}.bedgraph"
    t0 = time.perf_counter()
    st, factor = o.pileup_chroms_to_bedgraph(pile, W.filt, SAMPLE_CHROMS, out)  # open BAM -> bedGraph of the sample closed
    dt = time.perf_counter() - t0
    if not keep:
        out.unlink(missing_ok=True)
    return st["n_total"], dt, st, factor, out


def run_reference(args, W: Workload):
    """--impl reference: the reference's CPU algorithm (the C++ oracle port; the Rust crate cannot be built here:
    no cargo/rustc in the image) on the host cores, each step a bounded sample of the same workload."""
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        oracle_sample(W)
    n_tot, t_tot = 0, 0.0
    for _ in range(args.steps):
        n, t, *_ = oracle_sample(W)
        n_tot += n
        t_tot += t
    v = n_tot / t_tot
    sample = f"bam-coverage (pileup + bedGraph written) over {'+'.join(SAMPLE_CHROMS)} of the same BAM per step ({n_tot // max(args.steps, 1)} reads/step)"
    emit({
        "impl": "reference", "metric": "bam-coverage reads/sec", "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(args.steps, 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic", "config": W.describe(args),
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "compressed_GBps": v * W.meta["file_bytes"] / W.meta["records"] / 1e9,
    })


def measure_h2d_gbps(torch, n_bytes=1 << 30, reps=3):
    """Pinned host -> device copy bandwidth on this box (the roofline that binds the end-to-end path;
    MEASURED_PEAKS.json has no H2D entry).  Best of `reps` copies of `n_bytes`, CUDA events on the current stream."""
    try:
        host = torch.empty(n_bytes, dtype=torch.uint8, pin_memory=True)
        dev = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
        best = 0.0
        for _ in range(reps + 1):  # first copy warms up
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            dev.copy_(host, non_blocking=True)
            e1.record()
            e1.synchronize()
            best = max(best, n_bytes / (e0.elapsed_time(e1) * 1e-3) / 1e9)
        del host, dev
        return best
    except Exception:
        return None


# ---- parity at the benchmarked size ----------------------------------------------------------------------------------------------
def _bedgraph_blocks(path: Path, chroms):
    """Rows of the given chromosomes out of a bedGraph sorted by (chrom byte string, start) -> {chrom: (start, end, score) arrays}."""
    import numpy as np
    import pandas as pd

    data = np.memmap(path, dtype=np.uint8, mode="r") if path.stat().st_size else np.zeros(0, np.uint8)
    raw = data.tobytes() if len(data) < (1 << 20) else None
    out = {}
    import mmap

    with open(path, "rb") as f:
        mm = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ) if path.stat().st_size else None
        for c in chroms:
            key = c.encode() + b"\t"
            if mm is None:
                out[c] = None
                continue
            first = 0 if mm[: len(key)] == key else mm.find(b"\n" + key)
            if first < 0:
                out[c] = None
                continue
            if first > 0 or mm[: len(key)] != key:
                first += 1
            last = mm.rfind(b"\n" + key)
            last = first if last < first else last + 1
            end = mm.find(b"\n", last) + 1
            import io

            df = pd.read_csv(io.BytesIO(mm[first:end]), sep="\t", header=None, names=["c", "s", "e", "v"], dtype={"c": str, "s": np.int64, "e": np.int64, "v": np.float64},
                             float_precision="round_trip")
            assert (df["c"] == c).all()
            out[c] = (df["s"].to_numpy(), df["e"].to_numpy(), df["v"].to_numpy())
        if mm is not None:
            mm.close()
    del raw
    return out


def _bigwig_intervals(path: Path):
    """Fast reader for the parity check: -> ({name: id}, chrom_id[], start[], end[], value[]) of a bedGraph-section BigWig."""
    import struct
    import zlib

    import numpy as np

    d = open(path, "rb").read()
    magic, version, zooms, chrom_off, data_off, index_off, fc, dfc, sql_off, summ_off, ubuf, _ = struct.unpack_from("<IHHQQQHHQQIQ", d, 0)
    assert magic == 0x888FFC26
    chroms = {}

    def walk_chrom(off, key_size):
        is_leaf, _, count = struct.unpack_from("<BBH", d, off)
        off += 4
        for _ in range(count):
            key = d[off:off + key_size].rstrip(b"\0").decode()
            if is_leaf:
                chroms[key] = struct.unpack_from("<I", d, off + key_size)[0]
            else:
                walk_chrom(struct.unpack_from("<Q", d, off + key_size)[0], key_size)
            off += key_size + 8

    _, _, key_size, _, n_items, _ = struct.unpack_from("<IIIIQQ", d, chrom_off)
    if n_items:
        walk_chrom(chrom_off + 32, key_size)
    leaves = []

    def walk_r(off):
        is_leaf, _, count = struct.unpack_from("<BBH", d, off)
        off += 4
        if is_leaf:
            a = np.frombuffer(d, dtype="<u8", count=count * 4, offset=off).reshape(count, 4)  # 32 B per item: 4 x u32, offset, size
            leaves.extend(zip(a[:, 2].tolist(), a[:, 3].tolist()))
        else:
            for k in range(count):
                walk_r(struct.unpack_from("<Q", d, off + 24 * k + 16)[0])

    _, _, ritems = struct.unpack_from("<IIQ", d, index_off)
    if ritems:
        walk_r(index_off + 48)
    cid, st, en, val = [], [], [], []
    item = np.dtype([("s", "<u4"), ("e", "<u4"), ("v", "<f4")])
    for doff, dsize in sorted(leaves):
        sec = d[doff:doff + dsize]
        if ubuf:
            sec = zlib.decompress(sec)
        c, _, _, _, _, typ, _, count = struct.unpack_from("<IIIIIBBH", sec, 0)
        assert typ == 1
        a = np.frombuffer(sec, dtype=item, count=count, offset=24)
        cid.append(np.full(count, c, np.uint32))
        st.append(a["s"])
        en.append(a["e"])
        val.append(a["v"])
    cat = lambda xs, dt: np.concatenate(xs) if xs else np.zeros(0, dt)  # noqa: E731
    return chroms, cat(cid, np.uint32), cat(st, np.uint32), cat(en, np.uint32), cat(val, np.float32)


def parity_check(N, W: Workload, pile: dict, gpu_out: Path, gpu_factor: float, ost: dict, ofactor: float, opath: Path) -> dict:
    """GPU output rows of the sample chromosomes vs the oracle's rows for the same chromosomes (computed by the CPU leg of this
    run), plus the 13 filter counters of those chromosomes.  Raw counts are recovered from the scores of each arm with that
    arm's own factor (the oracle normalises by the sample's reads, the job by all reads)."""
    import numpy as np

    res = {"checked": True, "chroms": SAMPLE_CHROMS, "rows": 0, "mismatches": 0, "max_rel_err": 0.0, "counters_equal": None, "what": W.out_kind}
    ob = _bedgraph_blocks(opath, SAMPLE_CHROMS)
    if W.out_kind == "bedgraph":
        gb = _bedgraph_blocks(gpu_out, SAMPLE_CHROMS)
    else:
        names, cid, st, en, val = _bigwig_intervals(gpu_out)
        gb = {}
        for c in SAMPLE_CHROMS:
            m = cid == names[c]
            gb[c] = (st[m].astype(np.int64) + 1, en[m].astype(np.int64) + 1, val[m].astype(np.float64))  # BigWig rows are shifted by -1
    for c in SAMPLE_CHROMS:
        o, g = ob[c], gb[c]
        if o is None or g is None or len(o[0]) != len(g[0]):
            res["mismatches"] += 1 + (abs(len(o[0]) - len(g[0])) if o is not None and g is not None else 0)
            continue
        res["rows"] += len(o[0])
        bad = (o[0] != g[0]) | (o[1] != g[1])
        ocount = np.rint(o[2] / ofactor) if ofactor else np.zeros(len(o[2]))
        if W.out_kind == "bedgraph":
            gcount = np.rint(g[2] / gpu_factor) if gpu_factor else np.zeros(len(g[2]))
            expect = ocount * gpu_factor
        else:
            expect = (ocount * gpu_factor).astype(np.float32).astype(np.float64)  # to_bigwig casts the f64 product to f32
            gcount = ocount.copy()
            gcount[g[2] != expect] = -1
        bad |= ocount != gcount
        rel = np.abs(g[2] - expect) / np.maximum(np.abs(expect), 1e-300)
        rel[expect == 0] = np.abs(g[2][expect == 0])
        res["max_rel_err"] = max(res["max_rel_err"], float(rel.max()) if len(rel) else 0.0)
        bad |= rel > REL_TOL
        res["mismatches"] += int(bad.sum())
    # filter counters of the same chromosomes: one single-contig pass each (pileup_chromosome), summed
    tot = None
    with N.pileup(dict(pile, shard_rank=0, shard_world=1), W.filt) as h:
        for c in SAMPLE_CHROMS:
            h.chromosome(c)
            s = h.stats()
            tot = s if tot is None else {k: tot[k] + s[k] for k in s}
    res["counters_equal"] = tot == ost
    if not res["counters_equal"]:
        res["mismatches"] += 1
        res["counters"] = {"gpu": tot, "oracle": ost}
    return res


_REAL_STDOUT = None


def claim_stdout():
    """stdout must carry exactly one JSON line, but native libraries write there too (NCCL prints its version banner on
    the first collective): file descriptor 1 is pointed at stderr for the whole run and the result line goes to a saved
    duplicate of the original stdout."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line: dict):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5])
    ap.add_argument("--pairs", type=int, default=int(os.environ.get("BND_BENCH_PAIRS", "100000000")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-parity", action="store_true")
    args = ap.parse_args()
    if args.config == 5:
        import bench_multi

        return bench_multi.main(args)
    claim_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    W = Workload(args)

    if args.impl == "reference":
        if rank == 0:
            run_reference(args, W.ensure())
        return

    import torch
    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: bamnado_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if rank == 0:
        W.ensure()
    barrier()
    W.ensure()
    bam, meta = W.bam, W.meta

    from bamnado_b200 import _native

    N = _native.lib()
    pile = dict(W.pile, bam_path=str(bam), device=local_rank, shard_rank=rank, shard_world=world)
    ext = {"bedgraph": "bedgraph", "bigwig": "bw"}[W.out_kind]
    out_paths = []

    def allreduce_stats(h):
        st = h.stats()
        if world > 1:
            t = torch.tensor(list(st.values()), dtype=torch.int64, device="cuda")
            dist.all_reduce(t)  # 13 x u64 over NCCL: the only exchange on this path (CPM/RPKM denominator)
            h.set_total_stats([int(x) for x in t.tolist()])

    # ---- e2e leg: host file -> ONE output file through the C ABI ----
    trace = bool(os.environ.get("BND_BENCH_TRACE"))

    def e2e_step():
        out_path = outdir() / f"out_cfg{W.config}_{len(out_paths)}.{ext}"  # a fresh output file per job, shared by all ranks
        out_paths.append(out_path)
        stamps = [("start", time.perf_counter())]
        mark = lambda name: stamps.append((name, time.perf_counter()))  # noqa: E731
        with N.pileup(pile, W.filt) as h:
            mark("create")
            if world > 1:  # sharded: the ranks meet once to sum the 13 filter counters before normalising, once to place their pieces
                if W.out_kind == "bedgraph" and rank == 0:
                    h.preallocate_output(out_path)  # one rank allocates the merged file's pages behind its pile-up
                h.accumulate()
                mark("accumulate")
                allreduce_stats(h)
                mark("allreduce")
                h.finalize()
                mark("finalize")
                if W.out_kind == "bedgraph":
                    sizes = h.format_bedgraph()  # text on the device; bytes per reference of this shard
                    mark("format")
                    t = torch.tensor(sizes + [h.preallocated_bytes()], dtype=torch.int64, device="cuda")
                    allv = [torch.empty_like(t) for _ in range(world)]
                    dist.all_gather(allv, t)  # n_ref x u64 per rank: where every rank's pieces go in the one sorted file (+ the bytes already allocated)
                    tables = [[int(x) for x in v.tolist()] for v in allv]
                    mark("all_gather")
                    h.write_bedgraph_pieces(out_path, [tb[:-1] for tb in tables], rank, max(tb[-1] for tb in tables))
                    mark("write_pieces")
                    dist.barrier()
                    mark("barrier")
                else:
                    h.to_bigwig(str(out_path) + f".rank{rank}")  # BigWig sections of a shard: one file per rank
                    mark("to_bigwig")
            elif W.out_kind == "bedgraph":
                h.to_bedgraph(out_path)  # single GPU: one call = BamPileup::to_bedgraph (pile-up, normalise, write)
                mark("to_bedgraph")
            else:
                h.to_bigwig(out_path)
                mark("to_bigwig")
            res = h.shape(), h.timings(), h.norm_factor()
        mark("close")
        if trace:
            sys.stderr.write(f"[trace rank {rank}] " + ", ".join(f"{b[0]} {1e3 * (b[1] - a[1]):.1f}" for a, b in zip(stamps, stamps[1:])) + " ms\n")
        return res

    def clean_outputs(keep_last=False):
        for p in out_paths[:-1] if keep_last else out_paths:
            if rank == 0:
                p.unlink(missing_ok=True)
            Path(str(p) + f".rank{rank}").unlink(missing_ok=True)

    for _ in range(args.warmup):
        e2e_step()
    barrier()
    clean_outputs()
    out_paths.clear()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        shape, tm_e2e, gpu_factor = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    last_out = out_paths[-1]
    if world > 1 and W.out_kind == "bigwig":
        out_bytes = Path(str(last_out) + f".rank{rank}").stat().st_size
    else:
        out_bytes = last_out.stat().st_size if rank == 0 else 0
    barrier()
    clean_outputs(keep_last=True)

    # ---- resident leg: compressed bytes already in HBM ----
    h = N.pileup(pile, W.filt)
    h.stage()

    def step():
        h.accumulate_resident()
        allreduce_stats(h)
        return h.finalize()

    for _ in range(args.warmup):
        step()
    sampler = ClockSampler(local_rank)
    launches0 = N.launch_count()
    barrier()
    if rank == 0:
        sampler.start()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        rows = step()
    barrier()
    res_s = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    launches = N.launch_count() - launches0
    tm_res = h.timings()
    rows_local = rows
    if world > 1:
        t = torch.tensor([res_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res_s = float(t.item())
        t = torch.tensor([launches, shape["comp_bytes"], out_bytes, rows], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        launches, h2d_total, out_bytes, rows = [int(x) for x in t.tolist()]
    else:
        h2d_total = shape["comp_bytes"]

    # ---- roofline leg (rank 0): K1 launch durations without cross-stream overlap ----
    roofline = None
    if rank == 0:
        os.environ["BND_SERIALIZE"] = "1"
        h.accumulate_resident()
        ts, sh = h.timings(), h.shape()
        os.environ["BND_SERIALIZE"] = "0"
        n_launch = max(1, sh["segments"])
        alg_bytes = sh["comp_bytes"] + sh["inflated_bytes"]  # per pass; DESIGN.md: C + U bytes per record
        peak, peak_src = 6451.5, "fallback of B200_PROFILING.md"
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
            peak, peak_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
        except Exception:
            peak = 6650.0
        achieved = alg_bytes / (ts["inflate_ms"] * 1e-3) / 1e9
        step_ms = 1e3 * res_s / args.steps
        roofline = {"kernel": "bgzf_inflate_spec_kernel (K1)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None, "peak_source": peak_src, "launches_timed": n_launch,
                    "avg_launch_ms": ts["inflate_ms"] / n_launch, "algorithmic_bytes_per_launch": alg_bytes / n_launch,
                    "inflated_GBps": sh["inflated_bytes"] / (ts["inflate_ms"] * 1e-3) / 1e9,
                    "serialised_pass_k1_ms": ts["inflate_ms"], "pipelined_step_ms": step_ms,
                    "in_pipeline": {"inflated_GBps_lower_bound": sh["inflated_bytes"] / (step_ms * 1e-3) / 1e9,
                                    "note": "launch durations come from a pass with one segment at a time (every launch's tail is exposed); in the overlapped "
                                            "pipeline the next launch fills that tail, so the step can be shorter than the sum of the serialised launches"},
                    "other_kernels_ms": {"record_index": ts["index_ms"], "filter_scatter": ts["pileup_ms"], "scan_collapse": tm_res["finalize_ms"]},
                    "note": "K1 decodes one BGZF block per warp (32 self-synchronising spans side by side): bound by instruction issue and dependency latency, reported against HBM for honesty (SURVEY 8d)"}
        # the other kernels against the same HBM peak, algorithmic bytes of SURVEY 8(d): K2+K3 read the inflated records
        # and write 8 B of offset each, K4 adds 8 B of atomics per record, K5 moves 12 B per bin + 16 B per output row
        recs, bins_n = sh["records"], sh["bins"]
        k234_bytes = sh["inflated_bytes"] + 16 * recs
        k5_bytes = 12 * bins_n + 16 * rows_local
        k234_ms, k5_ms = ts["index_ms"] + ts["pileup_ms"], tm_res["finalize_ms"]
        roofline["other_kernels"] = [
            {"kernel": "K2 record index + K3/K4 filter, intervals, scatter", "bound": "hbm", "algorithmic_bytes": k234_bytes, "ms": k234_ms,
             "achieved": k234_bytes / max(k234_ms, 1e-9) / 1e6, "unit": "GB/s", "frac": k234_bytes / max(k234_ms, 1e-9) / 1e6 / peak},
            {"kernel": "K5 scan + run collapse", "bound": "hbm", "algorithmic_bytes": k5_bytes, "ms": k5_ms,
             "achieved": k5_bytes / max(k5_ms, 1e-9) / 1e6, "unit": "GB/s", "frac": k5_bytes / max(k5_ms, 1e-9) / 1e6 / peak}]
        traffic_file = ROOT / "profiles" / "traffic.json"
        if traffic_file.exists():
            try:
                tj = json.loads(traffic_file.read_text())
                roofline["traffic"] = tj.get("inflate_dram_bytes_per_launch")
                roofline["traffic_source"] = "static: " + tj.get("source", "profiles/traffic.json") + " (ncu --set full capture, not measured in this run)"
            except Exception:
                pass
    h.close()

    cpu, parity = None, None
    if rank == 0 and not args.no_cpu_baseline:
        n, t, ost, ofactor, opath = oracle_sample(W, keep=True)
        cpu = {"value": n / t, "unit": "reads/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"CPU oracle (C++ port of the reference algorithm; Rust toolchain absent), pileup + bedGraph written, over {'+'.join(SAMPLE_CHROMS)} of the same BAM: {n} reads in {t:.2f} s"}
        if not args.no_parity:
            gpu_out = last_out
            if world > 1 and W.out_kind == "bigwig":  # the check needs one whole file: an unsharded job outside the timed region
                gpu_out = outdir() / f"parity_cfg{W.config}.{ext}"
                with N.pileup(dict(pile, shard_rank=0, shard_world=1), W.filt) as hp:
                    hp.to_bigwig(gpu_out)
            parity = parity_check(N, W, pile, gpu_out, gpu_factor, ost, ofactor, opath)
            if gpu_out != last_out:
                gpu_out.unlink(missing_ok=True)
        opath.unlink(missing_ok=True)
    barrier()
    clean_outputs()

    if rank == 0:
        records = meta["records"]
        value = records * args.steps / res_s
        e2e = records * args.steps / e2e_s
        h2d_peak = measure_h2d_gbps(torch)
        line = {
            "metric": "bam-coverage reads/sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * res_s / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": W.describe(args),
            "compressed_GBps": meta["file_bytes"] * args.steps / res_s / 1e9,
            "e2e": {"value": e2e, "unit": "reads/s", "ms_per_step": 1e3 * e2e_s / args.steps, "compressed_GBps": meta["file_bytes"] * args.steps / e2e_s / 1e9,
                    "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": out_bytes, "stages_ms_last_step": tm_e2e,
                    "output": f"one {W.out_kind} file" if not (world > 1 and W.out_kind == "bigwig") else "one BigWig per rank",
                    # the end-to-end path against the measured pinned H2D bandwidth of this box
                    "h2d_peak_GBps": h2d_peak,
                    "frac_of_h2d_roofline": (meta["file_bytes"] * args.steps / e2e_s / 1e9 / h2d_peak / world) if h2d_peak else None},
            "gpu_launches": launches, "rows": rows, "host_cores": os.cpu_count(), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "parity": parity, "stages_ms_resident_last_step": tm_res,
        }
        emit(line)
        if parity is not None and parity["mismatches"]:
            sys.stderr.write(f"PARITY FAILURE: {json.dumps(parity)}\n")
            if world > 1:
                dist.destroy_process_group()
            sys.exit(3)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
