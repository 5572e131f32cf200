#!/usr/bin/env python
"""bench.py — bam-coverage throughput of the B200 path on BASELINE.json's configs[1].

Workload (config.workload): synthetic coordinate-sorted paired-end hg38 BAM (tools/synth_bam.cpp, seed 20260219,
zlib level 6), `bam-coverage --bin-size 50 --normalize rpkm --min-mapq 30 --proper-pairs`, bedGraph output.
A "step" is one whole coverage job over that BAM.

  value   reads/s with the compressed BAM already resident in HBM: K1 inflate -> K2 record index -> K3/K4 filter +
          scatter -> K5 scan/collapse into rows (device events + host clock around blocking C-ABI calls, max over ranks)
  e2e     the same job through the C ABI with HOST buffers: BAM file (page cache) -> pinned -> H2D -> kernels ->
          bedGraph text on device -> D2H -> output file closed
  roofline  dominant kernel (K1 inflate): algorithmic bytes = compressed bytes read + inflated bytes written per launch,
          over the launch duration measured with CUDA events on the launching stream in a serialised pass
  cpu_baseline  the CPU oracle (port of the reference algorithm) on a bounded sample of the same BAM, all host cores

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
BIN = 50
PILE = dict(bin_size=BIN, norm="rpkm", collapse=True)
FILT = dict(min_mapq=30, proper_pair=True, min_length=20, max_length=1000)  # CLI defaults + the config's flags
SAMPLE_CHROMS = ["chr1", "chr2", "chr3"]


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


def ensure_bam(pairs: int) -> tuple[Path, dict]:
    bam = workdir() / f"cfg2_pe_{pairs}_seed{SEED}.bam"
    meta = bam.with_suffix(".json")
    if not (bam.exists() and meta.exists() and Path(str(bam) + ".bai").exists()):
        tmp = bam.with_suffix(".tmp.bam")
        r = subprocess.run([str(synth_tool()), "--out", str(tmp), "--pairs", str(pairs), "--seed", str(SEED)], check=True, capture_output=True, text=True)
        os.replace(str(tmp) + ".bai", str(bam) + ".bai")
        os.replace(tmp, bam)
        meta.write_text(r.stdout.strip().splitlines()[-1])
    return bam, json.loads(meta.read_text())


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


def oracle_sample(bam: Path, threads: int = 0):
    """CPU oracle on a bounded sample (a few chromosomes of the same BAM, pileup + bedGraph written) -> (reads, seconds)."""
    from oracle import oracle_py as o

    pile = dict(PILE, bam_path=str(bam), n_threads=threads)
    out = outdir() / "oracle_sample.bedgraph"
    t0 = time.perf_counter()
    st, _ = o.pileup_chroms_to_bedgraph(pile, FILT, SAMPLE_CHROMS, out)  # open BAM -> bedGraph of the sample closed
    dt = time.perf_counter() - t0
    out.unlink(missing_ok=True)
    return st["n_total"], dt


def run_reference(args, bam: Path, meta: dict):
    """--impl reference: the reference's CPU algorithm (the C++ oracle port; the Rust crate cannot be built here:
    no cargo/rustc in the image) on the host cores, each step a bounded sample of the same workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    for _ in range(args.warmup):
        oracle_sample(bam)
    n_tot, t_tot = 0, 0.0
    for _ in range(args.steps):
        n, t = oracle_sample(bam)
        n_tot += n
        t_tot += t
    v = n_tot / t_tot
    sample = f"bam-coverage (pileup + bedGraph written) over {'+'.join(SAMPLE_CHROMS)} of the same BAM per step ({n_tot // max(args.steps, 1)} reads/step)"
    line = {
        "impl": "reference", "metric": "bam-coverage reads/sec", "value": v, "unit": "reads/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_tot / max(args.steps, 1), "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "int32", "data": "synthetic",
        "config": config_dict(args, meta),
        "cpu_baseline": {"value": v, "unit": "reads/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "compressed_GBps": v * meta["file_bytes"] / meta["records"] / 1e9,
    }
    emit(line)


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


def config_dict(args, meta):
    return {"workload": f"configs[1]: synthetic {args.pairs}-pair paired-end hg38 BAM (seed {SEED}, zlib-6, {meta['records']} records, "
                        f"{meta['file_bytes']} B compressed, {meta['inflated_bytes']} B inflated), bam-coverage --bin-size 50 --normalize rpkm "
                        f"--min-mapq 30 --proper-pairs, bedGraph",
            "pairs": args.pairs, "bin_size": BIN, "parallelism": f"chunk-sharded x{args.gpus}" if args.gpus > 1 else "single GPU",
            "l2": "inputs (compressed + inflated stream) are far larger than the 126 MB L2; no explicit flush"}


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
    ap.add_argument("--pairs", type=int, default=int(os.environ.get("BND_BENCH_PAIRS", "100000000")))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    claim_stdout()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        if rank == 0:
            bam, meta = ensure_bam(args.pairs)
            run_reference(args, bam, meta)
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
        bam, meta = ensure_bam(args.pairs)
    barrier()
    bam, meta = ensure_bam(args.pairs)

    from bamnado_b200 import _native

    N = _native.lib()
    pile = dict(PILE, bam_path=str(bam), device=local_rank, shard_rank=rank, shard_world=world)
    out_paths = []

    def allreduce_stats(h):
        st = h.stats()
        if world > 1:
            t = torch.tensor(list(st.values()), dtype=torch.int64, device="cuda")
            dist.all_reduce(t)  # 13 x u64 over NCCL: the only exchange on this path (CPM/RPKM denominator)
            h.set_total_stats([int(x) for x in t.tolist()])

    # ---- e2e leg: host file -> bedGraph file through the C ABI ----
    def e2e_step():
        out_path = outdir() / f"out_rank{rank}_{len(out_paths)}.bedgraph"  # a fresh output file per job
        out_paths.append(out_path)
        with N.pileup(pile, FILT) as h:
            if world > 1:  # sharded: the ranks meet once to sum the 13 filter counters before normalising
                h.accumulate()
                allreduce_stats(h)
                h.finalize()
            h.to_bedgraph(out_path)  # single GPU: one call = BamPileup::to_bedgraph (pile-up, normalise, write)
            return h.shape(), h.timings()

    for _ in range(args.warmup):
        e2e_step()
    for p in out_paths:
        p.unlink(missing_ok=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        shape, tm_e2e = e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    out_bytes = out_paths[-1].stat().st_size
    for p in out_paths:
        p.unlink(missing_ok=True)

    # ---- resident leg: compressed bytes already in HBM ----
    h = N.pileup(pile, FILT)
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
        n_launch = max(1, sh["launches"] // 5)  # 5 kernels per segment: inflate, guess+walk, fix+scan, compact, pileup
        alg_bytes = sh["comp_bytes"] + sh["inflated_bytes"]  # per pass; DESIGN.md: C + U bytes per record
        peak, peak_src = 6451.5, "fallback of B200_PROFILING.md"
        try:
            peaks = json.loads((ROOT / "MEASURED_PEAKS.json").read_text())
            peak, peak_src = float(peaks["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
        except Exception:
            peak = 6650.0
        achieved = alg_bytes / (ts["inflate_ms"] * 1e-3) / 1e9
        roofline = {"kernel": "bgzf_inflate_spec_kernel (K1)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None, "peak_source": peak_src, "launches_timed": n_launch,
                    "avg_launch_ms": ts["inflate_ms"] / n_launch, "algorithmic_bytes_per_launch": alg_bytes / n_launch,
                    "inflated_GBps": sh["inflated_bytes"] / (ts["inflate_ms"] * 1e-3) / 1e9,
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
                roofline["traffic"] = json.loads(traffic_file.read_text()).get("inflate_dram_bytes_per_launch")
            except Exception:
                pass
    h.close()

    cpu = None
    if rank == 0 and not args.no_cpu_baseline:
        n, t = oracle_sample(bam)
        cpu = {"value": n / t, "unit": "reads/s", "cores": os.cpu_count(), "kind": "port",
               "sample": f"CPU oracle (C++ port of the reference algorithm; Rust toolchain absent), pileup + bedGraph written, over {'+'.join(SAMPLE_CHROMS)} of the same BAM: {n} reads in {t:.2f} s"}

    if rank == 0:
        records = meta["records"]
        value = records * args.steps / res_s
        e2e = records * args.steps / e2e_s
        h2d_peak = measure_h2d_gbps(torch)
        line = {
            "metric": "bam-coverage reads/sec", "value": value, "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * res_s / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "int32",
            "data": "synthetic", "config": config_dict(args, meta),
            "compressed_GBps": meta["file_bytes"] * args.steps / res_s / 1e9,
            "e2e": {"value": e2e, "unit": "reads/s", "ms_per_step": 1e3 * e2e_s / args.steps, "compressed_GBps": meta["file_bytes"] * args.steps / e2e_s / 1e9,
                    "h2d_bytes_per_step": h2d_total, "d2h_bytes_per_step": out_bytes, "stages_ms_last_step": tm_e2e,
                    # the end-to-end path against the measured pinned H2D bandwidth of this box
                    "h2d_peak_GBps": h2d_peak,
                    "frac_of_h2d_roofline": (meta["file_bytes"] * args.steps / e2e_s / 1e9 / h2d_peak / world) if h2d_peak else None},
            "gpu_launches": launches, "rows": rows, "host_cores": os.cpu_count(), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "stages_ms_resident_last_step": tm_res,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
