"""N > 1 path on CPU: two `gloo` ranks, each piles up its shard of the genomic chunks (emulator backend), the 13 filter
counters are summed with one all-reduce, and the concatenated rows equal the oracle's whole-genome rows."""
import os
import subprocess

import pytest
import sys
import textwrap

import numpy as np

from conftest import ROOT

WORKER = textwrap.dedent("""
    import ctypes, os, sys
    import numpy as np
    import torch, torch.distributed as dist
    sys.path.insert(0, {root!r})
    from bamnado_b200 import _native
    from bamnado_b200.build import build_emul
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    N = _native.Native(ctypes.CDLL(str(build_emul())))
    pile = dict(bam_path={bam!r}, bin_size=50, norm="rpkm", shard_rank=rank, shard_world=world)
    filt = dict(min_mapq=30, proper_pair=True)
    with N.pileup(pile, filt) as h:
        if rank == 0:
            h.preallocate_output({out!r} + ".merged.bdg")  # one rank allocates the merged file's pages behind its pile-up
        h.accumulate()
        t = torch.tensor(list(h.stats().values()), dtype=torch.int64)
        dist.all_reduce(t)
        h.set_total_stats(t.tolist())
        h.finalize()
        np.save({out!r} + f".runs{{rank}}.npy", h.runs())
        if rank == 0:
            np.save({out!r} + ".stats.npy", t.numpy())
            open({out!r} + ".factor", "w").write(repr(h.norm_factor()))
        text = h.bedgraph_text()
        open({out!r} + f".text{{rank}}", "wb").write(text)
        # ONE sorted bedGraph written by all ranks: all-gather of the per-reference byte counts, then every rank fills its pieces
        sizes = torch.tensor(h.format_bedgraph() + [h.preallocated_bytes()], dtype=torch.int64)
        tables = [torch.empty_like(sizes) for _ in range(world)]
        dist.all_gather(tables, sizes)
        tables = [t.tolist() for t in tables]
        h.write_bedgraph_pieces({out!r} + ".merged.bdg", [t[:-1] for t in tables], rank, max(t[-1] for t in tables))
        dist.barrier()
    dist.destroy_process_group()
""")


def test_two_rank_gloo_sharded_coverage(oracle, make_bam, tmp_path, native_emul):
    bam = make_bam("pe", "--pairs", 20000, "--odd")
    out = str(tmp_path / "dist")
    script = tmp_path / "worker.py"
    script.write_text(WORKER.format(root=str(ROOT), bam=bam, out=out))
    env = dict(os.environ, BND_SEGMENT_BYTES="262144", BND_EMUL_WORKERS="2", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", "29517", str(script)], env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    pile, filt = dict(bam_path=bam, bin_size=50, norm="rpkm"), dict(min_mapq=30, proper_pair=True)
    oruns, ostats, ofactor = oracle.pileup_runs(pile, filt)
    runs = np.concatenate([np.load(out + f".runs{k}.npy") for k in range(2)])
    assert len(runs) == len(oruns) and (runs == oruns).all()
    assert list(np.load(out + ".stats.npy")) == list(ostats.values())
    assert float(open(out + ".factor").read()) == ofactor
    # every rank's text rows use the job-wide factor; together they are the oracle's bedGraph (as a set of lines)
    ref = tmp_path / "oracle.bdg"
    oracle.pileup_to_bedgraph(pile, filt, ref)
    got = sorted(open(out + ".text0", "rb").read().splitlines() + open(out + ".text1", "rb").read().splitlines())
    assert got == sorted(ref.read_bytes().splitlines())
    # the file the ranks wrote together is the reference's artefact byte for byte (write_bedgraph sorts by chrom, start)
    assert open(out + ".merged.bdg", "rb").read() == ref.read_bytes()


@pytest.mark.parametrize("n_bams", [2, 4])
def test_multi_bam_sharded_by_bam_gloo(oracle, make_bam, tmp_path, native_emul, n_bams):
    # §8(e): BAMs sharded over ranks (1 or 2 per rank), all-gather + OR of the 1-bit change masks, all-to-all of the compacted column
    # slices, every rank formats its rows and writes its pieces -> MultiBamPileup TSV
    bams = [make_bam(f"m{i}", "--pairs", 8000 if i < 3 else 3000, "--seed", 11 + i, "--mode", "se") for i in range(n_bams)]
    out = tmp_path / "dist.tsv"
    env = dict(os.environ, BND_EMUL_WORKERS="2", OMP_NUM_THREADS="1")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2", "--master-addr", "127.0.0.1",
                        "--master-port", str(29519 + n_bams), str(ROOT / "scripts" / "multi_bam_dist.py"), "--bams", *bams, "--bin-size", "500", "-o", str(out),
                        "--backend", "gloo", "--emul", "--slice-rows", "64000000" if n_bams == 2 else "1777"], env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    ref = tmp_path / "oracle.tsv"
    filt = dict(min_mapq=20, min_length=20, max_length=1000)
    oracle.multi_to_tsv([dict(bam_path=b, bin_size=500, collapse=False) for b in bams], [filt] * n_bams, ref)
    a, b = out.read_text().splitlines(), ref.read_text().splitlines()
    assert a[0] == b[0] and sorted(a[1:]) == sorted(b[1:])
    # the file the two ranks wrote together is byte for byte what one process holding all BAMs writes
    single = tmp_path / "single.tsv"
    native_emul.multi_to_tsv([dict(bam_path=b, bin_size=500, collapse=False) for b in bams], [filt] * n_bams, single)
    assert out.read_bytes() == single.read_bytes()
