"""World-size-2 tests of the multi-GPU host logic on CPU (gloo): window sharding needs no collective; the single
huge window is split by landmark, every rank eliminates its own landmarks, and one all-reduce of [S | g] gives
the full reduced system.  Compute on each rank is the oracle here (no GPU in this tier); the same partition
and reduction code runs with the CUDA path and NCCL in bench.py and in tests/test_gpu_parity.py."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, rel_err


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import __graft_entry__ as ge
    pkg, orc = ge.load_package(), ge.load_oracle()
    abi, synth, shard = pkg._abi, pkg.synth, pkg.shard
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    cfg = synth.euroc_config()
    flags = abi.OUT_HB | abi.OUT_SCHUR | abi.LOSS_CAUCHY
    # (1) independent windows: contiguous ranges, gather only to compare
    batch = synth.make_windows(7, seed=201, F=40)
    mine = shard.shard_windows(batch, rank, world)
    out = orc.linearize_batch(cfg, mine, flags)
    lo, hi = shard.split_range(batch.W, rank, world)
    full = orc.linearize_batch(cfg, batch, flags)
    ok1 = all(np.array_equal(out[k], full[k][lo:hi]) for k in out)
    counts = [None] * world
    dist.all_gather_object(counts, (lo, hi))
    # (2) one huge window: split by landmark, local Schur, all-reduce of [S | g]
    huge = synth.make_windows(1, seed=202, P=24, F=90, lines_per_frame=2)
    part = shard.split_huge_window(huge, rank, world)
    po = orc.linearize_batch(cfg, part, flags)
    buf = torch.from_numpy(shard.pack_sg(po["S"][0], po["g"][0]))
    dist.all_reduce(buf)
    S, g = shard.unpack_sg(buf.numpy(), huge.D)
    ref = orc.linearize_batch(cfg, huge, flags)
    nfac = torch.tensor([part.NP, part.NL])
    dist.all_reduce(nfac)
    q.put((rank, ok1, counts, rel_err(S, ref["S"][0]), rel_err(g, ref["g"][0]), nfac.tolist(), [huge.NP, huge.NL]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_huge_window_allreduce(pkg, orc):
    world, port = 2, 29500 + os.getpid() % 2000
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok1, counts, es, eg, nfac, total in res:
        assert ok1
        assert counts[0][0] == 0 and counts[-1][1] == 7 and all(counts[i][1] == counts[i + 1][0] for i in range(world - 1))
        assert es < 1e-12 and eg < 1e-12      # sums commute up to rounding
        assert nfac == total                  # every factor is owned by exactly one rank


def test_split_range_covers_everything(pkg):
    sr = pkg.shard.split_range
    for n in (0, 1, 7, 4096, 65536):
        for world in (1, 2, 3, 4, 8):
            parts = [sr(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1
