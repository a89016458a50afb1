"""Multi-GPU partitioning of the hot path.  Host-side logic only; one process per GPU.

  * batches of independent windows: contiguous window ranges per rank, no data-path collective;
  * association sweep: contiguous pose ranges per rank, the map is replicated;
  * ONE huge window: factors are split by LANDMARK so every rank can eliminate its own landmarks locally
    (their H_ll/H_lp blocks are disjoint); only the reduced pose system [S | g] is summed across ranks
    (one all-reduce).  Line factors are dealt round-robin.
"""
import numpy as np

from ._abi import Batch


def split_range(n, rank, world):
    """Contiguous share [lo, hi) of n items for `rank`; sizes differ by at most one."""
    lo = (n * rank) // world
    hi = (n * (rank + 1)) // world
    return lo, hi


def shard_windows(batch, rank, world):
    lo, hi = split_range(batch.W, rank, world)
    return batch.slice_windows(lo, hi)


def split_huge_window(batch, rank, world):
    """Rank's share of the factors of a single-window batch (W == 1): landmarks l with l % world == rank, renumbered densely
    (local landmark l // world; the rank's inverse-depth array and landmark rows hold F/world entries, not F), line factors k
    with k % world == rank.  Poses and the extrinsic are replicated."""
    assert batch.W == 1
    feat = (batch.pf_idx >> 16).astype(np.int64)
    keep = (feat % world) == rank
    lkeep = (np.arange(batch.NL) % world) == rank
    z = None if batch.pf_pts_i_z is None else batch.pf_pts_i_z[keep]
    idx = batch.pf_idx[keep]
    idx = ((idx & 0xffff) | (((idx >> 16) // world) << 16)).astype(np.uint32)
    owned = np.arange(rank, batch.F, world)
    dep = batch.inv_depth[:, owned] if len(owned) else np.zeros((1, 1))
    return Batch(batch.poses, batch.ex_pose, dep,
                 np.array([0, int(keep.sum())], dtype=np.int32), idx, batch.pf_obs[keep],
                 np.array([0, int(lkeep.sum())], dtype=np.int32), batch.lf_frame[lkeep], batch.lf_geom[:, lkeep], z)


def pack_sg(S, g):
    """[S | g] as one contiguous buffer (D*D + D doubles) for a single all-reduce."""
    return np.concatenate([np.asarray(S).reshape(-1), np.asarray(g).reshape(-1)])


def unpack_sg(buf, D):
    return buf[:D * D].reshape(D, D), buf[D * D:D * D + D]
