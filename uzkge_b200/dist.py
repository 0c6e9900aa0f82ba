"""One process per GPU: how the hot path shards over an 8 x B200 box.

* MSM: points and scalars are split into contiguous slices, one per rank; every rank keeps only ITS slice of the SRS
  resident (tables included) and computes a partial sum; the G partial sums (96 bytes each) are all-gathered and
  combined with G - 1 projective additions.  No other data-path collective exists.
* Independent commitments of one prover round (5 + 3 wire polynomials, 5 quotient pieces, 2 openings:
  /root/reference/uzkge/src/plonk/prover.rs:132-192, helpers.rs:1323-1408): round-robin over ranks with a fully
  replicated SRS; results (96 bytes each) are all-gathered.
* NTT < 2^22 and the Fiat-Shamir transcript: replicas only.

`torch.distributed` is plumbing (NCCL on GPUs; gloo in the CPU tests).  The compute callables default to the CUDA
backend; the CPU test-suite injects its checker to exercise the partitioning and the exchange without a GPU.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np

from . import ffi


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous slice [lo, hi) of n items owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def _all_gather_u64(arr: np.ndarray, group=None, device=None) -> np.ndarray:
    """All-gather a small uint64 array (same shape on every rank) -> (world, *shape)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group)
    t = torch.from_numpy(np.ascontiguousarray(arr).view(np.int64).reshape(-1))
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * t.numel(), dtype=torch.int64, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    return out.cpu().numpy().view(np.uint64).reshape((world,) + arr.shape)


class ShardedSrs:
    """The rank-local slice of an SRS (bases lo..hi resident on this rank's GPU)."""

    def __init__(self, affine_xy: np.ndarray, rank: int, world: int, window_bits: int = 0,
                 upload: Callable = None):
        pts = ffi.as_u64(affine_xy, 8)
        self.n = pts.shape[0]
        self.rank, self.world = rank, world
        self.lo, self.hi = shard_range(self.n, rank, world)
        self.local_points = pts[self.lo : self.hi]
        upload = upload or (lambda p: ffi.srs_upload(p, window_bits))
        self.handle = upload(self.local_points) if self.hi > self.lo else None


def msm_sharded(srs: ShardedSrs, scalars: np.ndarray, group=None, device=None,
                msm_fn: Callable = None, add_fn: Callable = None) -> np.ndarray:
    """sum_i scalars[i] * srs[i] over all ranks' slices; every rank returns the same Jacobian point.

    scalars: the full scalar vector (length <= srs.n, replicated) -- each rank reads only its slice."""
    msm_fn = msm_fn or (lambda handle, s: ffi.msm_g1(handle, s))
    add_fn = add_fn or ffi.g1_add
    s = ffi.as_u64(scalars, 4)
    lo, hi = min(srs.lo, s.shape[0]), min(srs.hi, s.shape[0])
    if hi > lo and srs.handle is not None:
        partial = msm_fn(srs.handle, s[lo:hi])
    else:
        partial = np.zeros(12, dtype=np.uint64)  # Z = 0: identity
    parts = _all_gather_u64(np.asarray(partial, dtype=np.uint64).reshape(12), group, device)
    acc = parts[0]
    for r in range(1, parts.shape[0]):
        acc = add_fn(acc, parts[r])
    return acc


def commit_distributed(polys: Sequence[np.ndarray], rank: int, world: int, commit_fn: Callable, group=None,
                       device=None) -> np.ndarray:
    """Independent commitments of one round, polynomial j on rank j % world (full SRS on every rank).
    Returns (len(polys), 12) on every rank."""
    k = len(polys)
    per = (k + world - 1) // world
    mine = np.zeros((per, 12), dtype=np.uint64)
    for slot, j in enumerate(range(rank, k, world)):
        mine[slot] = commit_fn(polys[j])
    allp = _all_gather_u64(mine, group, device)  # (world, per, 12)
    out = np.zeros((k, 12), dtype=np.uint64)
    for j in range(k):
        out[j] = allp[j % world, j // world]
    return out
