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


# ------------------------------------------------------------------------------------------------ distributed NTT
# Four-step transform of size N = G * L over G ranks: rank r owns the contiguous slice
# x[r*L : (r+1)*L] (n = n1*L + n2 with n1 = r).
#   1. all-to-all: column block j (n2 in [j*S, (j+1)*S), S = L / G) of every rank goes to rank j
#   2. cross kernel (uzkge_cuda_ntt_cross_fr_device): G-point transforms over n1 + twiddle w_N^(n2*k1)
#   3. all-to-all: row k1 goes to rank k1, which now holds B[k1][n2], n2 = 0..L-1
#   4. local size-L transform (uzkge_cuda_ntt_fr_device) -> X[k1 + G*k2], k2 = 0..L-1   ("cyclic" layout)
#   5. optional all-to-all + local interleave -> natural contiguous slices X[r*L : (r+1)*L]
# The inverse runs the same steps with conjugate roots and the 1/N factor split as 1/G (cross) * 1/L (local).


class CudaNttOps:
    """The compute steps on torch CUDA tensors (int64 views of (n, 4) uint64 limbs) through the device ABI."""

    def __init__(self, stream: int = 0):
        self.stream = stream

    def empty_like(self, t):
        import torch

        return torch.empty_like(t)

    def cross(self, t_in, log_g: int, cols: int, col_offset: int, n_total: int, inverse: bool):
        out = self.empty_like(t_in)
        ffi.ntt_cross_fr_device(t_in.data_ptr(), out.data_ptr(), log_g, cols, col_offset, n_total, inverse, self.stream)
        return out

    def local_ntt(self, t_in, n: int, inverse: bool):
        out, scratch = self.empty_like(t_in), self.empty_like(t_in)
        ffi.ntt_fr_device(t_in.data_ptr(), out.data_ptr(), scratch.data_ptr(), n, n, inverse, None, self.stream)
        return out


def _check_dist_sizes(n_total: int, world: int):
    if world not in (2, 4, 8):
        raise ValueError("distributed NTT supports 2, 4 or 8 ranks")
    if n_total & (n_total - 1) or n_total < world * world:
        raise ValueError("distributed NTT needs a power-of-two size >= world^2")
    L = n_total // world
    return L, L // world, world.bit_length() - 1


def ntt_fr_distributed(x_local, n_total: int, rank: int, world: int, inverse: bool = False, natural_output: bool = True,
                       group=None, ops=None):
    """x_local: this rank's contiguous slice as a flat int64 torch tensor of 4 * L words.  Returns the rank's slice of the
    transform: natural contiguous (default) or cyclic (X[rank + world * k2]) when natural_output is False."""
    import torch.distributed as dist

    ops = ops or CudaNttOps()
    L, S, log_g = _check_dist_sizes(n_total, world)
    assert x_local.numel() == 4 * L
    recv = ops.empty_like(x_local)
    dist.all_to_all_single(recv, x_local.contiguous(), group=group)                    # 1
    crossed = ops.cross(recv, log_g, S, rank * S, n_total, inverse)                    # 2
    rows = ops.empty_like(crossed)
    dist.all_to_all_single(rows, crossed, group=group)                                 # 3
    y = ops.local_ntt(rows, L, inverse)                                                # 4
    if not natural_output:
        return y
    back = ops.empty_like(y)
    dist.all_to_all_single(back, y, group=group)                                       # 5
    return back.view(world, S, 4).permute(1, 0, 2).contiguous().view(-1)


def ntt_fr_distributed_emulated(xs, n_total: int, inverse: bool = False, natural_output: bool = True, ops=None):
    """The same steps with all `world` ranks' slices held by ONE process (list of tensors) and the all-to-alls done by
    indexing: runs the multi-rank path on a single GPU (tests) -- no collective, no waiting kernels."""
    import torch

    ops = ops or CudaNttOps()
    world = len(xs)
    L, S, log_g = _check_dist_sizes(n_total, world)

    def all_to_all(ts):
        views = [t.view(world, S * 4) for t in ts]
        return [torch.cat([views[src][dst] for src in range(world)]).contiguous() for dst in range(world)]

    recv = all_to_all(xs)
    crossed = [ops.cross(recv[r], log_g, S, r * S, n_total, inverse) for r in range(world)]
    rows = all_to_all(crossed)
    ys = [ops.local_ntt(rows[r], L, inverse) for r in range(world)]
    if not natural_output:
        return ys
    back = all_to_all(ys)
    return [b.view(world, S, 4).permute(1, 0, 2).contiguous().view(-1) for b in back]


class _RawDeviceArray:
    """Zero-copy torch view of a library-owned device buffer (torch.as_tensor reads __cuda_array_interface__)."""

    def __init__(self, ptr: int, words: int):
        self.__cuda_array_interface__ = {"shape": (words,), "typestr": "<i8", "data": (ptr, False), "version": 2}


class PeerNtt:
    """The four-step transform with BOTH exchanges fused into the cross-rank kernel over peer memory (NVLink P2P), one process per
    GPU.  Every rank owns two buffers other ranks map through cudaIpc: `x` (its contiguous input slice) and `rows` (its receive
    buffer).  One launch of the cross kernel per rank loads the rank's column block straight from every peer's `x` (instead of
    all-to-all no. 1), does the G-point transforms and twiddles, and stores row k1 straight into rank k1's `rows` (instead of
    all-to-all no. 2); the size-L local transform follows.  Two stream-ordered barriers (1-element NCCL all-reduce) order the
    ranks: inputs in place before the kernel, stores landed after it.  Output: the cyclic layout X[rank + world * k2], the form a
    following pointwise stage consumes (ntt_fr_distributed(natural_output=False) is the NCCL-only equivalent)."""

    def __init__(self, n_total: int, rank: int, world: int, device, group=None):
        import torch
        import torch.distributed as dist

        self.n_total, self.rank, self.world, self.group = n_total, rank, world, group
        self.L, self.S, self.log_g = _check_dist_sizes(n_total, world)
        nbytes = 32 * self.L
        self.x, self.rows, self.y, self.scratch = (ffi.dev_alloc(nbytes) for _ in range(4))
        mine = (ffi.ipc_export(self.x), ffi.ipc_export(self.rows), ffi.ipc_export(self.y))
        handles = [None] * world
        dist.all_gather_object(handles, mine, group=group)
        self.peer_x = [self.x if r == rank else ffi.ipc_open(handles[r][0]) for r in range(world)]
        self.peer_rows = [self.rows if r == rank else ffi.ipc_open(handles[r][1]) for r in range(world)]
        self.peer_y = [self.y if r == rank else ffi.ipc_open(handles[r][2]) for r in range(world)]
        self.x_view = torch.as_tensor(_RawDeviceArray(self.x, 4 * self.L), device=device)
        self.y_view = torch.as_tensor(_RawDeviceArray(self.y, 4 * self.L), device=device)
        self._flag = torch.zeros(1, dtype=torch.float32, device=device)

    def _barrier(self):
        import torch.distributed as dist

        dist.all_reduce(self._flag, group=self.group)   # stream-ordered: no host synchronisation

    def transform(self, inverse: bool = False):
        """Transforms the slice currently in `x_view`; returns `y_view` (cyclic layout), valid until the next call."""
        off = 32 * self.rank * self.S
        self._barrier()
        ffi.ntt_cross_rows_fr_device([p + off for p in self.peer_x], [p + off for p in self.peer_rows], self.log_g, self.S,
                                     self.rank * self.S, self.n_total, inverse)
        self._barrier()
        ffi.ntt_fr_device(self.rows, self.y, self.scratch, self.L, self.L, inverse, None)
        return self.y_view

    def transform_natural(self, inverse: bool = False):
        """The same transform with NATURAL output: `y_view` of rank r receives X[r L .. (r + 1) L).  The local transform's last pass
        stores every output straight into its owner's slice over peer memory (uzkge_cuda_ntt_fr_scatter_device): no third exchange,
        one more stream-ordered barrier so that every rank's stores have landed before the result is read."""
        off = 32 * self.rank * self.S
        self._barrier()
        ffi.ntt_cross_rows_fr_device([p + off for p in self.peer_x], [p + off for p in self.peer_rows], self.log_g, self.S,
                                     self.rank * self.S, self.n_total, inverse)
        self._barrier()
        ffi.ntt_fr_scatter_device(self.rows, self.peer_y, self.scratch, self.L, inverse, self.log_g, self.rank)
        self._barrier()
        return self.y_view

    def close(self):
        import torch
        import torch.distributed as dist

        torch.cuda.synchronize()
        dist.barrier(group=self.group)
        for r in range(self.world):
            if r != self.rank:
                ffi.ipc_close(self.peer_x[r])
                ffi.ipc_close(self.peer_rows[r])
                ffi.ipc_close(self.peer_y[r])
        dist.barrier(group=self.group)
        for p in (self.x, self.rows, self.y, self.scratch):
            ffi.dev_free(p)


# ------------------------------------------------------------------------------------------------ point-split commitments for the prover
_FQ = 21888242871839275222246405745257275088696311157297823662689037894645226208583
_FQ_R = (1 << 256) % _FQ
_FQ_R_INV = pow(_FQ_R, -1, _FQ)


def _jac_add_host(a, b) -> np.ndarray:
    """a + b on Jacobian points given as 12 Montgomery limbs each: the world - 1 combinations of a split commitment, O(1) group
    operations on the host (a device round trip per addition costs ten times the arithmetic).  add-2007-bl / dbl-2009-l, a = 0."""
    def load(v):
        v = np.asarray(v, dtype=np.uint64).reshape(12)
        return [(int(v[4 * i]) | int(v[4 * i + 1]) << 64 | int(v[4 * i + 2]) << 128 | int(v[4 * i + 3]) << 192) * _FQ_R_INV % _FQ for i in range(3)]

    def store(x, y, z):
        out = np.zeros(12, dtype=np.uint64)
        for i, c in enumerate((x, y, z)):
            m = c % _FQ * _FQ_R % _FQ
            for j in range(4):
                out[4 * i + j] = (m >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
        return out

    x1, y1, z1 = load(a)
    x2, y2, z2 = load(b)
    if z1 == 0:
        return store(x2, y2, z2) if z2 else store(1, 1, 0)
    if z2 == 0:
        return store(x1, y1, z1)
    z1z1, z2z2 = z1 * z1 % _FQ, z2 * z2 % _FQ
    u1, u2 = x1 * z2z2 % _FQ, x2 * z1z1 % _FQ
    s1, s2 = y1 * z2 * z2z2 % _FQ, y2 * z1 * z1z1 % _FQ
    if u1 == u2:
        if s1 != s2:
            return store(1, 1, 0)
        a_, b_ = x1 * x1 % _FQ, y1 * y1 % _FQ
        c_ = b_ * b_ % _FQ
        d_ = 2 * ((x1 + b_) * (x1 + b_) - a_ - c_) % _FQ
        e_ = 3 * a_ % _FQ
        x3 = (e_ * e_ - 2 * d_) % _FQ
        return store(x3, e_ * (d_ - x3) - 8 * c_, 2 * y1 * z1)
    h, r = (u2 - u1) % _FQ, (s2 - s1) % _FQ
    hh = h * h % _FQ
    hhh, v = h * hh % _FQ, u1 * hh % _FQ
    x3 = (r * r - hhh - 2 * v) % _FQ
    return store(x3, r * (v - x3) - s1 * hhh, z1 * z2 * h)


class _Slice:
    """Window into a DevVec-like object (`.t`, element offset): what CosetParams.quotient writes a coset's values to."""

    def __init__(self, base, first: int):
        self.t, self.ptr = base.t, base.t.data_ptr() + 32 * first


class SplitCommitter:
    """KZG commitments of DEVICE-RESIDENT coefficient vectors with the SRS points split over the ranks (SURVEY 8e, row "MSM"),
    driven by rank 0 -- the rank that runs the prover (uzkge_b200/plonk.py) and owns the polynomials.

    Rank r keeps the bases [r * chunk, (r + 1) * chunk) resident (window tables included).  Per commitment: rank 0 broadcasts a
    16-byte header, scatters the scalar vector in `chunk`-sized pieces (NCCL scatter over NVLink: 32 B per point leave rank 0
    once), every rank runs the MSM over its piece, the 96-byte partial sums are gathered to rank 0 and combined with world - 1
    projective additions.  Ranks > 0 sit in `serve()` until rank 0 calls `shutdown()`.

    `msm_fn(handle, t_scalars, n) -> 12 x int64 tensor` and `add_fn` default to the CUDA backend; the CPU tests (gloo) inject the
    checker to exercise the protocol without a GPU.
    """

    OP_EXIT, OP_COMMIT, OP_TRANSFORM, OP_QSETUP, OP_QUOTIENT = 0, 1, 2, 3, 4

    def __init__(self, affine_xy: np.ndarray, rank: int, world: int, device=None, group=None, window_bits: int = 0,
                 upload: Callable = None, msm_fn: Callable = None, add_fn: Callable = None, ntt_fn: Callable = None):
        import torch

        pts = ffi.as_u64(affine_xy, 8)
        self.n = pts.shape[0]
        self.rank, self.world, self.group = rank, world, group
        self.device = device if device is not None else torch.device("cpu")
        self.chunk = (self.n + world - 1) // world
        self.lo = min(self.n, rank * self.chunk)
        self.hi = min(self.n, self.lo + self.chunk)
        upload = upload or (lambda p: ffi.srs_upload(p, window_bits))
        self.handle = upload(pts[self.lo: self.hi]) if self.hi > self.lo else None
        self._msm = msm_fn or self._msm_cuda
        self._add = add_fn or _jac_add_host
        self._recv = torch.zeros(4 * self.chunk, dtype=torch.int64, device=self.device)
        self._pad = torch.zeros(4 * self.chunk * world, dtype=torch.int64, device=self.device) if rank == 0 else None
        self._hdr = torch.zeros(8, dtype=torch.int64, device=self.device)   # op, a, b, c, 4 limbs of a coset shift
        self._ntt = ntt_fn or self._ntt_cuda
        self._tbuf = {}
        self._part = torch.zeros(12, dtype=torch.int64, device=self.device)
        self._parts = [torch.zeros(12, dtype=torch.int64, device=self.device) for _ in range(world)] if rank == 0 else None
        self.commits = 0

    def max_degree(self) -> int:
        return self.n - 1

    def _msm_cuda(self, handle, t_scalars, n: int):
        ffi.msm_g1_device(handle, t_scalars.data_ptr(), n, self._part.data_ptr())
        return self._part

    def _step(self, length: int, src_tensor=None) -> np.ndarray | None:
        """One commitment; collective: every rank calls it with the same `length`."""
        import torch.distributed as dist

        if self.rank == 0:
            self._pad[: 4 * length].copy_(src_tensor[: 4 * length])
            dist.scatter(self._recv, list(self._pad.split(4 * self.chunk)), src=0, group=self.group)
        else:
            dist.scatter(self._recv, None, src=0, group=self.group)
        mine = max(0, min(length, self.hi) - self.lo)
        if mine and self.handle is not None:
            part = self._msm(self.handle, self._recv, mine)
        else:
            part = self._part.zero_()   # Z = 0: the identity
        dist.gather(part, self._parts, dst=0, group=self.group)
        if self.rank != 0:
            return None
        parts = [p.cpu().numpy().view(np.uint64) for p in self._parts]
        acc = parts[0]
        for p in parts[1:]:
            acc = self._add(acc, p)
        return np.asarray(acc, dtype=np.uint64).reshape(12).copy()

    def commit_device(self, vecs) -> list:
        """Rank 0: commit to device vectors (objects with `.t` = int64 tensor of 4 words per element and `.len`)."""
        import torch.distributed as dist

        from .errors import DegreeError
        from .poly_commit import KZGCommitment

        assert self.rank == 0, "only rank 0 drives commitments; the other ranks call serve()"
        out = []
        for v in vecs:
            if v.len > self.n:
                raise DegreeError("DegreeError")
            self._hdr[0], self._hdr[1] = self.OP_COMMIT, v.len
            dist.broadcast(self._hdr, src=0, group=self.group)
            out.append(KZGCommitment(self._step(v.len, v.t)))
            self.commits += 1
        return out

    # ---- independent transforms of one prover round, one polynomial per rank ("replicas": each fits one GPU)
    def _ntt_cuda(self, t_in, len_in: int, domain_size: int, inverse: bool, shift, t_out, t_scratch):
        ffi.ntt_fr_device(t_in.data_ptr(), t_out.data_ptr(), t_scratch.data_ptr(), len_in, domain_size, inverse, shift)
        return t_out

    def _buffers(self, len_in: int, domain_size: int):
        import torch

        key = (len_in, domain_size)
        if key not in self._tbuf:
            mk = lambda n: torch.zeros(4 * n, dtype=torch.int64, device=self.device)
            self._tbuf = {key: (mk(len_in), mk(domain_size), mk(domain_size))}    # one size at a time: the 6n vectors are large
        return self._tbuf[key]

    def _transform_step(self, count: int, len_in: int, domain_size: int, inverse: bool, shift, jobs=None):
        """Collective.  Polynomial i is transformed on rank i % world; rank 0 sends the coefficients and receives the values."""
        import torch.distributed as dist

        if self.rank == 0:
            # rounds of `world` polynomials: at most one per rank and round, so every rank's send / recv sequence pairs up in
            # order with rank 0's (no rank waits on a send whose receive is queued behind another send)
            scratch = self._buffers(len_in, domain_size)[2]
            for base in range(0, count, self.world):
                batch = list(range(base, min(count, base + self.world)))
                pending = [dist.isend(jobs[i][0][: 4 * len_in].contiguous(), i % self.world, group=self.group) for i in batch if i % self.world]
                for i in batch:
                    if i % self.world == 0:
                        self._ntt(jobs[i][0], len_in, domain_size, inverse, shift, jobs[i][1], scratch)
                for w_ in pending:
                    w_.wait()
                for i in batch:
                    if i % self.world:
                        dist.recv(jobs[i][1][: 4 * domain_size], i % self.world, group=self.group)
            return
        mine = [i for i in range(count) if i % self.world == self.rank]
        if not mine:
            return
        t_in, t_out, t_scr = self._buffers(len_in, domain_size)
        for _ in mine:
            dist.recv(t_in, 0, group=self.group)
            out = self._ntt(t_in, len_in, domain_size, inverse, shift, t_out, t_scr)
            dist.send(out[: 4 * domain_size], 0, group=self.group)

    def transform_many(self, jobs, len_in: int, domain_size: int, inverse: bool = False, coset_shift=None) -> None:
        """Rank 0: `jobs` = [(src tensor holding >= len_in elements, dst tensor holding domain_size elements)]: dst = the
        (coset) transform of src, natural order -- FpPolynomial::coset_fft_with_domain for the wire and z polynomials of the
        quotient round (plonk/helpers.rs:256-266), one polynomial per GPU."""
        import torch.distributed as dist

        assert self.rank == 0
        self._hdr.zero_()
        self._hdr[0], self._hdr[1], self._hdr[2], self._hdr[3] = self.OP_TRANSFORM, len(jobs), len_in, domain_size
        shift = None
        if coset_shift is not None:
            shift = np.ascontiguousarray(coset_shift, dtype=np.uint64).reshape(4)
            self._hdr[4:8] = self._hdr.new_tensor(shift.view(np.int64).tolist())
            self._hdr[3] = domain_size | (1 << 62)
        if inverse:
            self._hdr[2] = len_in | (1 << 62)
        dist.broadcast(self._hdr, src=0, group=self.group)
        self._transform_step(len(jobs), len_in, domain_size, inverse, shift, jobs)

    # ---- the quotient round, one coset (or a few) of the 6n domain per rank (plonk.CosetParams)
    def _my_cosets(self):
        return [j for j in range(6) if j % self.world == self.rank]

    def _quotient_setup_step(self, n: int, k_t, poly_ts) -> None:
        """Collective.  k_t: 20 int64 (k[0..5) limbs), poly_ts: the 19 coefficient vectors q[9], s[5], qb, q_prk[4] (n elements each;
        filled on rank 0, receive buffers elsewhere)."""
        import torch
        import torch.distributed as dist

        from . import plonk

        dist.broadcast(k_t, src=0, group=self.group)
        for t in poly_ts:
            dist.broadcast(t, src=0, group=self.group)
        k = [plonk.unmont(row) for row in k_t.cpu().numpy().view(np.uint64).reshape(5, 4)]

        class _Vec:                               # (pointer, length) view of a received tensor
            def __init__(self, t):
                self.t, self.ptr, self.len = t, t.data_ptr(), n

        vecs = [_Vec(t) for t in poly_ts]
        self._q = {"n": n, "k": k, "params": {j: plonk.CosetParams(vecs[0:9], vecs[9:14], vecs[14], vecs[15:19], k, n, j, self.device)
                                              for j in self._my_cosets()},
                   "polys": [torch.zeros(4 * (n + 4), dtype=torch.int64, device=self.device) for _ in range(7)],
                   "scal": torch.zeros(12, dtype=torch.int64, device=self.device),
                   "out": plonk.DevVec(n, self.device, zero=False)}

    def _quotient_step(self, n: int, has_pi: bool, t_cosets=None) -> None:
        """Collective.  The polynomials (w0..w4, z of n + 3 coefficients, pi of n) and alpha, beta, gamma are already in
        self._q["polys"] / ["scal"] on rank 0."""
        import torch.distributed as dist

        from . import plonk

        q = self._q
        dist.broadcast(q["scal"], src=0, group=self.group)
        for i in range(7 if has_pi else 6):
            dist.broadcast(q["polys"][i], src=0, group=self.group)
        alpha, beta, gamma = (plonk.unmont(row) for row in q["scal"].cpu().numpy().view(np.uint64).reshape(3, 4))
        polys = [(q["polys"][i], n + 3) for i in range(6)] + [(q["polys"][6], n) if has_pi else None]
        if self.rank == 0:
            for j in self._my_cosets():
                q["params"][j].quotient(polys, q["k"], alpha, beta, gamma, _Slice(t_cosets, j * n))
            for j in range(6):
                if j % self.world:
                    dist.recv(t_cosets.t[4 * j * n: 4 * (j + 1) * n], j % self.world, group=self.group)
        else:
            for j in self._my_cosets():
                q["params"][j].quotient(polys, q["k"], alpha, beta, gamma, q["out"])
                dist.send(q["out"].t[: 4 * n], 0, group=self.group)

    def quotient_by_cosets(self, P, polys, k, alpha: int, beta: int, gamma: int, t_cosets) -> None:
        """Rank 0 (called by plonk.prover): t on the six cosets of the quotient domain into t_cosets[j * n + i], the cosets dealt to
        the ranks.  The preprocessed polynomials travel once (first call), the round's seven polynomials every proof."""
        import torch
        import torch.distributed as dist

        from . import plonk

        assert self.rank == 0
        n = P.n
        if getattr(self, "_q", None) is None or self._q["n"] != n or self._q.get("owner") is not P:
            self._hdr.zero_()
            self._hdr[0], self._hdr[1] = self.OP_QSETUP, n
            dist.broadcast(self._hdr, src=0, group=self.group)
            k_t = torch.from_numpy(plonk.mont_rows(k).view(np.int64).reshape(-1).copy()).to(self.device)
            poly_ts = []
            for v in list(P.q_polys) + list(P.s_polys) + [P.qb_poly] + list(P.q_prk_polys):
                t = torch.zeros(4 * n, dtype=torch.int64, device=self.device)
                t[: 4 * min(v.len, n)].copy_(v.t[: 4 * min(v.len, n)])
                poly_ts.append(t)
            self._quotient_setup_step(n, k_t, poly_ts)
            self._q["owner"] = P
        q = self._q
        q["scal"].copy_(torch.from_numpy(plonk.mont_rows([alpha, beta, gamma]).view(np.int64).reshape(-1).copy()))
        has_pi = polys[6] is not None
        for i, pl in enumerate(polys):
            if pl is not None:
                q["polys"][i][: 4 * pl[1]].copy_(pl[0][: 4 * pl[1]])
        self._hdr.zero_()
        self._hdr[0], self._hdr[1], self._hdr[2] = self.OP_QUOTIENT, n, 1 if has_pi else 0
        dist.broadcast(self._hdr, src=0, group=self.group)
        self._quotient_step(n, has_pi, t_cosets)

    def serve(self) -> int:
        """Ranks > 0: answer rank 0's commitments and transforms until it shuts the service down.  Returns the number served."""
        import torch.distributed as dist

        assert self.rank != 0
        while True:
            dist.broadcast(self._hdr, src=0, group=self.group)
            h = [int(x) for x in self._hdr.cpu()]
            op = h[0]
            if op == self.OP_EXIT:
                return self.commits
            if op == self.OP_QSETUP:
                import torch

                n = h[1]
                k_t = torch.zeros(20, dtype=torch.int64, device=self.device)
                poly_ts = [torch.zeros(4 * n, dtype=torch.int64, device=self.device) for _ in range(19)]
                self._quotient_setup_step(n, k_t, poly_ts)
                continue
            if op == self.OP_QUOTIENT:
                self._quotient_step(h[1], bool(h[2]))
                continue
            if op == self.OP_TRANSFORM:
                count, len_in, dom = h[1], h[2] & ((1 << 62) - 1), h[3] & ((1 << 62) - 1)
                inverse, has_shift = bool(h[2] >> 62), bool(h[3] >> 62)
                shift = np.array(h[4:8], dtype=np.int64).view(np.uint64) if has_shift else None
                self._transform_step(count, len_in, dom, inverse, shift)
                continue
            self._step(h[1])
            self.commits += 1

    def shutdown(self) -> None:
        import torch.distributed as dist

        assert self.rank == 0
        self._hdr[0], self._hdr[1] = self.OP_EXIT, 0
        dist.broadcast(self._hdr, src=0, group=self.group)

    def close(self) -> None:
        if self.handle is not None:
            ffi.srs_free(self.handle)
            self.handle = None
