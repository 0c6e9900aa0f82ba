"""CPU, world_size 2, gloo: the N > 1 host logic (slice ownership, the 96-byte all-gather, the combine order) with the
CPU oracle injected as the compute callable -- the CUDA kernels themselves are covered by the -m gpu tests."""
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp


def free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import cpu as oc
        from uzkge_b200 import dist as udist

        pts = oc.g1_random_points(n, 3)
        sc = oc.random_fr(n - 5, 4)  # shorter than the SRS: the last slice is ragged
        store = {}

        def upload(p):
            store[1] = p
            return 1

        srs = udist.ShardedSrs(pts, rank, world, upload=upload)
        assert (srs.lo, srs.hi) == udist.shard_range(n, rank, world)
        got = udist.msm_sharded(srs, sc, msm_fn=lambda h, s: oc.msm_g1(store[h][: s.shape[0]], s), add_fn=oc.g1_add_jac)
        want = oc.msm_g1(pts[: n - 5], sc)
        ok1 = np.array_equal(oc.g1_to_affine(got), oc.g1_to_affine(want))

        polys = [oc.random_fr(m, 10 + m) for m in (7, 64, 33, 1, 50)]
        outs = udist.commit_distributed(polys, rank, world, lambda c: oc.msm_g1(pts[: c.shape[0]], c))
        ok2 = all(np.array_equal(oc.g1_to_affine(outs[j]), oc.g1_to_affine(oc.msm_g1(pts[: p.shape[0]], p)))
                  for j, p in enumerate(polys))
        # scalars shorter than the first slice: the other rank contributes the identity
        short = sc[:3]
        got3 = udist.msm_sharded(srs, short, msm_fn=lambda h, s: oc.msm_g1(store[h][: s.shape[0]], s), add_fn=oc.g1_add_jac)
        ok3 = np.array_equal(oc.g1_to_affine(got3), oc.g1_to_affine(oc.msm_g1(pts[:3], short)))
        # distributed NTT: the exchange pattern with the compute steps injected (numpy + oracle on CPU tensors)
        import torch

        class CpuOps:
            def empty_like(self, t):
                return torch.empty_like(t)

            def cross(self, t_in, log_g, cols, col_offset, n_total, inverse):
                g = 1 << log_g
                a = t_in.numpy().view(np.uint64).reshape(g, cols, 4)
                w = oc.fr_root_of_unity(n_total)
                if inverse:
                    w = oc.fr_inv(w)
                out = np.zeros_like(a)
                for t in range(cols):
                    col = np.ascontiguousarray(a[:, t, :])
                    y = oc.ntt_fr(col, g, inverse=inverse)  # G-point transform (inverse includes 1/G)
                    for k1 in range(g):
                        tw = oc.fr_pow(w, ((col_offset + t) * k1) % n_total)
                        out[k1, t] = oc.fr_mul(y[k1 : k1 + 1], tw.reshape(1, 4))[0]
                return torch.from_numpy(out.view(np.int64).reshape(-1))

            def local_ntt(self, t_in, n, inverse):
                a = t_in.numpy().view(np.uint64).reshape(n, 4)
                return torch.from_numpy(oc.ntt_fr(a, n, inverse=inverse).view(np.int64).reshape(-1))

        nt = 64
        L = nt // world
        x = oc.random_fr(nt, 99)
        mine = torch.from_numpy(np.ascontiguousarray(x[rank * L : (rank + 1) * L]).view(np.int64).reshape(-1))
        y = udist.ntt_fr_distributed(mine, nt, rank, world, ops=CpuOps())
        want = oc.ntt_fr(x, nt)
        ok4 = np.array_equal(y.numpy().view(np.uint64).reshape(L, 4), want[rank * L : (rank + 1) * L])
        yc = udist.ntt_fr_distributed(mine, nt, rank, world, natural_output=False, ops=CpuOps())
        ok4 = ok4 and np.array_equal(yc.numpy().view(np.uint64).reshape(L, 4), want[rank::world])
        back = udist.ntt_fr_distributed(y, nt, rank, world, inverse=True, ops=CpuOps())
        ok4 = ok4 and np.array_equal(back.numpy().view(np.uint64).reshape(L, 4), x[rank * L : (rank + 1) * L])
        q.put((rank, ok1, ok2, ok3 and ok4))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_everything():
    from uzkge_b200.dist import shard_range

    for n in (0, 1, 7, 8, 1 << 20, (1 << 20) + 3):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_sharded_msm_and_distributed_commits_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, n = 2, 301
    port = free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] and r[3] for r in res), res


def _split_worker(rank, world, port, n, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch

        from oracle import cpu as oc
        from uzkge_b200 import dist as udist

        pts = oc.g1_random_points(n, 3)
        store = {}

        def upload(p):
            store[1] = p
            return 1

        def msm_fn(handle, t, k):
            s = t.numpy().view(np.uint64).reshape(-1, 4)[:k]
            return torch.from_numpy(oc.msm_g1(store[handle][:k], s).view(np.int64).copy())

        sc = udist.SplitCommitter(pts, rank, world, upload=upload, msm_fn=msm_fn, add_fn=oc.g1_add_jac)
        assert sc.max_degree() == n - 1
        if rank != 0:
            served = sc.serve()
            q.put((rank, served == 4))
            return

        class Vec:
            def __init__(self, a):
                self.t = torch.from_numpy(a.view(np.int64).reshape(-1).copy())
                self.len = a.shape[0]

        # full length, ragged, shorter than the first slice (the other rank adds the identity), and a single coefficient
        polys = [oc.random_fr(m, 20 + m) for m in (n, n - 7, 5, 1)]
        cms = sc.commit_device([Vec(p) for p in polys])
        ok = all(np.array_equal(oc.g1_to_affine(c.value), oc.g1_to_affine(oc.msm_g1(pts[: p.shape[0]], p))) for c, p in zip(cms, polys))
        sc.shutdown()
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_split_committer_world2():
    """The prover's point-split commitment service (dist.SplitCommitter): header broadcast, scatter, partial MSMs, gather, combine."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world, n = 2, 203
    port = free_port()
    procs = [ctx.Process(target=_split_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res


def _transform_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch

        from oracle import cpu as oc
        from uzkge_b200 import dist as udist

        def ntt_fn(t_in, len_in, dom, inverse, shift, t_out, t_scr):
            a = t_in.numpy().view(np.uint64).reshape(-1, 4)[:len_in]
            y = oc.ntt_fr(a, dom, inverse=inverse, coset=shift)
            t_out[: 4 * dom].copy_(torch.from_numpy(y.view(np.int64).reshape(-1)))
            return t_out

        pts = oc.g1_random_points(8, 3)
        sc = udist.SplitCommitter(pts, rank, world, upload=lambda p: 1, msm_fn=lambda h, t, k: torch.zeros(12, dtype=torch.int64),
                                  add_fn=oc.g1_add_jac, ntt_fn=ntt_fn)
        if rank != 0:
            sc.serve()
            q.put((rank, True))
            return
        len_in, dom = 19, 96                      # n + 3 coefficients on the 6 n domain, like the quotient round
        k = oc.random_fr(1, 5)[0]
        polys = [oc.random_fr(len_in, 30 + i) for i in range(5)]
        jobs = [(torch.from_numpy(p.view(np.int64).reshape(-1).copy()), torch.zeros(4 * dom, dtype=torch.int64)) for p in polys]
        sc.transform_many(jobs, len_in, dom, False, k)
        ok = all(np.array_equal(j[1].numpy().view(np.uint64).reshape(dom, 4), oc.ntt_fr(p, dom, coset=k)) for j, p in zip(jobs, polys))
        jobs2 = [(j[1], torch.zeros(4 * dom, dtype=torch.int64)) for j in jobs[:3]]
        sc.transform_many(jobs2, dom, dom, True, None)          # plain inverse transforms, no shift
        ok = ok and all(np.array_equal(j[1].numpy().view(np.uint64).reshape(dom, 4), oc.ntt_fr(jobs[i][1].numpy().view(np.uint64).reshape(dom, 4), dom, inverse=True))
                        for i, j in enumerate(jobs2))
        sc.shutdown()
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_split_committer_distributes_transforms_world2():
    """dist.SplitCommitter.transform_many: one polynomial per rank and round, coefficients out / evaluations back, with the oracle's
    transform injected -- more polynomials than ranks, with and without the coset shift, forward and inverse."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    world = 2
    port = free_port()
    procs = [ctx.Process(target=_transform_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    assert all(r[1] for r in res), res


def test_host_jacobian_add_matches_oracle(oc):
    """dist._jac_add_host (the world - 1 combinations of a split commitment, on the host): addition, doubling, identity operands,
    P + (-P), against the oracle's group law."""
    from uzkge_b200.dist import _jac_add_host

    FQ = 21888242871839275222246405745257275088696311157297823662689037894645226208583
    pts = oc.g1_random_points(4, 3)
    a = oc.g1_mul(pts[0], oc.random_fr(1, 1)[0])
    b = oc.g1_mul(pts[1], oc.random_fr(1, 2)[0])
    ident = np.zeros(12, dtype=np.uint64)
    for x, y in ((a, b), (a, a), (a, ident), (ident, b), (ident, ident)):
        assert np.array_equal(oc.g1_to_affine(_jac_add_host(x, y)), oc.g1_to_affine(oc.g1_add_jac(x, y)))
    neg = a.copy()
    yv = sum(int(a[4 + j]) << (64 * j) for j in range(4))
    for j in range(4):
        neg[4 + j] = (((FQ - yv) % FQ) >> (64 * j)) & 0xFFFFFFFFFFFFFFFF
    assert not oc.g1_to_affine(_jac_add_host(a, neg)).any()
