"""GPU: the distributed (four-step) NTT with the ranks emulated on one GPU: same kernels (cross step + local
transforms through the C ABI), the all-to-alls done by indexing.  Bit-exact against the oracle's single transform."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64).reshape(-1)).cuda()


def to_np(t):
    return t.cpu().numpy().view(np.uint64).reshape(-1, 4)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("log_n", [6, 12, 17])
def test_distributed_ntt_matches_oracle(gpu, oc, world, log_n):
    from uzkge_b200 import dist as udist

    n = 1 << log_n
    L = n // world
    x = oc.random_fr(n, 40 + log_n + world)
    xs = [to_dev(x[r * L : (r + 1) * L]) for r in range(world)]
    want = oc.ntt_fr(x, n)
    ys = udist.ntt_fr_distributed_emulated(xs, n)
    got = np.concatenate([to_np(y) for y in ys])
    assert np.array_equal(got, want)
    # cyclic layout: rank r holds X[r + world * k2]
    yc = udist.ntt_fr_distributed_emulated(xs, n, natural_output=False)
    for r in range(world):
        assert np.array_equal(to_np(yc[r]), want[r::world])
    # inverse brings the natural slices back
    back = udist.ntt_fr_distributed_emulated(ys, n, inverse=True)
    assert np.array_equal(np.concatenate([to_np(b) for b in back]), x)
    assert np.array_equal(np.concatenate([to_np(b) for b in back]), oc.ntt_fr(want, n, inverse=True))


def test_distributed_ntt_full_size_round_trip(gpu, oc):
    from uzkge_b200 import dist as udist

    n, world = 1 << 22, 8
    L = n // world
    x = oc.random_fr(n, 3)
    xs = [to_dev(x[r * L : (r + 1) * L]) for r in range(world)]
    ys = udist.ntt_fr_distributed_emulated(xs, n)
    single = gpu.ntt_fr(x, n)
    assert np.array_equal(np.concatenate([to_np(y) for y in ys]), single)
    back = udist.ntt_fr_distributed_emulated(ys, n, inverse=True)
    assert np.array_equal(np.concatenate([to_np(b) for b in back]), x)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("log_n", [6, 13, 18])
def test_peer_fused_cross_step_matches_oracle(gpu, oc, world, log_n):
    """uzkge_cuda_ntt_cross_rows_fr_device: the cross-rank step with one base pointer per row, which is how dist.PeerNtt fuses both
    exchanges into the kernel's loads and stores over peer memory.  The `world` ranks are emulated on one GPU: every "rank" launches
    the kernel with in_rows = every rank's slice + its column offset, out_rows = every rank's receive buffer + that offset."""
    n = 1 << log_n
    L, S, log_g = n // world, n // world // world, world.bit_length() - 1
    x = oc.random_fr(n, 60 + log_n + world)
    want = oc.ntt_fr(x, n)
    for inverse in (False, True):
        src = want if inverse else x
        xs = [to_dev(src[r * L:(r + 1) * L]) for r in range(world)]
        rows = [torch.zeros(4 * L, dtype=torch.int64, device="cuda") for _ in range(world)]
        for r in range(world):
            off = 32 * r * S
            gpu.ntt_cross_rows_fr_device([t.data_ptr() + off for t in xs], [t.data_ptr() + off for t in rows], log_g, S, r * S, n, inverse)
        outs = []
        for r in range(world):
            y, scr = torch.empty_like(rows[r]), torch.empty_like(rows[r])
            gpu.ntt_fr_device(rows[r].data_ptr(), y.data_ptr(), scr.data_ptr(), L, L, inverse, None)
            outs.append(to_np(y))
        ref = oc.ntt_fr(src, n, inverse=inverse)
        for r in range(world):                       # cyclic layout: rank r holds X[r + world * k2]
            assert np.array_equal(outs[r], ref[r::world]), (inverse, r)


@pytest.mark.parametrize("world", [2, 4, 8])
@pytest.mark.parametrize("log_n", [6, 13, 18, 22])
def test_natural_output_folded_into_the_final_store(gpu, oc, world, log_n):
    """uzkge_cuda_ntt_fr_scatter_device: the local transform of the four-step whose last pass writes every output into its OWNER's
    natural slice (dist.PeerNtt.transform_natural does this over peer memory: no third exchange).  Ranks emulated on one GPU:
    the "peer" slices are buffers of the same device.  Forward and inverse, against the oracle's single transform."""
    if log_n == 22 and world != 8:
        pytest.skip("the full size once")
    n = 1 << log_n
    L, S, log_g = n // world, n // world // world, world.bit_length() - 1
    x = oc.random_fr(n, 80 + log_n + world)
    want = oc.ntt_fr(x, n) if log_n < 20 else gpu.ntt_fr(x, n)
    for inverse in (False, True):
        src = want if inverse else x
        ref = x if inverse else want
        xs = [to_dev(src[r * L:(r + 1) * L]) for r in range(world)]
        rows = [torch.zeros(4 * L, dtype=torch.int64, device="cuda") for _ in range(world)]
        for r in range(world):
            off = 32 * r * S
            gpu.ntt_cross_rows_fr_device([t.data_ptr() + off for t in xs], [t.data_ptr() + off for t in rows], log_g, S, r * S, n, inverse)
        nat = [torch.zeros(4 * L, dtype=torch.int64, device="cuda") for _ in range(world)]
        scr = torch.empty(4 * L, dtype=torch.int64, device="cuda")
        for r in range(world):
            keep = rows[r].clone()
            gpu.ntt_fr_scatter_device(rows[r].data_ptr(), [t.data_ptr() for t in nat], scr.data_ptr(), L, inverse, log_g, r)
            assert torch.equal(rows[r], keep)                # the input is left intact
        got = np.concatenate([to_np(t) for t in nat])
        assert np.array_equal(got, ref), inverse


def test_ipc_export_of_library_buffers(gpu):
    """Buffers from uzkge_cuda_dev_alloc can be exported for other processes (dist.PeerNtt maps them on the peer GPUs; opening
    a handle needs a second process, covered by scripts/dist_check.py under torchrun)."""
    p = gpu.dev_alloc(1 << 20)
    try:
        h = gpu.ipc_export(p)
        assert len(h) == 64 and any(h)
    finally:
        gpu.dev_free(p)
