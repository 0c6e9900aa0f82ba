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
