"""GPU, two or more devices: the multi-GPU paths with REAL NCCL / peer memory, one process per GPU (tests/multi/worker.py under
torch.distributed.run) -- the four-step transform against the oracle at BASELINE's 2^22, the point-split MSM, the round's
commitments dealt to the ranks, and the split proof against the single-GPU proof.  Skipped on a one-GPU box (the emulated-rank
tests in test_gpu_dist_ntt.py and the gloo tests in test_dist_gloo.py still run there)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_multi_gpu_paths_against_oracle_and_single_gpu(gpu, tmp_path):
    count = gpu.device_count()
    if count < 2:
        pytest.skip("needs at least two GPUs (run under gpurun --gpus 2)")
    world = 2 if count < 4 else 4
    out = tmp_path / "multi.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "multi", "worker.py"), str(out)]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    assert res.returncode == 0, res.stdout[-2000:] + res.stderr[-4000:]
    checks = json.load(open(out))
    assert checks.pop("world") == world
    assert len(checks) >= 9 and all(checks.values()), checks


@pytest.mark.parametrize("mode", ["split", "replicated"])
def test_in_process_device_group_msm(gpu, oc, mode):
    """One process driving every GPU of the box through the C ABI (uzkge_cuda_init_devices + uzkge_cuda_srs_upload_multi): point-split
    MSMs with the partial sums combined on the host, and a round's independent commitments dealt to the devices, against the oracle.
    Runs on any number of devices (a group of one exercises the same fan-out and combine code)."""
    G = gpu.init_devices(0)
    assert G == gpu.device_count() >= 1
    n = 6000
    pts = oc.g1_random_points(n, 4100)
    pts[17] = 0                                       # an identity base (padded SRS)
    h = gpu.srs_upload_multi(pts, gpu.MULTI_SPLIT if mode == "split" else gpu.MULTI_REPLICATED)
    try:
        info = gpu.srs_info(h)
        assert info["n"] == n and info["reserved"] == G
        sc = oc.random_fr(n, 4101)
        assert np.array_equal(oc.g1_to_affine(gpu.msm_g1(h, sc)), oc.g1_to_affine(oc.msm_g1(pts, sc)))
        # a prefix, an offset range that straddles slice boundaries, an empty range
        m = n // 3 + 5
        assert np.array_equal(oc.g1_to_affine(gpu.msm_g1(h, sc[:m])), oc.g1_to_affine(oc.msm_g1(pts[:m], sc[:m])))
        off = n // 2 - 700
        assert np.array_equal(oc.g1_to_affine(gpu.msm_g1(h, sc[:1500], base_offset=off)), oc.g1_to_affine(oc.msm_g1(pts[off:off + 1500], sc[:1500])))
        assert not gpu.msm_g1(h, sc[:0])[8:].any()
        # a round's ragged batch (5 wire polynomials, z, ...): more jobs than devices and fewer
        for k in (1, 3, 8, 11):
            vecs = [oc.random_fr(n - 313 * j, 4200 + j) for j in range(k)]
            got = gpu.msm_g1_batch(h, vecs)
            for j in range(k):
                assert np.array_equal(oc.g1_to_affine(got[j]), oc.g1_to_affine(oc.msm_g1(pts[: n - 313 * j], vecs[j]))), (mode, k, j)
        with pytest.raises(Exception):
            gpu.msm_g1(h, oc.random_fr(n + 1, 1))
    finally:
        gpu.srs_free(h)
    with pytest.raises(Exception):
        gpu.msm_g1(h, sc)                              # the handle is gone


def test_device_calls_on_different_streams_share_workspaces_safely(gpu, oc):
    """Advisor finding (round 1): two *_device MSMs on ONE handle issued on two streams used to overwrite each other's sort / bucket
    buffers.  The library now orders the users of a shared workspace with events: both results must be right, every time; the same
    for the polynomial engine's scan workspace."""
    import torch

    n = 1 << 14
    pts = oc.g1_random_points(n, 5100)
    h = gpu.srs_upload(pts)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    try:
        sc = [oc.random_fr(n, 5101 + j) for j in range(2)]
        want = [oc.g1_to_affine(oc.msm_g1(pts, s)) for s in sc]
        d_sc = [torch.from_numpy(s.view(np.int64)).cuda() for s in sc]
        outs = [torch.zeros(12, dtype=torch.int64, device="cuda") for _ in range(2)]
        vals = [torch.zeros(4, dtype=torch.int64, device="cuda") for _ in range(2)]
        z = oc.random_fr(1, 5200)[0]
        want_ev = [oc.fr_eval(s, z) for s in sc]
        torch.cuda.synchronize()
        for _ in range(10):
            for j, st in enumerate((s1, s2)):
                gpu.msm_g1_device(h, d_sc[j].data_ptr(), n, outs[j].data_ptr(), st.cuda_stream)
                gpu.poly_horner_fr_device(d_sc[j].data_ptr(), n, z, 0, vals[j].data_ptr(), st.cuda_stream)
            torch.cuda.synchronize()
            for j in range(2):
                assert np.array_equal(oc.g1_to_affine(outs[j].cpu().numpy().view(np.uint64)), want[j])
                assert np.array_equal(vals[j].cpu().numpy().view(np.uint64), want_ev[j])
    finally:
        gpu.srs_free(h)
