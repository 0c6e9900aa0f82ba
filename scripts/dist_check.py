"""Multi-GPU parity check, one process per GPU (NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py
Checks the distributed four-step NTT (NCCL all-to-all) and the point-split MSM against single-GPU results, and times them."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, torch.distributed as dist
from uzkge_b200 import ffi, dist as udist
import bench as B

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); lr = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dev = torch.device("cuda", lr)
dist.init_process_group("nccl", device_id=dev)
ffi.init(lr)
ok = True

def sync_time(fn, reps=5, warm=2):
    for _ in range(warm): fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); dist.barrier(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())

for lg in (16, 22, 24, 26):
    n = 1 << lg
    L = n // world
    x = B.random_fr(n, 7) if lg <= 24 else None      # same seed on every rank
    if x is not None:
        mine = torch.from_numpy(np.ascontiguousarray(x[rank * L:(rank + 1) * L]).view(np.int64).reshape(-1)).to(dev)
    else:
        mine = torch.from_numpy(B.random_fr(L, 100 + rank).view(np.int64).reshape(-1)).to(dev)
    y = udist.ntt_fr_distributed(mine, n, rank, world)
    back = udist.ntt_fr_distributed(y, n, rank, world, inverse=True)
    good = bool(torch.equal(back, mine))
    if x is not None and lg <= 24:
        full = torch.from_numpy(x.view(np.int64).reshape(-1)).to(dev)
        o = torch.empty_like(full); s = torch.empty_like(full)
        ffi.ntt_fr_device(full.data_ptr(), o.data_ptr(), s.data_ptr(), n, n)
        torch.cuda.synchronize()
        good = good and bool(torch.equal(o[rank * L * 4:(rank + 1) * L * 4], y))
        t1 = sync_time(lambda: ffi.ntt_fr_device(full.data_ptr(), o.data_ptr(), s.data_ptr(), n, n))
        del full, o, s
    else:
        t1 = float("nan")
    # both exchanges fused into the cross kernel over peer memory (cudaIpc / NVLink P2P)
    peer = udist.PeerNtt(n, rank, world, dev)
    peer.x_view.copy_(mine)
    yp = peer.transform()
    yc = udist.ntt_fr_distributed(mine, n, rank, world, natural_output=False)
    good_peer = bool(torch.equal(yp, yc))
    peer.x_view.copy_(yc)      # the inverse of the cyclic output is NOT the input (layouts differ): check the inverse kernel path
    ypi = peer.transform(inverse=True).clone()
    yci = udist.ntt_fr_distributed(yc, n, rank, world, inverse=True, natural_output=False)
    good_peer = good_peer and bool(torch.equal(ypi, yci))
    peer.x_view.copy_(mine)
    tp = sync_time(lambda: peer.transform())
    peer.close()
    good = good and good_peer
    tn = sync_time(lambda: udist.ntt_fr_distributed(mine, n, rank, world))
    tc = sync_time(lambda: udist.ntt_fr_distributed(mine, n, rank, world, natural_output=False))
    ok &= good
    if rank == 0:
        print(f"dist ntt 2^{lg} world={world}: parity={'ok' if good else 'FAIL'}  single-GPU {t1*1e3:.0f} us  distributed {tn*1e3:.0f} us (cyclic out {tc*1e3:.0f} us, peer-memory fused {tp*1e3:.0f} us, parity={'ok' if good_peer else 'FAIL'})", flush=True)

# MSM: points split per rank, 96-byte partial sums gathered and added
n = 1 << 20
tau = B.random_fr(1, 5)[0]
bases = ffi.srs_generate(tau, n)           # every rank builds the same SRS (setup), keeps only its slice resident
sc = B.random_fr(n, 6)
srs = udist.ShardedSrs(bases, rank, world)
got = udist.msm_sharded(srs, sc, device=dev)
if rank == 0:
    h = ffi.srs_upload(bases)
    want = ffi.msm_g1(h, sc)
    same = np.array_equal(ffi.g1_to_affine(got), ffi.g1_to_affine(want))
    ok &= same
    print(f"sharded msm 2^20 world={world}: parity={'ok' if same else 'FAIL'}", flush=True)
flag = torch.tensor([1 if ok else 0], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
dist.destroy_process_group()
sys.exit(0 if int(flag.item()) == 1 else 1)
