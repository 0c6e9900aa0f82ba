"""ONE process, two GPUs with peer access: the four-step transform's cross kernel and scattering local transform reading / writing the
OTHER GPU's memory -- for `ncu --metrics nvlrx__bytes.sum,nvltx__bytes.sum,...` (a torchrun job cannot run under ncu).
    python scripts/gpu_peer_ntt_once.py LOG_N
Both "ranks" are driven from this process, one after the other (the kernels are the ones dist.PeerNtt launches per rank)."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import torch

import bench as B
from uzkge_b200 import ffi

lg = int(sys.argv[1]) if len(sys.argv) > 1 else 22
world, log_g = 2, 1
assert torch.cuda.device_count() >= 2, "needs two GPUs"
assert ffi.init_devices(2) == 2          # enables peer access both ways
n = 1 << lg
L, S = n // world, n // world // world
x = B.random_fr(n, 11)
devs = [torch.device("cuda", r) for r in range(world)]
xs = [torch.from_numpy(x[r * L:(r + 1) * L].view(np.int64).reshape(-1)).to(devs[r]) for r in range(world)]
rows = [torch.zeros(4 * L, dtype=torch.int64, device=devs[r]) for r in range(world)]
nat = [torch.zeros(4 * L, dtype=torch.int64, device=devs[r]) for r in range(world)]
scr = [torch.empty(4 * L, dtype=torch.int64, device=devs[r]) for r in range(world)]
for rep in range(2):
    for r in range(world):
        ffi.set_device(r)
        off = 32 * r * S
        ffi.ntt_cross_rows_fr_device([t.data_ptr() + off for t in xs], [t.data_ptr() + off for t in rows], log_g, S, r * S, n, False)
    for r in range(world):
        torch.cuda.synchronize(devs[r])
    for r in range(world):
        ffi.set_device(r)
        ffi.ntt_fr_scatter_device(rows[r].data_ptr(), [t.data_ptr() for t in nat], scr[r].data_ptr(), L, False, log_g, r)
    for r in range(world):
        torch.cuda.synchronize(devs[r])
got = np.concatenate([t.cpu().numpy().view(np.uint64).reshape(-1, 4) for t in nat])
ffi.set_device(0)
want = ffi.ntt_fr(x, n)
print("peer four-step 2^%d over 2 GPUs in one process: %s" % (lg, "ok" if np.array_equal(got, want) else "MISMATCH"))
sys.exit(0 if np.array_equal(got, want) else 1)
