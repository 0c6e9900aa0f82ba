"""A/B of the quotient kernel's register budget on random coset evaluations (every term active) and inside a proof (zero selectors)."""
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch
from uzkge_b200 import ffi, plonk, KZGCommitmentSchemeBN254
from uzkge_b200.rng import ChaChaRng
from uzkge_b200.transcript import Transcript
import bench as B
ffi.init(0)
dev = torch.device("cuda", 0)
def timeit(fn, reps=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def d(a): return torch.from_numpy(np.ascontiguousarray(a).view(np.int64).reshape(-1)).to(dev)
n, factor = 1 << 20, 6
m = n * factor
base = B.random_fr(m, 1)
arrs = [d(np.roll(base, 7 * i, axis=0)) for i in range(28)]
sc = B.random_fr(8, 2); k = B.random_fr(5, 3); zh = B.random_fr(factor, 4)
out = torch.empty(4 * m, dtype=torch.int64, device=dev)
p = [t.data_ptr() for t in arrs]
fn = lambda: ffi.plonk_quotient_fr_device(p[0:5], p[5:14], p[14], p[15], p[16:21], p[21], p[22], p[23], p[24:28], k, sc[0], sc[1], sc[2], sc[3], sc[4], zh, m, factor, out.data_ptr())
for mb in (1, 4, 1, 4):
    ffi.configure("quotient_min_blocks", mb)
    print(f"random inputs, min_blocks={mb}: {timeit(fn)*1e3:9.1f} us", flush=True)
del arrs, out
cs = plonk.TurboCS.synthetic(20)
pcs = KZGCommitmentSchemeBN254.new(cs.size + 2, plonk.mont(12345))
params = plonk.indexer(cs, pcs)
wit = plonk.DevVec.from_numpy(cs.get_witness_array(), dev)
for mb in (1, 4, 1, 4):
    ffi.configure("quotient_min_blocks", mb)
    t = {}
    for _ in range(3):
        plonk.prover(ChaChaRng.from_seed(bytes(32)), Transcript(b"bench"), pcs, cs, params, wit, timings=t)
    print(f"proof 2^20, min_blocks={mb}: round3_quotient {t['round3_quotient']/3:.2f} ms", flush=True)
