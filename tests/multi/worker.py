"""Worker of tests/test_gpu_multi.py, one process per GPU (NCCL):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/multi/worker.py OUT.json
Every check compares the multi-GPU path with the CPU oracle or with the single-GPU result; rank 0 writes the verdicts to OUT.json
and the exit code is non-zero when any of them failed."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import torch.distributed as dist

from oracle import cpu as oc          # checker only
from uzkge_b200 import KZGCommitmentSchemeBN254, ffi, plonk
from uzkge_b200 import dist as udist
from uzkge_b200.rng import ChaChaRng
from uzkge_b200.transcript import Transcript

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
ffi.init(lrank)
oc.set_num_threads(max(1, len(os.sched_getaffinity(0)) // world))
checks = {}


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64).reshape(-1)).to(dev)


def to_np(t):
    return t.cpu().numpy().view(np.uint64).reshape(-1, 4)


def agree(flag: bool) -> bool:
    t = torch.tensor([1 if flag else 0], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return bool(t.item())


# 1. the four-step transform over the ranks (NCCL all-to-all, and both exchanges fused over peer memory) against the oracle's single
#    transform: BASELINE's 2^22 and a small size
for lg in (16, 22):
    n = 1 << lg
    L = n // world
    x = oc.random_fr(n, 700 + lg)                 # same seed on every rank
    want = oc.ntt_fr(x, n)
    mine = to_dev(x[rank * L:(rank + 1) * L])
    nat = udist.ntt_fr_distributed(mine, n, rank, world)
    ok = np.array_equal(to_np(nat), want[rank * L:(rank + 1) * L])
    cyc = udist.ntt_fr_distributed(mine, n, rank, world, natural_output=False)
    ok = ok and np.array_equal(to_np(cyc), want[rank::world])
    back = udist.ntt_fr_distributed(nat, n, rank, world, inverse=True)
    ok = ok and bool(torch.equal(back, mine))
    checks[f"ntt_four_step_nccl_2^{lg}_vs_oracle"] = agree(ok)
    peer = udist.PeerNtt(n, rank, world, dev)
    peer.x_view.copy_(mine)
    okp = np.array_equal(to_np(peer.transform()), want[rank::world])
    checks[f"ntt_four_step_peer_memory_2^{lg}_vs_oracle"] = agree(okp)
    # natural output folded into the local transform's final store over peer memory (no third exchange), forward and back
    okn = np.array_equal(to_np(peer.transform_natural()), want[rank * L:(rank + 1) * L])
    peer.x_view.copy_(peer.y_view)
    okn = okn and bool(torch.equal(peer.transform_natural(inverse=True), mine))
    peer.close()
    checks[f"ntt_four_step_peer_memory_natural_output_2^{lg}_vs_oracle"] = agree(okn)

# 2. the point-split MSM against the oracle (2^16) and against the single-GPU MSM (2^20)
for lg in (16, 20):
    n = 1 << lg
    tau = oc.random_fr(1, 55)[0]
    bases = ffi.srs_generate(tau, n)
    sc = oc.random_fr(n, 56 + lg)
    srs = udist.ShardedSrs(bases, rank, world)
    got = udist.msm_sharded(srs, sc, device=dev)
    if lg == 16:
        ok = np.array_equal(oc.g1_to_affine(got), oc.g1_to_affine(oc.msm_g1(bases, sc)))
        checks["msm_point_split_2^16_vs_oracle"] = agree(ok)
    else:
        ok = np.array_equal(oc.g1_to_affine(got), oc.g1_to_affine(oc.g1_mul(bases[0], oc.fr_eval(sc, tau))))
        checks["msm_point_split_2^20_vs_trapdoor"] = agree(ok)
    ffi.srs_free(srs.handle)

# 3. the independent commitments of a round dealt to the ranks (SRS replicated), every MSM on its rank's GPU
n = 1 << 14
tau = oc.random_fr(1, 77)[0]
bases = ffi.srs_generate(tau, n)
h = ffi.srs_upload(bases)
polys = [oc.random_fr(n - 7 * j, 800 + j) for j in range(8)]
got = udist.commit_distributed(polys, rank, world, lambda p: ffi.msm_g1(h, p), device=dev)
ok = all(np.array_equal(oc.g1_to_affine(got[j]), oc.g1_to_affine(oc.g1_mul(bases[0], oc.fr_eval(polys[j], tau)))) for j in range(8))
checks["round_commitments_dealt_to_ranks_vs_trapdoor"] = agree(ok)
ffi.srs_free(h)

# 4. one proof with every commitment point-split over the ranks == the single-GPU proof, byte for byte
for lg in (12, 16):
    n = 1 << lg
    tau_m = plonk.mont(0x1234567890ABCDEF1234567890ABCDEF)
    bases = ffi.srs_generate(tau_m, n + 3)
    sc = udist.SplitCommitter(bases, rank, world, device=dev)
    if rank != 0:
        sc.serve()
        sc.close()
    else:
        cs = plonk.TurboCS.synthetic(lg)
        pcs = KZGCommitmentSchemeBN254(bases)
        wit = plonk.DevVec.from_numpy(cs.get_witness_array(), dev)
        proofs = []
        for p in (pcs, sc):
            params = plonk.indexer(cs, p)
            proofs.append(plonk.prover(ChaChaRng.from_seed(bytes(32)), Transcript(b"x"), p, cs, params, wit).to_bytes_be())
            del params
        checks[f"split_proof_2^{lg}_equals_single_gpu_proof"] = proofs[0] == proofs[1] and len(proofs[0]) > 0
        sc.shutdown()
        sc.close()
        pcs.close()
    dist.barrier()

if rank == 0:
    checks["world"] = world
    with open(sys.argv[1], "w") as f:
        json.dump(checks, f, indent=1)
    print(json.dumps(checks), flush=True)
flag = agree(all(v for k, v in checks.items() if k != "world"))
dist.destroy_process_group()
sys.exit(0 if flag else 1)
