"""torchrun --nproc-per-node G scripts/dist_plonk_check.py [log_n]: a proof whose commitments are point-split over G GPUs
(dist.SplitCommitter) must be identical to the single-GPU proof; prints both timings."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np, torch, torch.distributed as dist
from uzkge_b200 import ffi, plonk, KZGCommitmentSchemeBN254
from uzkge_b200 import dist as udist
from uzkge_b200.rng import ChaChaRng
from uzkge_b200.transcript import Transcript

rank, world, lrank = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lrank)
dev = torch.device("cuda", lrank)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
dist.init_process_group("nccl", device_id=dev)
ffi.init(lrank)
for lg in [int(a) for a in sys.argv[1:]] or [12, 16, 20]:
    n = 1 << lg
    tau = plonk.mont(0x1234567890ABCDEF1234567890ABCDEF)
    bases = ffi.srs_generate(tau, n + 3)
    sc = udist.SplitCommitter(bases, rank, world, device=dev)
    if rank != 0:
        served = sc.serve()
        sc.close()
        dist.barrier()
        continue
    cs = plonk.TurboCS.synthetic(lg)
    pcs = KZGCommitmentSchemeBN254(bases)
    wit = plonk.DevVec.from_numpy(cs.get_witness_array(), dev)
    res = {}
    for name, p in (("single", pcs), ("split", sc)):
        params = plonk.indexer(cs, p)
        for _ in range(2):
            proof = plonk.prover(ChaChaRng.from_seed(bytes(32)), Transcript(b"x"), p, cs, params, wit)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            proof = plonk.prover(ChaChaRng.from_seed(bytes(32)), Transcript(b"x"), p, cs, params, wit)
        torch.cuda.synchronize()
        res[name] = (proof, (time.perf_counter() - t0) / reps, params.verifier_params)
        del params
    a, b = res["single"][0], res["split"][0]
    same = (a.cm_w_vec == b.cm_w_vec and a.cm_t_vec == b.cm_t_vec and a.cm_z == b.cm_z and a.opening_witness_zeta == b.opening_witness_zeta
            and a.opening_witness_zeta_omega == b.opening_witness_zeta_omega and a.w_polys_eval_zeta == b.w_polys_eval_zeta
            and res["single"][2].cm_q_vec == res["split"][2].cm_q_vec and res["single"][2].cm_s_vec == res["split"][2].cm_s_vec)
    print(f"{'PASS' if same else 'FAIL'} plonk split 2^{lg} over {world} GPUs: single {res['single'][1]*1e3:.1f} ms, split {res['split'][1]*1e3:.1f} ms", flush=True)
    sc.shutdown()
    sc.close(); pcs.close()
    dist.barrier()
dist.destroy_process_group()
