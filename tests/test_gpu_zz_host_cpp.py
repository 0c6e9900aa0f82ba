"""GPU: the reference's boundary tests (test_fft / check_fft, test_commit, test_homomorphic_poly_com_elem, test_pcs_eval,
apply_blind_factors, the Lagrange SRS) restated in C++ on include/uzkge_host.hpp, every expected value from the CPU oracle."""
import pytest

from host_cpp import run

pytestmark = pytest.mark.gpu


def test_cpp_host_layer_matches_the_oracle(gpu):
    out = run("gpu")
    assert out.returncode == 0 and "PASS gpu" in out.stdout, out.stdout + out.stderr


def test_cpp_host_proves_zshuffle_through_the_abi(gpu, tmp_path):
    """A COMPILED host (tests/host/host_api_test.cpp, no Python in the loop) calls uzkge_cuda_plonk_params_upload /
    uzkge_cuda_srs_upload_lagrange_commit / uzkge_cuda_plonk_prove with the data the Rust PlonkProverParams holds, for a zshuffle
    circuit (remark + permutation gadgets, `shuffle` feature set, all-Lagrange route) and prints the 1632-byte proof; it must equal
    the proof of the call-by-call mirror (uzkge_b200/plonk.py), which the restated verifier accepts."""
    import struct

    import numpy as np

    from plonk_circuits import build_shuffle_circuit, shuffle_inputs
    from uzkge_b200 import KZGCommitmentSchemeBN254, plonk
    from uzkge_b200.rng import ChaChaRng, fr_rand_mont
    from uzkge_b200.transcript import Transcript, transcript_init_plonk

    TAU = 0x1234567890ABCDEF1234567890ABCDEF
    inp = shuffle_inputs(3, 21)
    cs, _ = build_shuffle_circuit(plonk.TurboCS(), inp)
    n = cs.size
    pcs = KZGCommitmentSchemeBN254.new(n + 2, plonk.mont(TAU))
    lagrange = KZGCommitmentSchemeBN254.new_lagrange(n, plonk.mont(TAU))
    params = plonk.indexer(cs, pcs, shuffle=True, lagrange_pcs=lagrange)
    plonk.refresh_prover_params_public_key(cs, params, pcs, inp["pk"], lagrange_pcs=lagrange)
    wit = cs.get_witness_array()

    def transcript():
        tr = Transcript(b"Plonk shuffle Proof")
        tr.append_u64(3)
        return tr

    want = plonk.prover(ChaChaRng.from_seed(bytes(32)), transcript(), pcs, cs, params, wit, lagrange_pcs=lagrange, lagrange_all=True).to_bytes_be()

    recs = []

    def rec(tag, arr):
        b = np.ascontiguousarray(arr).tobytes() if not isinstance(arr, (bytes, bytearray)) else bytes(arr)
        recs.append(struct.pack("<QQ", tag, len(b)) + b)

    (SIZES, WIRING, PERM, K, Q, S, QB, PRK, ANEMOI, PUB_ROWS, PUB_WIT, Q_ECC, GEN, PK, EDWARDS, WITNESS, W_SEL, BLINDS, TRANSCRIPT, SRS, LAGRANGE,
     FLAGS) = range(1, 23)
    vp = params.verifier_params
    rec(SIZES, np.array([n, params.m, cs.num_vars, 1], dtype=np.uint64))
    rec(WIRING, cs.wiring.reshape(-1).astype(np.uint32))
    rec(PERM, cs.compute_permutation().astype(np.uint64))
    rec(K, plonk.mont_rows(vp.k))
    for p_ in params.q_polys:
        rec(Q, p_.numpy(n))
    for p_ in params.s_polys:
        rec(S, p_.numpy(n))
    rec(QB, params.qb_poly.numpy(n))
    for p_ in params.q_prk_polys:
        rec(PRK, p_.numpy(n))
    rec(ANEMOI, plonk.mont_rows([vp.anemoi_generator, vp.anemoi_generator_inv]))
    rec(PUB_ROWS, np.asarray(vp.public_vars_constraint_indices, dtype=np.uint64))
    rec(PUB_WIT, np.asarray(cs.public_vars_witness_indices, dtype=np.uint64))
    rec(Q_ECC, params.q_ecc_poly.numpy(n))
    for p_ in params.q_shuffle_generator_polys:
        rec(GEN, p_.numpy(n))
    for p_ in params.q_shuffle_public_key_polys:
        rec(PK, p_.numpy(n))
    rec(EDWARDS, plonk.mont(vp.edwards_a))
    rec(WITNESS, wit)
    for ev in cs.compute_witness_selectors():
        rec(W_SEL, ev)
    prng = ChaChaRng.from_seed(bytes(32))
    blinds = np.zeros((27, 4), dtype=np.uint64)
    for j in range(27):
        raw = fr_rand_mont(prng)
        blinds[j] = [(raw >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)]
    rec(BLINDS, blinds)
    tr = transcript()
    transcript_init_plonk(tr, vp, [plonk.unmont(wit[i]) for i in cs.public_vars_witness_indices], params.root)
    rec(TRANSCRIPT, bytes(tr.state))
    rec(SRS, pcs.public_parameter_group_1)
    rec(LAGRANGE, lagrange.public_parameter_group_1)
    rec(FLAGS, np.array([1], dtype=np.uint64))
    path = tmp_path / "zshuffle_job.bin"
    path.write_bytes(b"".join(recs))
    for p_ in (pcs, lagrange):
        p_.close()

    out = run("prove", str(path))
    assert out.returncode == 0 and "PASS prove" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
    got = [line.split()[1] for line in out.stdout.splitlines() if line.startswith("proof ")][0]
    assert len(want) == 1632 and bytes.fromhex(got) == want
