"""GPU: ONE proof on a device group behind the C ABI (uzkge_cuda_plonk_params_upload_multi + uzkge_cuda_plonk_prove, csrc/prover.cu).
Every member runs the prover on replicated polynomials, commits only its slice of the SRS (partial sums added on the host) and
evaluates the quotient only on its cosets of the quotient domain (exchanged over peer memory).  The proof must be byte-identical to
the single-device proof -- and through it to the Python mirror and the big-integer restatement (tests/test_gpu_prover_native.py).

On a one-GPU box the group is VIRTUAL (uzkge_cuda_configure("virtual_devices", G)): G members, each with its own SRS slice, parameter
copy, streams and worker thread, all on device 0 -- the same code paths, including the peer copies.  On a box with several GPUs the
same tests also run over the real devices."""
import numpy as np
import pytest

from plonk_circuits import build_circuit, build_shuffle_circuit, shuffle_inputs

pytestmark = pytest.mark.gpu

TAU = 0x1234567890ABCDEF1234567890ABCDEF


def _group(gpu, members: int) -> int:
    """members > 0: a virtual group of that size; 0: the real devices of the box."""
    gpu.configure("virtual_devices", members)
    return gpu.init_devices(0)


def _prove_both(cs, params, pcs, lagrange, lagrange_all, label=b"group", seed=bytes(32)):
    from uzkge_b200.native import NativeProver
    from uzkge_b200.rng import ChaChaRng
    from uzkge_b200.transcript import Transcript

    wit = cs.get_witness_array()
    out = []
    for multi in (False, True):
        native = NativeProver(cs, params, pcs, lagrange, lagrange_all, multi=multi)
        try:
            tr = Transcript(label)
            proof = native.prove(ChaChaRng.from_seed(seed), tr, wit)
            again = native.prove(ChaChaRng.from_seed(seed), Transcript(label), wit)     # buffers and barriers are reused
            assert again.to_bytes_be() == proof.to_bytes_be()
            out.append((proof.to_bytes_be(), bytes(tr.state), dict(native.last_stats)))
        finally:
            native.close()
    return out


@pytest.mark.parametrize("deal", [False, True])
@pytest.mark.parametrize("members", [0, 2, 3, 8])
def test_group_proof_equals_single_device_proof(gpu, members, deal):
    """members = 0: the real GPUs of the box (a group of one on a one-GPU box, NVLink peers under gpurun --gpus N).
    deal: the large-circuit path forced onto these small circuits -- a round's interpolations and the linearisation polynomial are
    dealt to the members, who pull each other's results over peer memory (default: from 2^19 gates on)."""
    from uzkge_b200 import KZGCommitmentSchemeBN254, plonk

    try:
        gpu.configure("group_deal_min_log_n", 0 if deal else 19)
        assert _group(gpu, members) == (members or gpu.device_count())
        for n_gates, n_public, n_boolean in ((100, 3, 2), (900, 0, 0)):
            cs = build_circuit(plonk.TurboCS(), n_gates, 40 + n_gates, n_public, n_boolean)
            pcs = KZGCommitmentSchemeBN254.new(cs.size + 2, plonk.mont(TAU))
            lagrange = KZGCommitmentSchemeBN254.new_lagrange(cs.size, plonk.mont(TAU))
            params = plonk.indexer(cs, pcs)
            for lag, la in ((None, None), (lagrange, None), (lagrange, True)):
                (one, st1, _), (grp, st2, stats) = _prove_both(cs, params, pcs, lag, la)
                assert grp == one and st1 == st2, (members, n_gates, la)
                assert stats["msm"] == 13 and stats["coset_ifft_m"] == 0 and stats["coset_fft_m"] == 0    # no transform of the whole 6n domain
            pcs.close()
            lagrange.close()
    finally:
        gpu.configure("group_deal_min_log_n", 19)
        _group(gpu, 0)


@pytest.mark.parametrize("deal", [False, True])
def test_group_proof_of_the_shuffle_feature_set(gpu, deal):
    """zshuffle's circuit (2 cards) with its 1632-byte proof format: witness selectors, quotient terms 12-18, all-Lagrange route."""
    from uzkge_b200 import KZGCommitmentSchemeBN254, plonk

    try:
        gpu.configure("group_deal_min_log_n", 0 if deal else 19)
        assert _group(gpu, 4) == 4
        inp = shuffle_inputs(2, 77)
        cs, _ = build_shuffle_circuit(plonk.TurboCS(), inp)
        pcs = KZGCommitmentSchemeBN254.new(cs.size + 2, plonk.mont(TAU))
        lagrange = KZGCommitmentSchemeBN254.new_lagrange(cs.size, plonk.mont(TAU))
        params = plonk.indexer(cs, pcs, shuffle=True)
        plonk.refresh_prover_params_public_key(cs, params, pcs, inp["pk"])
        for lag, la in ((None, None), (lagrange, True)):
            (one, st1, _), (grp, st2, _) = _prove_both(cs, params, pcs, lag, la, label=b"shuffle")
            assert len(one) == 1632 and grp == one and st1 == st2
        pcs.close()
        lagrange.close()
    finally:
        gpu.configure("group_deal_min_log_n", 19)
        _group(gpu, 0)


def test_group_parameters_take_a_new_public_key(gpu):
    """uzkge_cuda_plonk_params_set_public_key on a multi-device handle: parameters uploaded to the group BEFORE the joint key was loaded
    prove, after the refresh, the bytes of a single-device prover built after it (shuffle/src/gen_params/params.rs:57-129)."""
    from uzkge_b200 import KZGCommitmentSchemeBN254, plonk
    from uzkge_b200.native import NativeProver
    from uzkge_b200.rng import ChaChaRng
    from uzkge_b200.transcript import Transcript

    try:
        assert _group(gpu, 3) == 3
        inp = shuffle_inputs(1, 9)
        cs, _ = build_shuffle_circuit(plonk.TurboCS(), inp)
        pcs = KZGCommitmentSchemeBN254.new(cs.size + 2, plonk.mont(TAU))
        params = plonk.indexer(cs, pcs, shuffle=True)
        early = NativeProver(cs, params, pcs, multi=True)
        plonk.refresh_prover_params_public_key(cs, params, pcs, inp["pk"])
        early.refresh_public_key()
        late = NativeProver(cs, params, pcs)
        wit = cs.get_witness_array()
        a = early.prove(ChaChaRng.from_seed(bytes(32)), Transcript(b"k"), wit).to_bytes_be()
        b = late.prove(ChaChaRng.from_seed(bytes(32)), Transcript(b"k"), wit).to_bytes_be()
        assert a == b and len(a) == 1632
        early.close()
        late.close()
        pcs.close()
    finally:
        _group(gpu, 0)


def test_group_prover_rejects_what_it_cannot_run(gpu):
    from uzkge_b200 import KZGCommitmentSchemeBN254, plonk
    from uzkge_b200.native import NativeProver
    from uzkge_b200.rng import ChaChaRng
    from uzkge_b200.transcript import Transcript

    try:
        assert _group(gpu, 2) == 2
        cs = build_circuit(plonk.TurboCS(), 50, 5, 1, 1)
        pcs = KZGCommitmentSchemeBN254.new(cs.size + 2, plonk.mont(TAU))
        params = plonk.indexer(cs, pcs)
        native = NativeProver(cs, params, pcs, multi=True)
        try:
            native.srs_handle, keep = pcs.handle, native.srs_handle          # a single-device SRS under a group handle
            with pytest.raises(Exception):
                native.prove(ChaChaRng.from_seed(bytes(32)), Transcript(b"x"), cs.get_witness_array())
            native.srs_handle = keep
            # an unsatisfied witness: the reference's DegreeError from every member, no deadlock
            bad = cs.get_witness_array().copy()
            bad[5, 0] ^= np.uint64(1)
            with pytest.raises(Exception):
                native.prove(ChaChaRng.from_seed(bytes(32)), Transcript(b"x"), bad)
            good = native.prove(ChaChaRng.from_seed(bytes(32)), Transcript(b"x"), cs.get_witness_array())
            assert len(good.to_bytes_be()) == 1312
        finally:
            native.close()
        pcs.close()
    finally:
        _group(gpu, 0)
