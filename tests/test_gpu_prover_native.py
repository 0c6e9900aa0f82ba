"""GPU: the compiled prover behind the C ABI (uzkge_cuda_plonk_params_upload / uzkge_cuda_plonk_prove, csrc/prover.cu) against the
call-by-call mirror (uzkge_b200/plonk.py::prover) and against the big-integer restatement of the reference's prover
(oracle/plonk_prover.py: prover.rs:88-394): same circuit, SRS trapdoor, ChaCha seed and transcript label -> identical proof bytes, on
every feature set (default / shuffle), every commitment route (monomial, prover_with_lagrange, all-Lagrange, holed SRS) and on the
zshuffle and zmatchmaking circuits; an unsatisfied witness yields the reference's DegreeError, no proof."""
import numpy as np
import pytest

from plonk_circuits import build_circuit, build_shuffle_circuit, shuffle_inputs

pytestmark = pytest.mark.gpu

TAU = 0x1234567890ABCDEF1234567890ABCDEF


def _pair(cs, params, pcs, label, lagrange=None, lagrange_all=None, prefix=None, seed=bytes(32), wit=None):
    """(proof bytes of the Python mirror, proof bytes of the compiled prover, its statistics)."""
    from uzkge_b200 import plonk
    from uzkge_b200.native import NativeProver
    from uzkge_b200.rng import ChaChaRng
    from uzkge_b200.transcript import Transcript

    def transcript():
        tr = Transcript(label)
        if prefix is not None:
            tr.append_u64(prefix)
        return tr

    wit = cs.get_witness_array() if wit is None else wit
    params.workspace.pop("lagrange_scheme", None)
    params.workspace.pop("srs_truncated", None)
    kw = {} if lagrange_all is None else {"lagrange_all": lagrange_all}
    t1 = transcript()
    want = plonk.prover(ChaChaRng.from_seed(seed), t1, pcs, cs, params, wit, lagrange_pcs=lagrange, **kw)
    native = NativeProver(cs, params, pcs, lagrange, lagrange_all)
    try:
        t2 = transcript()
        got = native.prove(ChaChaRng.from_seed(seed), t2, wit)
        assert bytes(t2.state) == bytes(t1.state), "the transcripts diverged"
        # a second proof on the same handle (buffers are reused) and one from a witness already in HBM
        again = native.prove(ChaChaRng.from_seed(seed), transcript(), plonk.DevVec.from_numpy(wit, plonk._dev()))
        assert again.to_bytes_be() == got.to_bytes_be()
        stats = native.last_stats
    finally:
        native.close()
    return want.to_bytes_be(), got.to_bytes_be(), stats


@pytest.mark.parametrize("n_gates,n_public,n_boolean", [(25, 1, 1), (100, 3, 2), (200, 0, 0)])
def test_native_prover_matches_mirror_and_restatement(gpu, bn, n_gates, n_public, n_boolean):
    from oracle import plonk_prover as pp
    from uzkge_b200 import KZGCommitmentSchemeBN254, plonk

    seed = 11 + n_gates
    cs = build_circuit(plonk.TurboCS(), n_gates, seed, n_public, n_boolean)
    ocs = build_circuit(pp.TurboCS(), n_gates, seed, n_public, n_boolean)
    pcs, opcs = KZGCommitmentSchemeBN254.new(cs.size + 2, plonk.mont(TAU)), pp.Kzg(cs.size + 2, TAU)
    params, oparams = plonk.indexer(cs, pcs), pp.indexer(ocs, opcs)
    want, got, stats = _pair(cs, params, pcs, b"test")
    assert got == want
    ref = pp.prover(pp.ChaCha(bytes(32)), pp.Transcript(b"test"), opcs, ocs, oparams, ocs.witness)
    assert got == pp.proof_to_bytes_be(ref)
    assert stats["msm"] == 13 and stats["coset_ifft_m"] == 1 and stats["coset_fft_m"] == 6 + (1 if n_public else 0)
    # prover_with_lagrange: wires and z over the Lagrange SRS, then everything, then with the production files' holes in the SRS
    lagrange = KZGCommitmentSchemeBN254.new_lagrange(cs.size, plonk.mont(TAU))
    for la in (None, True):
        w2, g2, _ = _pair(cs, params, pcs, b"test", lagrange, la)
        assert g2 == w2 == want, la
    holes = pcs.public_parameter_group_1.copy()
    holes[3:cs.size] = 0
    sparse = KZGCommitmentSchemeBN254(holes)
    w3, g3, _ = _pair(cs, params, sparse, b"test", lagrange)
    assert g3 == w3 == want
    # other blinds, other proof; same bytes from both provers
    w4, g4, _ = _pair(cs, params, pcs, b"test", seed=bytes([7] * 32))
    assert g4 == w4 != want
    for p in (pcs, lagrange, sparse):
        p.close()


def test_native_prover_shuffle_feature_set_and_remark_gates(gpu, bn):
    """The 1632-byte proof format of zshuffle's deployed verifier: a circuit without remark gates (zero witness selectors) and the
    real remark + permutation gadgets of a 2-card deck with a joint key loaded, on the monomial and the all-Lagrange route."""
    from oracle import plonk_prover as pp
    from uzkge_b200 import KZGCommitmentSchemeBN254, plonk

    cs = build_circuit(plonk.TurboCS(), 60, 3, 2, 1)
    pcs = KZGCommitmentSchemeBN254.new(cs.size + 2, plonk.mont(TAU))
    params = plonk.indexer(cs, pcs, shuffle=True)
    want, got, stats = _pair(cs, params, pcs, b"sel")
    assert len(got) == 1632 and got == want and stats["msm"] == 16
    pcs.close()

    inp = shuffle_inputs(2, 5)
    cs, _ = build_shuffle_circuit(plonk.TurboCS(), inp)
    ocs, _ = build_shuffle_circuit(pp.TurboCS(), inp)
    n = cs.size
    pcs, opcs = KZGCommitmentSchemeBN254.new(n + 2, plonk.mont(TAU)), pp.Kzg(n + 2, TAU)
    lagrange = KZGCommitmentSchemeBN254.new_lagrange(n, plonk.mont(TAU))
    params, oparams = plonk.indexer(cs, pcs, shuffle=True), pp.indexer(ocs, opcs, shuffle=True)
    plonk.refresh_prover_params_public_key(cs, params, pcs, inp["pk"])
    pp.refresh_public_key(oparams, ocs, opcs, inp["pk"])
    otr = pp.Transcript(b"Plonk shuffle Proof")
    otr.u64(2)
    ref = pp.proof_to_bytes_be(pp.prover(pp.ChaCha(bytes(32)), otr, opcs, ocs, oparams, ocs.witness))
    for lag, la in ((None, None), (lagrange, None), (lagrange, True)):
        want, got, _ = _pair(cs, params, pcs, b"Plonk shuffle Proof", lag, la, prefix=2)
        assert got == want == ref, (lag is not None, la)
    for p in (pcs, lagrange):
        p.close()


def test_native_prover_refreshes_the_public_key(gpu):
    """uzkge_cuda_plonk_params_set_public_key: parameters uploaded BEFORE the joint key was loaded prove the same bytes after the
    refresh as parameters uploaded after it (refresh_prover_params_public_key, shuffle/src/gen_params/params.rs:57-129)."""
    from uzkge_b200 import KZGCommitmentSchemeBN254, plonk
    from uzkge_b200.native import NativeProver
    from uzkge_b200.rng import ChaChaRng
    from uzkge_b200.transcript import Transcript

    inp = shuffle_inputs(1, 9)
    cs, _ = build_shuffle_circuit(plonk.TurboCS(), inp)
    pcs = KZGCommitmentSchemeBN254.new(cs.size + 2, plonk.mont(TAU))
    params = plonk.indexer(cs, pcs, shuffle=True)
    early = NativeProver(cs, params, pcs)
    plonk.refresh_prover_params_public_key(cs, params, pcs, inp["pk"])
    early.refresh_public_key()
    late = NativeProver(cs, params, pcs)
    wit = cs.get_witness_array()
    a = early.prove(ChaChaRng.from_seed(bytes(32)), Transcript(b"k"), wit).to_bytes_be()
    b = late.prove(ChaChaRng.from_seed(bytes(32)), Transcript(b"k"), wit).to_bytes_be()
    c = plonk.prover(ChaChaRng.from_seed(bytes(32)), Transcript(b"k"), pcs, cs, params, wit).to_bytes_be()
    assert a == b == c
    early.close()
    late.close()
    pcs.close()


def test_native_prover_matchmaking_circuit(gpu):
    """Anemoi gates (quotient terms 8-11, the prk parts of the linearisation) through the compiled prover: a small zmatchmaking
    circuit, both feature sets, monomial and all-Lagrange routes."""
    import random

    from plonk_circuits import FR
    from uzkge_b200 import KZGCommitmentSchemeBN254, plonk
    from uzkge_b200 import matchmaking as mm

    rnd = random.Random(12)
    cs, _ = mm.build_cs(plonk.TurboCS(), [rnd.randrange(FR) for _ in range(3)], rnd.randrange(FR), rnd.randrange(FR))
    pcs = KZGCommitmentSchemeBN254.new(cs.size + 2, plonk.mont(TAU))
    lagrange = KZGCommitmentSchemeBN254.new_lagrange(cs.size, plonk.mont(TAU))
    for shuffle in (False, True):
        params = plonk.indexer(cs, pcs, shuffle=shuffle)
        for lag, la in ((None, None), (lagrange, True)):
            want, got, _ = _pair(cs, params, pcs, mm.PLONK_PROOF_TRANSCRIPT, lag, la, prefix=3)
            assert got == want, (shuffle, la)
    for p in (pcs, lagrange):
        p.close()


def test_native_prover_rejects_an_unsatisfied_witness(gpu):
    from uzkge_b200 import KZGCommitmentSchemeBN254, plonk
    from uzkge_b200.errors import UzkgeError
    from uzkge_b200.native import NativeProver
    from uzkge_b200.rng import ChaChaRng
    from uzkge_b200.transcript import Transcript

    cs = build_circuit(plonk.TurboCS(), 30, 5)
    pcs = KZGCommitmentSchemeBN254.new(cs.size + 2, plonk.mont(TAU))
    params = plonk.indexer(cs, pcs)
    native = NativeProver(cs, params, pcs)
    w = cs.get_witness_array().copy()
    w[7] = plonk.mont(plonk.unmont(w[7]) + 1)
    with pytest.raises(UzkgeError):
        native.prove(ChaChaRng.from_seed(bytes(32)), Transcript(b"test"), w)
    # the handle is still usable afterwards
    ok = native.prove(ChaChaRng.from_seed(bytes(32)), Transcript(b"test"), cs.get_witness_array())
    assert ok.to_bytes_be() == plonk.prover(ChaChaRng.from_seed(bytes(32)), Transcript(b"test"), pcs, cs, params, cs.get_witness_array()).to_bytes_be()
    native.close()
    pcs.close()
