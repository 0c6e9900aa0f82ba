"""CPU: property tests (hypothesis) of the host-side mirrors that build circuits -- no GPU, no reference data: group laws of the
Baby Jubjub arithmetic in both restatements, permutations drawn from the RNG mirror, range checks over random values and widths,
Anemoi sponge / stream-cipher gadgets over random inputs.  Every generated witness must satisfy the restated constraint check."""
import pytest
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from plonk_circuits import FR

SCALARS = st.integers(min_value=1, max_value=(1 << 250) - 1)
FIELD = st.integers(min_value=0, max_value=FR - 1)
FAST = settings(max_examples=15, deadline=None, suppress_health_check=[HealthCheck.too_slow])


@FAST
@given(a=SCALARS, b=SCALARS)
def test_baby_jubjub_group_laws(a, b):
    from oracle import babyjubjub as bj
    from uzkge_b200 import shuffle as sh

    pa, pb = sh.ed_mul(a, sh.GENERATOR), sh.ed_mul(b, sh.GENERATOR)
    assert pa == bj.mul(a, bj.GEN) and sh.ed_is_on_curve(pa) and bj.on_curve(pb)
    assert sh.ed_add(pa, pb) == sh.ed_add(pb, pa) == bj.add(pa, pb) == sh.ed_mul(a + b, sh.GENERATOR)
    assert sh.ed_add(pa, sh.ed_neg(pa)) == sh.IDENTITY and sh.ed_add(pa, sh.IDENTITY) == pa
    assert sh.ed_mul(b, pa) == sh.ed_mul(a * b % sh.SUBGROUP_ORDER, sh.GENERATOR)
    msg = pb
    card = sh.Ciphertext(sh.ed_mul(b, sh.GENERATOR), sh.ed_add(msg, sh.ed_mul(b, pa)))      # encrypted to the key a G
    assert card.verify(msg, a) and not card.verify(sh.ed_add(msg, sh.GENERATOR), a)


@FAST
@given(seed=st.binary(min_size=32, max_size=32), n=st.integers(min_value=1, max_value=60))
def test_permutation_rand_is_a_permutation(seed, n):
    from uzkge_b200 import shuffle as sh
    from uzkge_b200.rng import ChaChaRng

    p = sh.Permutation.rand(ChaChaRng.from_seed(seed), n)
    assert len(p) == n
    p.sanity_check()
    bits = sh.BabyJubjubShuffle.sample_random_scalar_bits(ChaChaRng.from_seed(seed))
    assert len(bits) == 84 and all(isinstance(b, bool) for row in bits for b in row)


@FAST
@given(n_bits=st.integers(min_value=2, max_value=40), data=st.data())
def test_range_check_accepts_exactly_the_range(n_bits, data):
    from uzkge_b200 import plonk
    from uzkge_b200.errors import UzkgeError

    value = data.draw(st.integers(min_value=0, max_value=(1 << n_bits) - 1))
    cs = plonk.TurboCS()
    bits = cs.range_check(cs.new_variable(value), n_bits)
    cs.pad()
    cs.verify_witness(cs.witness, [])
    assert sum(cs.witness[b] << i for i, b in enumerate(bits)) == value
    cs = plonk.TurboCS()
    cs.range_check(cs.new_variable(value + (1 << n_bits)), n_bits)
    cs.pad()
    with pytest.raises(UzkgeError):
        cs.verify_witness(cs.witness, [])


@settings(max_examples=6, deadline=None, suppress_health_check=[HealthCheck.too_slow])
@given(values=st.lists(FIELD, min_size=1, max_size=7), out_len=st.integers(min_value=1, max_value=8))
def test_anemoi_gadgets_over_random_inputs(values, out_len):
    from uzkge_b200 import plonk
    from uzkge_b200.anemoi import AnemoiJive254 as A

    trace = A.eval_variable_length_hash_with_trace(values)
    assert trace.output == A.eval_variable_length_hash(values)
    cs = plonk.TurboCS()
    cs.load_anemoi_parameters()
    cs.anemoi_variable_length_hash(trace, [cs.new_variable(v) for v in values], cs.new_variable(trace.output))
    sc = A.eval_stream_cipher_with_trace(values, out_len)
    assert sc.output == A.eval_stream_cipher(values, out_len) and len(sc.output) == out_len
    cs.anemoi_stream_cipher(sc, [cs.new_variable(v) for v in values], [cs.new_variable(v) for v in sc.output])
    cs.pad()
    cs.verify_witness(cs.witness, [])
