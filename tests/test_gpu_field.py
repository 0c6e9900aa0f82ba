"""GPU: K1 field kernels through the C ABI, bit-exact against the oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("field", ["fr", "fq"])
def test_field_mul_matches_oracle(gpu, oc, bn, field):
    mod = bn.FR if field == "fr" else bn.FQ
    vals = [0, 1, 2, mod - 1, mod - 2, (1 << 256) % mod, (1 << 255) % mod, (mod + 1) // 2, 0xFFFFFFFF, 1 << 32, (1 << 64) - 1]
    e = np.array([bn.int_to_limbs(v) for v in vals], dtype=np.uint64)
    rnd = oc.random_fr(100000, 3)
    a = np.concatenate([np.repeat(e, len(e), axis=0), rnd])
    b = np.concatenate([np.tile(e, (len(e), 1)), rnd[::-1]])
    want = oc.fr_mul(a, b) if field == "fr" else oc.fq_mul(a, b)
    assert np.array_equal(gpu.field_mul(a, b, field), want)


def test_field_mul_empty(gpu):
    z = np.zeros((0, 4), dtype=np.uint64)
    assert gpu.field_mul(z, z).shape == (0, 4)
