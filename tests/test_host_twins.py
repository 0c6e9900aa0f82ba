"""CPU: the product's device headers (ff.cuh, ec.cuh, ntt_plan.h) compiled for the host (tests/host/uzhost.cpp)
against the oracle: pins the exact limb-level algorithms and the NTT pass decomposition the kernels run."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "host", "uzhost.cpp")
OUT = os.path.join(HERE, "host", "libuzhost.so")
CSRC = os.path.join(HERE, "..", "uzkge_b200", "csrc")


@pytest.fixture(scope="module")
def uz():
    deps = [SRC] + [os.path.join(CSRC, f) for f in ("ff.cuh", "ec.cuh", "ntt_plan.h")]
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in deps):
        subprocess.check_call(["g++", "-O2", "-shared", "-fPIC", "-o", OUT, SRC])
    return C.CDLL(OUT)


def p(a):
    return a.ctypes.data_as(C.c_void_p)


EDGE = None


def edge_values(oc, bn, mod):
    vals = [0, 1, 2, mod - 1, mod - 2, (1 << 256) % mod, (1 << 255) % mod, (mod + 1) // 2, 0xFFFFFFFF, 1 << 32, (1 << 64) - 1]
    return np.array([bn.int_to_limbs(v) for v in vals], dtype=np.uint64)


@pytest.mark.parametrize("field", ["fr", "fq"])
def test_field_ops_match_oracle(uz, oc, bn, field):
    mod = bn.FR if field == "fr" else bn.FQ
    e = edge_values(oc, bn, mod)
    rnd = oc.random_fr(2000, 3)  # < r < q: valid for both fields
    a = np.concatenate([np.repeat(e, len(e), axis=0), rnd])
    b = np.concatenate([np.tile(e, (len(e), 1)), rnd[::-1]])
    n = a.shape[0]
    o = np.empty_like(a)
    getattr(uz, f"uzhost_{field}_mul")(p(a), p(b), p(o), C.c_size_t(n))
    want = oc.fr_mul(a, b) if field == "fr" else oc.fq_mul(a, b)
    assert np.array_equal(o, want)
    ai, bi = bn.array_to_ints(a, None, mont=False), bn.array_to_ints(b, None, mont=False)
    getattr(uz, f"uzhost_{field}_add")(p(a), p(b), p(o), C.c_size_t(n))
    assert bn.array_to_ints(o, None, mont=False) == [(x + y) % mod for x, y in zip(ai, bi)]
    getattr(uz, f"uzhost_{field}_sub")(p(a), p(b), p(o), C.c_size_t(n))
    assert bn.array_to_ints(o, None, mont=False) == [(x - y) % mod for x, y in zip(ai, bi)]


@pytest.mark.parametrize("field", ["fr", "fq"])
def test_dedicated_squaring_matches_the_product(uz, oc, bn, field):
    """fe_sqr (ff.cuh: 108 multiplier instructions, separate reduction with deferred carries) == fe_mul(a, a) == the oracle's product, on
    edge values, on limb patterns that maximise the cross products and the carries (all-ones limbs below the modulus), and on random
    values."""
    mod = bn.FR if field == "fr" else bn.FQ
    pats = []
    for k in range(8):
        for fill in (0xFFFFFFFF, 0x80000000, 0xFFFF0000, 1):
            v = sum(fill << (32 * j) for j in range(k + 1))
            pats += [v % mod, (mod - v) % mod, (v << (32 * (7 - k))) % mod]
    pats += [mod - 1 - (1 << (32 * j)) for j in range(8)] + [(mod >> (32 * j)) << (32 * j) for j in range(1, 8)]
    a = np.concatenate([edge_values(oc, bn, mod), np.array([bn.int_to_limbs(v) for v in pats], dtype=np.uint64), oc.random_fr(20000, 31)])
    n = a.shape[0]
    o = np.empty_like(a)
    getattr(uz, f"uzhost_{field}_sqr")(p(a), p(o), C.c_size_t(n))
    want = oc.fr_mul(a, a) if field == "fr" else oc.fq_mul(a, a)
    assert np.array_equal(o, want)
    o2 = np.empty_like(a)
    getattr(uz, f"uzhost_{field}_mul")(p(a), p(a), p(o2), C.c_size_t(n))
    assert np.array_equal(o, o2)


def test_constants_and_inverse(uz, oc, bn):
    out = np.zeros((4, 4), dtype=np.uint64)
    uz.uzhost_consts(p(out))
    R = 1 << 256
    assert bn.limbs_to_int(out[0]) == R % bn.FQ and bn.limbs_to_int(out[1]) == R * R % bn.FQ
    assert bn.limbs_to_int(out[2]) == R % bn.FR and bn.limbs_to_int(out[3]) == R * R % bn.FR
    a = oc.random_fr(20, 5)
    o = np.empty_like(a)
    uz.uzhost_fr_inv(p(a), p(o), C.c_size_t(20))
    one = bn.ints_to_array([1], bn.FR)
    assert np.array_equal(oc.fr_mul(a, o), np.repeat(one, 20, axis=0))


def test_xyzz_group_law_matches_oracle(uz, oc, bn):
    n = 50
    pts = oc.g1_random_points(n, 9)
    # exercise the special cases: identity inputs, a repeated point (doubling), P then -P (cancellation)
    pts[3] = 0
    pts[10] = pts[9]
    pts[21] = pts[20]
    neg = np.zeros(n, dtype=np.int32)
    neg[21] = 1
    neg[30:40] = 1
    out = np.zeros(12, dtype=np.uint64)
    uz.uzhost_madd_chain(p(pts), p(neg), C.c_size_t(n), p(out))
    P = bn.array_to_affine(pts)
    acc = None
    for i in range(n):
        q = bn.g1_neg(P[i]) if neg[i] else P[i]
        acc = bn.g1_add(acc, q)
    assert bn.jac_array_to_affine(out) == acc
    # first two points only: P + P path from the identity accumulator
    two = np.stack([pts[0], pts[0]])
    uz.uzhost_madd_chain(p(two), p(np.zeros(2, dtype=np.int32)), C.c_size_t(2), p(out))
    assert bn.jac_array_to_affine(out) == bn.g1_add(P[0], P[0])

    s, ks, aff = np.zeros(12, dtype=np.uint64), np.zeros(12, dtype=np.uint64), np.zeros(8, dtype=np.uint64)
    uz.uzhost_add_tree(p(pts), C.c_size_t(n), C.c_uint32(1234567), p(s), p(ks), p(aff))
    tot = None
    for q in P:
        tot = bn.g1_add(tot, q)
    assert bn.jac_array_to_affine(s) == tot
    assert bn.jac_array_to_affine(ks) == bn.g1_mul(tot, 1234567)
    assert bn.array_to_affine(aff.reshape(1, 8))[0] == tot


def test_root_of_unity_twin_matches_golden(uz, bn, domain_kat):
    for n in (1, 2, 3, 6, 4096, 8192, 16384, 49152, 98304, 1 << 22, 1 << 28, 3 << 23):
        out = np.zeros(4, dtype=np.uint64)
        ok = C.c_int(0)
        uz.uzhost_root_of_unity(C.c_uint64(n), p(out), C.byref(ok))
        assert ok.value == 1
        assert bn.array_to_ints(out.reshape(1, 4), bn.FR)[0] == bn.root_of_unity(n)
    for n in (0, 5, 27, 1 << 29):
        ok = C.c_int(1)
        uz.uzhost_root_of_unity(C.c_uint64(n), p(np.zeros(4, dtype=np.uint64)), C.byref(ok))
        assert ok.value == 0


CONFIGS = [(12, 11, 22), (6, 5, 8), (4, 4, 6), (8, 3, 4)]


@pytest.mark.parametrize("cfg", CONFIGS)
def test_ntt_plan_emulation_matches_oracle(uz, oc, cfg):
    """Every pass structure (1, 2, 3 passes; mixed radix; every fused scaling) on small sizes: shrinking the tile
    limits makes small transforms take the multi-pass code paths that 2^22..2^24 take on the GPU."""
    log_tile, max_r, two_max = cfg
    k = oc.random_fr(1, 77)[0]
    seen = set()
    sizes = [1, 2, 3, 4, 6, 8, 12, 16, 24, 48, 64, 96, 128, 192, 256, 384, 512, 768, 1024, 1536, 2048, 4096, 3 << 11]
    for n in sizes:
        for inverse in (0, 1):
            for coset in (None, k):
                for len_in in sorted({n, n // 2 + 1}):
                    x = oc.random_fr(n, n + 7)
                    buf = np.zeros((n, 4), dtype=np.uint64)
                    buf[:len_in] = x[:len_in]
                    npass = C.c_uint32(0)
                    rc = uz.uzhost_ntt(p(buf), C.c_uint64(len_in), C.c_uint64(n), C.c_int(inverse),
                                       p(coset) if coset is not None else None, C.c_uint32(log_tile), C.c_uint32(max_r),
                                       C.c_uint32(two_max), C.byref(npass))
                    if rc == 1:
                        continue  # size too large for this (shrunken) configuration
                    seen.add(npass.value)
                    want = oc.ntt_fr(x[:len_in], n, bool(inverse), coset)
                    assert np.array_equal(buf, want), (cfg, n, inverse, coset is not None, len_in, npass.value)
    if cfg != (12, 11, 22):
        assert {1, 2, 3} <= seen


def test_public_inputs_leave_montgomery_form_in_one_c_call(bn, oc):
    """uzkge_host_fr_mont_to_be (csrc/hostutil.c): Montgomery limbs -> the canonical big-endian strings that enter the transcript
    (plonk/transcript.rs:27-30), against Python integers, edge values included."""
    from uzkge_b200.transcript import fr_mont_rows_to_bytes_be

    xs = [0, 1, bn.FR - 1, (1 << 253) + 17] + bn.array_to_ints(oc.random_fr(500, 12), bn.FR)
    got = fr_mont_rows_to_bytes_be(bn.ints_to_array(xs, bn.FR))
    assert got == b"".join(x.to_bytes(32, "big") for x in xs)
    assert fr_mont_rows_to_bytes_be(np.zeros((0, 4), dtype=np.uint64)) == b""
