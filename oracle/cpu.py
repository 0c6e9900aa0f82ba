"""ctypes front-end of the C oracle (oracle/oracle.c).  TEST INFRASTRUCTURE ONLY.

Arrays are numpy uint64: Fr/Fq elements (n, 4), affine points (n, 8), Jacobian (12,), all
little-endian limbs in Montgomery form -- the same layout as the CUDA C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(_build.build())
        u64p = C.POINTER(C.c_uint64)
        sz = C.c_size_t
        L.oracle_init.restype = None
        for name in ("oracle_fr_mul", "oracle_fq_mul"):
            getattr(L, name).argtypes = [u64p, u64p, u64p, sz]
            getattr(L, name).restype = None
        for name in ("oracle_fr_to_mont", "oracle_fr_from_mont", "oracle_fq_to_mont", "oracle_fq_from_mont"):
            getattr(L, name).argtypes = [u64p, u64p, sz]
            getattr(L, name).restype = None
        L.oracle_msm_g1.argtypes = [u64p, u64p, sz, u64p]
        L.oracle_msm_g1.restype = C.c_int
        L.oracle_g1_mul.argtypes = [u64p, u64p, u64p]
        L.oracle_g1_mul.restype = None
        L.oracle_g1_add_jac.argtypes = [u64p, u64p, u64p]
        L.oracle_g1_add_jac.restype = None
        L.oracle_g1_to_affine.argtypes = [u64p, u64p]
        L.oracle_g1_to_affine.restype = None
        L.oracle_g1_on_curve.argtypes = [u64p]
        L.oracle_g1_on_curve.restype = C.c_int
        L.oracle_g1_random_points.argtypes = [C.c_uint64, sz, u64p]
        L.oracle_g1_random_points.restype = None
        L.oracle_g1_progression.argtypes = [u64p, u64p, sz, u64p]
        L.oracle_g1_progression.restype = None
        L.oracle_fr_root_of_unity.argtypes = [sz, u64p]
        L.oracle_fr_root_of_unity.restype = None
        L.oracle_ntt_fr.argtypes = [u64p, sz, sz, C.c_int, u64p]
        L.oracle_ntt_fr.restype = C.c_int
        L.oracle_fr_eval.argtypes = [u64p, sz, u64p, u64p]
        L.oracle_fr_eval.restype = None
        L.oracle_fr_pow_u64.argtypes = [u64p, C.c_uint64, u64p]
        L.oracle_fr_pow_u64.restype = None
        L.oracle_fr_inv.argtypes = [u64p, u64p]
        L.oracle_fr_inv.restype = None
        L.oracle_fr_weighted_sums.argtypes = [u64p, sz, u64p, u64p]
        L.oracle_fr_weighted_sums.restype = None
        L.oracle_num_threads.restype = C.c_int
        L.oracle_set_num_threads.argtypes = [C.c_int]
        L.oracle_set_num_threads.restype = None
        L.oracle_init()
        _lib = L
    return _lib


def _p(a: np.ndarray):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.POINTER(C.c_uint64))


def _c(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint64)


def set_num_threads(n: int) -> None:
    lib().oracle_set_num_threads(int(n))


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def _binop(name, a, b):
    a, b = _c(a).reshape(-1, 4), _c(b).reshape(-1, 4)
    assert a.shape == b.shape
    o = np.empty_like(a)
    getattr(lib(), name)(_p(a), _p(b), _p(o), a.shape[0])
    return o


def _unop(name, a):
    a = _c(a).reshape(-1, 4)
    o = np.empty_like(a)
    getattr(lib(), name)(_p(a), _p(o), a.shape[0])
    return o


def fr_mul(a, b):
    return _binop("oracle_fr_mul", a, b)


def fq_mul(a, b):
    return _binop("oracle_fq_mul", a, b)


def fr_to_mont(a):
    return _unop("oracle_fr_to_mont", a)


def fr_from_mont(a):
    return _unop("oracle_fr_from_mont", a)


def fq_to_mont(a):
    return _unop("oracle_fq_to_mont", a)


def fq_from_mont(a):
    return _unop("oracle_fq_from_mont", a)


def _ge_modulus(a: np.ndarray, m: int) -> np.ndarray:
    from .bn254 import int_to_limbs

    ml = np.array(int_to_limbs(m), dtype=np.uint64)
    ge = np.zeros(a.shape[0], dtype=bool)
    undecided = np.ones(a.shape[0], dtype=bool)
    for k in (3, 2, 1, 0):
        gt = undecided & (a[:, k] > ml[k])
        lt = undecided & (a[:, k] < ml[k])
        ge |= gt
        undecided &= ~(gt | lt)
    return ge | undecided


def random_fr(n: int, seed: int) -> np.ndarray:
    """n uniform Fr elements as raw limbs (254-bit draws, redrawn while >= r): like ``Fr::rand``, the
    limbs are used directly as the Montgomery representation."""
    from .bn254 import FR

    rng = np.random.default_rng(seed)

    def draw(k):
        a = rng.integers(0, 1 << 63, size=(k, 4), dtype=np.uint64) * np.uint64(2) + rng.integers(
            0, 2, size=(k, 4), dtype=np.uint64
        )
        a[:, 3] &= np.uint64((1 << 62) - 1)
        return a

    a = draw(n)
    while True:
        bad = np.nonzero(_ge_modulus(a, FR))[0]
        if bad.size == 0:
            return a
        a[bad] = draw(bad.size)


def msm_g1(bases, scalars) -> np.ndarray:
    bases, scalars = _c(bases).reshape(-1, 8), _c(scalars).reshape(-1, 4)
    assert bases.shape[0] == scalars.shape[0]
    out = np.zeros(12, dtype=np.uint64)
    rc = lib().oracle_msm_g1(_p(bases), _p(scalars), bases.shape[0], _p(out))
    if rc:
        raise MemoryError("oracle_msm_g1")
    return out


def g1_mul(base_aff, scalar) -> np.ndarray:
    out = np.zeros(12, dtype=np.uint64)
    lib().oracle_g1_mul(_p(_c(base_aff).reshape(8)), _p(_c(scalar).reshape(4)), _p(out))
    return out


def g1_add_jac(a, b) -> np.ndarray:
    out = np.zeros(12, dtype=np.uint64)
    lib().oracle_g1_add_jac(_p(_c(a).reshape(12)), _p(_c(b).reshape(12)), _p(out))
    return out


def g1_to_affine(jac) -> np.ndarray:
    out = np.zeros(8, dtype=np.uint64)
    lib().oracle_g1_to_affine(_p(_c(jac).reshape(12)), _p(out))
    return out


def g1_on_curve(aff) -> bool:
    return bool(lib().oracle_g1_on_curve(_p(_c(aff).reshape(8))))


def g1_random_points(n: int, seed: int) -> np.ndarray:
    out = np.zeros((n, 8), dtype=np.uint64)
    lib().oracle_g1_random_points(seed, n, _p(out))
    return out


def g1_progression(p0_aff, q_aff, n: int) -> np.ndarray:
    out = np.zeros((n, 8), dtype=np.uint64)
    lib().oracle_g1_progression(_p(_c(p0_aff).reshape(8)), _p(_c(q_aff).reshape(8)), n, _p(out))
    return out


def fr_root_of_unity(n: int) -> np.ndarray:
    out = np.zeros(4, dtype=np.uint64)
    lib().oracle_fr_root_of_unity(n, _p(out))
    return out


def ntt_fr(data, n: int, inverse: bool = False, coset=None) -> np.ndarray:
    """Returns a new (n, 4) array; ``data`` holds the first len_in elements."""
    data = _c(data).reshape(-1, 4)
    len_in = data.shape[0]
    buf = np.zeros((n, 4), dtype=np.uint64)
    buf[:len_in] = data
    cp = _p(_c(coset).reshape(4)) if coset is not None else None
    rc = lib().oracle_ntt_fr(_p(buf), len_in, n, 1 if inverse else 0, cp)
    if rc:
        raise ValueError(f"oracle_ntt_fr: bad size {len_in}/{n}")
    return buf


def fr_eval(coefs, x) -> np.ndarray:
    coefs = _c(coefs).reshape(-1, 4)
    out = np.zeros(4, dtype=np.uint64)
    lib().oracle_fr_eval(_p(coefs), coefs.shape[0], _p(_c(x).reshape(4)), _p(out))
    return out


def fr_pow(a, e: int) -> np.ndarray:
    out = np.zeros(4, dtype=np.uint64)
    lib().oracle_fr_pow_u64(_p(_c(a).reshape(4)), e, _p(out))
    return out


def fr_inv(a) -> np.ndarray:
    out = np.zeros(4, dtype=np.uint64)
    lib().oracle_fr_inv(_p(_c(a).reshape(4)), _p(out))
    return out


def fr_weighted_sums(s):
    s = _c(s).reshape(-1, 4)
    s0 = np.zeros(4, dtype=np.uint64)
    s1 = np.zeros(4, dtype=np.uint64)
    lib().oracle_fr_weighted_sums(_p(s), s.shape[0], _p(s0), _p(s1))
    return s0, s1
