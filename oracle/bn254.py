"""Pure-Python big-int restatement of the arithmetic behind uzkge's hot path.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Small sizes only; the C
restatement in oracle/oracle.c covers 2^16..2^24.

What is restated, and which reference call site it stands behind:

* Fq / Fr, 4 x u64 little-endian limbs in Montgomery form R = 2^256 -- the
  in-memory ``ark_ff::Fp`` that crosses the C ABI.
* BN254 G1 (y^2 = x^3 + 3, generator (1, 2)), Jacobian coordinates, identity Z = 0
  -- ``G1Projective`` in /root/reference/uzkge/src/poly_commit/kzg_poly_commitment.rs:278-293.
* ``msm``: sum_i s_i * P_i  -- ``G1Projective::msm`` at kzg_poly_commitment.rs:290.
* ``root_of_unity``: the arkworks ``FftField`` generator for N = 3^a * 2^b
  (used by both ``Radix2EvaluationDomain`` and ``MixedRadixEvaluationDomain``,
  /root/reference/uzkge/src/poly_commit/field_polynomial.rs:554-567); pinned by the
  generators in the generated verifier keys (tests/golden/domain_kat.json).
* ``fft`` / ``ifft`` / ``coset_fft`` / ``coset_ifft``: natural-order DFT with zero
  padding to the domain size -- field_polynomial.rs:570-607; the executable
  contract is ``check_fft`` (field_polynomial.rs:632-646): fft[i] == poly.eval(w^i).
"""
from __future__ import annotations

import numpy as np

# ---------------------------------------------------------------- fields (S1)
FQ = 21888242871839275222246405745257275088696311157297823662689037894645226208583
FR = 21888242871839275222246405745257275088548364400416034343698204186575808495617
R256 = 1 << 256
MASK64 = (1 << 64) - 1

FR_GENERATOR = 5          # multiplicative generator of Fr* (ark-bn254 FrConfig)
FR_TWO_ADICITY = 28
FR_SMALL_SUBGROUP_BASE = 3
FR_SMALL_SUBGROUP_ADICITY = 2

G1_B = 3
G1_GEN = (1, 2)


def inv_mod(a: int, m: int) -> int:
    return pow(a, -1, m)


def to_mont(x: int, m: int) -> int:
    return (x * R256) % m


def from_mont(x: int, m: int) -> int:
    return (x * inv_mod(R256, m)) % m


def int_to_limbs(x: int) -> list[int]:
    return [(x >> (64 * i)) & MASK64 for i in range(4)]


def limbs_to_int(l) -> int:
    return int(l[0]) | (int(l[1]) << 64) | (int(l[2]) << 128) | (int(l[3]) << 192)


def ints_to_array(xs, m: int | None = None, mont: bool = True) -> np.ndarray:
    """list of canonical ints -> (n, 4) uint64 array (Montgomery limbs if ``mont``)."""
    out = np.empty((len(xs), 4), dtype=np.uint64)
    for i, x in enumerate(xs):
        if mont:
            x = to_mont(x, m)
        out[i] = int_to_limbs(x)
    return out


def array_to_ints(a: np.ndarray, m: int | None = None, mont: bool = True) -> list[int]:
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 4)
    rinv = inv_mod(R256, m) if mont else 1
    out = []
    for row in a:
        x = limbs_to_int(row)
        if mont:
            x = (x * rinv) % m
        out.append(x)
    return out


# ---------------------------------------------------------------- curve (S2)
def g1_is_on_curve(P) -> bool:
    if P is None:
        return True
    x, y = P
    return (y * y - x * x * x - G1_B) % FQ == 0


def g1_neg(P):
    if P is None:
        return None
    return (P[0], (-P[1]) % FQ)


def g1_add(P, Q):
    """Affine addition (None = identity)."""
    if P is None:
        return Q
    if Q is None:
        return P
    x1, y1 = P
    x2, y2 = Q
    if x1 == x2:
        if (y1 + y2) % FQ == 0:
            return None
        lam = (3 * x1 * x1) * inv_mod(2 * y1, FQ) % FQ
    else:
        lam = (y2 - y1) * inv_mod(x2 - x1, FQ) % FQ
    x3 = (lam * lam - x1 - x2) % FQ
    y3 = (lam * (x1 - x3) - y1) % FQ
    return (x3, y3)


def jac_double(P):
    X, Y, Z = P
    if Z == 0:
        return (1, 1, 0)
    A = X * X % FQ
    B = Y * Y % FQ
    C = B * B % FQ
    D = 2 * ((X + B) * (X + B) - A - C) % FQ
    E = 3 * A % FQ
    F = E * E % FQ
    X3 = (F - 2 * D) % FQ
    Y3 = (E * (D - X3) - 8 * C) % FQ
    Z3 = 2 * Y * Z % FQ
    return (X3, Y3, Z3)


def jac_add(P, Q):
    X1, Y1, Z1 = P
    X2, Y2, Z2 = Q
    if Z1 == 0:
        return Q
    if Z2 == 0:
        return P
    Z1Z1 = Z1 * Z1 % FQ
    Z2Z2 = Z2 * Z2 % FQ
    U1 = X1 * Z2Z2 % FQ
    U2 = X2 * Z1Z1 % FQ
    S1 = Y1 * Z2 * Z2Z2 % FQ
    S2 = Y2 * Z1 * Z1Z1 % FQ
    if U1 == U2:
        if S1 == S2:
            return jac_double(P)
        return (1, 1, 0)
    H = (U2 - U1) % FQ
    Rr = (S2 - S1) % FQ
    HH = H * H % FQ
    HHH = H * HH % FQ
    V = U1 * HH % FQ
    X3 = (Rr * Rr - HHH - 2 * V) % FQ
    Y3 = (Rr * (V - X3) - S1 * HHH) % FQ
    Z3 = Z1 * Z2 * H % FQ
    return (X3, Y3, Z3)


def jac_from_affine(P):
    if P is None:
        return (1, 1, 0)
    return (P[0], P[1], 1)


def jac_to_affine(P):
    X, Y, Z = P
    if Z % FQ == 0:
        return None
    zi = inv_mod(Z, FQ)
    zi2 = zi * zi % FQ
    return (X * zi2 % FQ, Y * zi2 * zi % FQ)


def g1_mul(P, k: int):
    """Scalar multiplication, affine in / affine out (double-and-add on Jacobian)."""
    k %= FR
    acc = (1, 1, 0)
    base = jac_from_affine(P)
    while k:
        if k & 1:
            acc = jac_add(acc, base)
        base = jac_double(base)
        k >>= 1
    return jac_to_affine(acc)


def msm_naive(points, scalars):
    """sum_i scalars[i] * points[i]; the definition (test_commit, kzg_poly_commitment.rs:526-548)."""
    assert len(points) == len(scalars)
    acc = (1, 1, 0)
    for P, s in zip(points, scalars):
        if P is None or s % FR == 0:
            continue
        acc = jac_add(acc, jac_from_affine(g1_mul(P, s)))
    return jac_to_affine(acc)


def msm_window(n: int) -> int:
    """arkworks VariableBaseMSM window rule (ark-ec 0.4 variable_base/mod.rs, [memory])."""
    if n < 32:
        return 3
    return (int(np.floor(np.log2(n))) * 69) // 100 + 2


def msm_pippenger(points, scalars, c: int | None = None):
    """Bucket method with unsigned windows; same result as ``msm_naive`` (canonical output)."""
    assert len(points) == len(scalars)
    n = len(points)
    if n == 0:
        return None
    if c is None:
        c = msm_window(n)
    nwin = (254 + c - 1) // c
    total = (1, 1, 0)
    for w in reversed(range(nwin)):
        for _ in range(c):
            total = jac_double(total)
        buckets = [(1, 1, 0)] * ((1 << c) - 1)
        for P, s in zip(points, scalars):
            if P is None:
                continue
            d = ((s % FR) >> (w * c)) & ((1 << c) - 1)
            if d:
                buckets[d - 1] = jac_add(buckets[d - 1], jac_from_affine(P))
        run = (1, 1, 0)
        acc = (1, 1, 0)
        for b in reversed(buckets):
            run = jac_add(run, b)
            acc = jac_add(acc, run)
        total = jac_add(total, acc)
    return jac_to_affine(total)


# ------------------------------------------------------- point / scalar codecs
def affine_to_array(points) -> np.ndarray:
    """list of affine points (None = identity -> x = y = 0) -> (n, 8) uint64 Montgomery limbs."""
    out = np.zeros((len(points), 8), dtype=np.uint64)
    for i, P in enumerate(points):
        if P is None:
            continue
        out[i, :4] = int_to_limbs(to_mont(P[0], FQ))
        out[i, 4:] = int_to_limbs(to_mont(P[1], FQ))
    return out


def array_to_affine(a: np.ndarray):
    a = np.asarray(a, dtype=np.uint64).reshape(-1, 8)
    rinv = inv_mod(R256, FQ)
    out = []
    for row in a:
        x = limbs_to_int(row[:4])
        y = limbs_to_int(row[4:])
        if x == 0 and y == 0:
            out.append(None)
        else:
            out.append((x * rinv % FQ, y * rinv % FQ))
    return out


def jac_array_to_affine(a) -> tuple[int, int] | None:
    """12 x u64 Montgomery Jacobian (the C ABI's MSM output) -> canonical affine or None."""
    a = np.asarray(a, dtype=np.uint64).reshape(12)
    rinv = inv_mod(R256, FQ)
    X = limbs_to_int(a[0:4]) * rinv % FQ
    Y = limbs_to_int(a[4:8]) * rinv % FQ
    Z = limbs_to_int(a[8:12]) * rinv % FQ
    return jac_to_affine((X, Y, Z))


def parse_srs_g1(raw: bytes):
    """G1 part of an SRS file written by ``to_unchecked_bytes``
    (/root/reference/uzkge/src/poly_commit/kzg_poly_commitment.rs:207-228):
    u32 len1 | u32 len2 | len1 x (x LE 32 B | y LE 32 B, flag bits in the top two bits of byte 63)."""
    len1 = int.from_bytes(raw[0:4], "little")
    pts = []
    for i in range(len1):
        off = 8 + 64 * i
        x = int.from_bytes(raw[off:off + 32], "little")
        yb = bytearray(raw[off + 32:off + 64])
        flags = yb[31] & 0xC0
        yb[31] &= 0x3F
        y = int.from_bytes(bytes(yb), "little")
        if flags & 0x40:  # infinity
            pts.append(None)
        else:
            pts.append((x, y))
    return pts


# ---------------------------------------------------------------- domains (S4)
def _factor_domain(n: int) -> tuple[int, int]:
    """n = 3^a * 2^b with a <= 2, b <= 28; raises otherwise."""
    a = 0
    m = n
    while m % 3 == 0 and m > 0:
        m //= 3
        a += 1
    b = 0
    while m % 2 == 0 and m > 0:
        m //= 2
        b += 1
    if m != 1 or a > FR_SMALL_SUBGROUP_ADICITY or b > FR_TWO_ADICITY:
        raise ValueError(f"unsupported domain size {n}")
    return a, b


def large_subgroup_root() -> int:
    q = (FR - 1) // ((1 << FR_TWO_ADICITY) * FR_SMALL_SUBGROUP_BASE ** FR_SMALL_SUBGROUP_ADICITY)
    return pow(FR_GENERATOR, q, FR)


def root_of_unity(n: int) -> int:
    """``FftField::get_root_of_unity(n)`` for n = 3^a * 2^b (canonical integer)."""
    a, b = _factor_domain(n)
    e = FR_SMALL_SUBGROUP_BASE ** (FR_SMALL_SUBGROUP_ADICITY - a) * (1 << (FR_TWO_ADICITY - b))
    return pow(large_subgroup_root(), e, FR)


# ---------------------------------------------------------------- transforms (S5)
def poly_eval(coefs, x: int) -> int:
    acc = 0
    for c in reversed(coefs):
        acc = (acc * x + c) % FR
    return acc


def dft_naive(coefs, n: int, root: int | None = None):
    """O(n^2) definition: out[i] = sum_j c_j w^{ij}; input zero-padded to n."""
    w = root_of_unity(n) if root is None else root
    assert len(coefs) <= n
    return [poly_eval(coefs, pow(w, i, FR)) for i in range(n)]


def _ntt_rec(x, w):
    n = len(x)
    if n == 1:
        return x
    if n % 2 == 0:
        w2 = w * w % FR
        ev = _ntt_rec(x[0::2], w2)
        od = _ntt_rec(x[1::2], w2)
        out = [0] * n
        t = 1
        h = n // 2
        for k in range(h):
            v = t * od[k] % FR
            out[k] = (ev[k] + v) % FR
            out[k + h] = (ev[k] - v) % FR
            t = t * w % FR
        return out
    assert n % 3 == 0
    w3 = pow(w, 3, FR)
    s = [_ntt_rec(x[i::3], w3) for i in range(3)]
    m = n // 3
    out = [0] * n
    for k in range(n):
        wk = pow(w, k, FR)
        out[k] = (s[0][k % m] + wk * s[1][k % m] + wk * wk % FR * s[2][k % m]) % FR
    return out


def fft(coefs, n: int):
    """``domain.fft(&coefs)`` (field_polynomial.rs:583-586): zero-pad to n, natural order."""
    _factor_domain(n)
    assert len(coefs) <= n
    x = [c % FR for c in coefs] + [0] * (n - len(coefs))
    return _ntt_rec(x, root_of_unity(n))


def ifft(evals, n: int):
    """``domain.ifft(&values)`` (field_polynomial.rs:594-597) WITHOUT the trailing-zero trim."""
    _factor_domain(n)
    assert len(evals) <= n
    x = [c % FR for c in evals] + [0] * (n - len(evals))
    y = _ntt_rec(x, inv_mod(root_of_unity(n), FR))
    ninv = inv_mod(n, FR)
    return [v * ninv % FR for v in y]


def coset_fft(coefs, n: int, k: int):
    """``coset_fft_with_domain`` (field_polynomial.rs:589-591): c_j * k^j then fft."""
    scaled = []
    p = 1
    for c in coefs:
        scaled.append(c * p % FR)
        p = p * k % FR
    return fft(scaled, n)


def coset_ifft(evals, n: int, k_inv: int):
    """``coset_ifft_with_domain`` (field_polynomial.rs:601-607): ifft then c_j * k_inv^j."""
    c = ifft(evals, n)
    out = []
    p = 1
    for v in c:
        out.append(v * p % FR)
        p = p * k_inv % FR
    return out


def trim(coefs):
    """``FpPolynomial::from_coefs`` trailing-zero trim (field_polynomial.rs:86-90); zero poly = [0]."""
    c = list(coefs)
    while len(c) > 1 and c[-1] == 0:
        c.pop()
    return c if c else [0]
