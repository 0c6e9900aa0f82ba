"""Host-side mirror of the reference's `poly_commit` interface for the MSM / NTT hot path, on top of the C ABI.

No Rust toolchain exists in this image, so the layer that in production is the two patched Rust bodies
 is restated here in Python with the reference's names, argument meaning and error behaviour, so
that the parity tests read like the reference's own tests:

  FpPolynomial                     /root/reference/uzkge/src/poly_commit/field_polynomial.rs:13-17, 86-90, 554-607
  Radix2EvaluationDomain,
  MixedRadixEvaluationDomain       ark-poly domains as used at field_polynomial.rs:554-567 (size, group_gen, fft, ifft)
  KZGCommitmentSchemeBN254         /root/reference/uzkge/src/poly_commit/kzg_poly_commitment.rs:169-177, 268-313
  KZGCommitment                    kzg_poly_commitment.rs:24-53 (transcript bytes = affine x BE || y BE)

Field elements are numpy uint64 rows of 4 little-endian limbs in Montgomery form (arkworks' in-memory `Fr`).
Every transform and every MSM runs on the GPU through `uzkge_b200.ffi`; nothing here computes them on the CPU.
"""
from __future__ import annotations

import numpy as np

from . import ffi
from .errors import DegreeError, ParameterError

FR_MODULUS = 21888242871839275222246405745257275088548364400416034343698204186575808495617
FQ_MODULUS = 21888242871839275222246405745257275088696311157297823662689037894645226208583
_FQ_R = (1 << 256) % FQ_MODULUS
_FQ_R_INV = pow(_FQ_R, -1, FQ_MODULUS)


def _limbs_to_int(row) -> int:
    return int(row[0]) | (int(row[1]) << 64) | (int(row[2]) << 128) | (int(row[3]) << 192)


def _int_to_limbs(x: int) -> np.ndarray:
    return np.array([(x >> (64 * i)) & 0xFFFFFFFFFFFFFFFF for i in range(4)], dtype=np.uint64)


def _is_pow2(n: int) -> bool:
    return n > 0 and (n & (n - 1)) == 0


class _Domain:
    """What the hot path needs from an ark-poly EvaluationDomain: `size()`, `group_gen`, `fft`, `ifft`."""

    def __init__(self, num_coeffs: int):
        self._size = num_coeffs
        self.group_gen = ffi.fr_root_of_unity(num_coeffs)

    def size(self) -> int:
        return self._size

    def fft(self, coefs: np.ndarray) -> np.ndarray:
        """domain.fft(&coefs): zero-pads to size(), natural order in and out."""
        return ffi.ntt_fr(coefs, self._size, inverse=False)

    def ifft(self, evals: np.ndarray) -> np.ndarray:
        return ffi.ntt_fr(evals, self._size, inverse=True)


class Radix2EvaluationDomain(_Domain):
    @classmethod
    def new(cls, num_coeffs: int):
        """Radix2EvaluationDomain::new: the smallest power of two >= num_coeffs, None above 2^28."""
        size = 1
        while size < num_coeffs:
            size <<= 1
        if size > (1 << 28):
            return None
        return cls(size)


class MixedRadixEvaluationDomain(_Domain):
    @classmethod
    def new(cls, num_coeffs: int):
        """Sizes the reference asks for: 2^k or 3 * 2^k (field_polynomial.rs:561-567)."""
        if not (_is_pow2(num_coeffs) or (num_coeffs % 3 == 0 and _is_pow2(num_coeffs // 3))):
            return None
        if num_coeffs > 3 * (1 << 28):
            return None
        return cls(num_coeffs)


class FpPolynomial:
    """Coefficient vector, low order first, trailing zeros trimmed (field_polynomial.rs:13-17, 86-90, 154-159)."""

    def __init__(self, coefs: np.ndarray):
        self.coefs = coefs

    # ---- construction
    @classmethod
    def from_coefs(cls, coefs) -> "FpPolynomial":
        c = ffi.as_u64(coefs, 4)
        p = cls(c)
        p.trim_coefs()
        return p

    @classmethod
    def zero(cls) -> "FpPolynomial":
        return cls(np.zeros((1, 4), dtype=np.uint64))

    def trim_coefs(self) -> None:
        nz = np.nonzero(self.coefs.any(axis=1))[0]
        keep = int(nz[-1]) + 1 if nz.size else 1
        if self.coefs.shape[0] == 0:
            self.coefs = np.zeros((1, 4), dtype=np.uint64)
        else:
            self.coefs = self.coefs[:keep]

    def get_coefs_ref(self) -> np.ndarray:
        return self.coefs

    def degree(self) -> int:
        return self.coefs.shape[0] - 1

    def is_zero(self) -> bool:
        return self.degree() == 0 and not self.coefs[0].any()

    def __eq__(self, other) -> bool:
        return isinstance(other, FpPolynomial) and np.array_equal(self.coefs, other.coefs)

    # ---- evaluation and division by (X - z): the serial loops of field_polynomial.rs:198-209 and :519-550 as GPU scans
    def eval(self, point) -> np.ndarray:
        return ffi.poly_eval_fr(self.coefs, point)

    def div_rem_linear(self, z) -> tuple["FpPolynomial", "FpPolynomial"]:
        """self.div_rem(&FpPolynomial::from_coefs(vec![-z, 1])): (quotient, remainder), both trimmed like the reference."""
        if self.coefs.shape[0] < 2:
            return FpPolynomial.zero(), FpPolynomial.from_coefs(self.coefs.copy())
        q, r = ffi.poly_div_linear_fr(self.coefs, z)
        return FpPolynomial.from_coefs(q), FpPolynomial.from_coefs(r.reshape(1, 4))

    # ---- domains (field_polynomial.rs:554-567)
    @staticmethod
    def evaluation_domain(num_coeffs: int):
        assert _is_pow2(num_coeffs)
        return Radix2EvaluationDomain.new(num_coeffs)

    @staticmethod
    def quotient_evaluation_domain(num_coeffs: int):
        assert _is_pow2(num_coeffs) or (num_coeffs % 3 == 0 and _is_pow2(num_coeffs // 3))
        return MixedRadixEvaluationDomain.new(num_coeffs)

    # ---- transforms (field_polynomial.rs:570-607)
    def fft(self, num_coeffs: int):
        assert num_coeffs > self.degree()
        if _is_pow2(num_coeffs):
            domain = self.evaluation_domain(num_coeffs)
        else:
            domain = self.quotient_evaluation_domain(num_coeffs)
        if domain is None:
            return None
        return self.fft_with_domain(domain)

    def fft_with_domain(self, domain: _Domain) -> np.ndarray:
        assert domain.size() > self.degree()
        return domain.fft(self.coefs)

    def coset_fft_with_domain(self, domain: _Domain, k) -> np.ndarray:
        """self.mul_var(k).fft_with_domain(domain): the power scaling is fused into the transform's first read."""
        assert domain.size() > self.degree()
        return ffi.ntt_fr(self.coefs, domain.size(), inverse=False, coset_shift=k)

    @classmethod
    def ifft_with_domain(cls, domain: _Domain, values) -> "FpPolynomial":
        return cls.from_coefs(domain.ifft(ffi.as_u64(values, 4)))

    @classmethod
    def coset_ifft_with_domain(cls, domain: _Domain, values, k_inv) -> "FpPolynomial":
        """ifft_with_domain(domain, values).mul_var(k_inv): the scaling is fused into the transform's last store."""
        return cls.from_coefs(ffi.ntt_fr(ffi.as_u64(values, 4), domain.size(), inverse=True, coset_shift=k_inv))


class KZGCommitment:
    """KZGCommitment(G1Projective): Jacobian X, Y, Z Montgomery limbs as returned by the backend."""

    def __init__(self, jac: np.ndarray):
        self.value = ffi.as_u64(jac).reshape(12)
        self._affine = None   # canonical (x, y) integers, or () for the identity

    def _affine_ints(self):
        """`into_affine()` + leaving Montgomery form, as the reference does on the host when it serialises a commitment
        (kzg_poly_commitment.rs:37-53): one modular inversion, O(1) per commitment."""
        if self._affine is None:
            x, y, z = (_limbs_to_int(self.value[4 * i: 4 * i + 4]) * _FQ_R_INV % FQ_MODULUS for i in range(3))
            if z == 0:
                self._affine = ()
            else:
                zi = pow(z, -1, FQ_MODULUS)
                zi2 = zi * zi % FQ_MODULUS
                self._affine = (x * zi2 % FQ_MODULUS, y * zi2 % FQ_MODULUS * zi % FQ_MODULUS)
        return self._affine

    def to_affine(self) -> np.ndarray:
        """(x, y) Montgomery limbs, zeros for the identity."""
        a = self._affine_ints()
        if not a:
            return np.zeros(8, dtype=np.uint64)
        return np.concatenate([_int_to_limbs(a[0] * _FQ_R % FQ_MODULUS), _int_to_limbs(a[1] * _FQ_R % FQ_MODULUS)])

    def is_identity(self) -> bool:
        return not self.value[8:].any()

    def add(self, other: "KZGCommitment") -> "KZGCommitment":
        return KZGCommitment(ffi.g1_add(self.value, other.value))

    def to_transcript_bytes(self) -> bytes:
        """kzg_poly_commitment.rs:37-53: affine x big-endian || y big-endian (canonical), 64 zero bytes for the
        identity."""
        a = self._affine_ints()
        if not a:
            return bytes(64)
        return a[0].to_bytes(32, "big") + a[1].to_bytes(32, "big")

    def __eq__(self, other) -> bool:
        return isinstance(other, KZGCommitment) and np.array_equal(self.to_affine(), other.to_affine())


class KZGCommitmentSchemeBN254:
    """KZG over BN254 with the G1 bases resident on the GPU.

    `public_parameter_group_1` is the affine SRS (n, 8) Montgomery limbs, identity = zeros (the reference keeps
    `Vec<G1Projective>` with Z = 1 and re-normalises on every commit, kzg_poly_commitment.rs:287-288; here the
    normalised bases and their window tables are uploaded once).
    """

    def __init__(self, public_parameter_group_1, window_bits: int = 0):
        self.public_parameter_group_1 = ffi.as_u64(public_parameter_group_1, 8)
        self.window_bits = window_bits          # as requested (0 = the engine's rule); info() reports the one in use
        self._handle = ffi.srs_upload(self.public_parameter_group_1, window_bits)

    @classmethod
    def new(cls, max_degree: int, tau, window_bits: int = 0) -> "KZGCommitmentSchemeBN254":
        """KZGCommitmentScheme::new (kzg_poly_commitment.rs:183-204) with the trapdoor given explicitly
        (the reference draws it from the caller's RNG): public_parameter_group_1[i] = tau^i * G, built on the GPU."""
        return cls(ffi.srs_generate(tau, max_degree + 1), window_bits)

    @classmethod
    def new_lagrange(cls, n: int, tau, window_bits: int = 0) -> "KZGCommitmentSchemeBN254":
        """The Lagrange-basis scheme of the same trapdoor: public_parameter_group_1[i] = L_i(tau) * G over the size-n domain -- what
        `lagrange-srs-*.bin` holds (gen_params/mod.rs:42-65) and prover_with_lagrange commits evaluation vectors against."""
        return cls(ffi.srs_generate_lagrange(tau, n), window_bits)

    def derive_lagrange(self, n: int, window_bits: int = 0) -> "KZGCommitmentSchemeBN254":
        """The Lagrange-basis scheme of THIS SRS for the size-n domain, derived without the trapdoor by an inverse transform over
        the first n G1 points (uzkge_cuda_srs_lagrange_from_monomial)."""
        if n > self.public_parameter_group_1.shape[0]:
            raise ParameterError("the monomial SRS is shorter than the domain")
        return KZGCommitmentSchemeBN254(ffi.srs_lagrange_from_monomial(self.public_parameter_group_1, n), window_bits)

    def close(self) -> None:
        if self._handle:
            ffi.srs_free(self._handle)
            self._handle = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self) -> int:
        return self._handle

    def info(self) -> dict:
        return ffi.srs_info(self._handle)

    def max_degree(self) -> int:
        return self.public_parameter_group_1.shape[0] - 1

    def commit(self, polynomial: FpPolynomial) -> KZGCommitment:
        """PolyComScheme::commit (kzg_poly_commitment.rs:278-293)."""
        coefs = polynomial.get_coefs_ref()
        degree = polynomial.degree()
        if degree + 1 > self.public_parameter_group_1.shape[0]:
            raise DegreeError("DegreeError")
        return KZGCommitment(ffi.msm_g1(self._handle, coefs[: degree + 1]))

    def commit_batch(self, polynomials) -> list[KZGCommitment]:
        """Independent commitments of one prover round (plonk/prover.rs:132-192, helpers.rs:1323-1408) in one call."""
        vecs = []
        for p in polynomials:
            if p.degree() + 1 > self.public_parameter_group_1.shape[0]:
                raise DegreeError("DegreeError")
            vecs.append(p.get_coefs_ref()[: p.degree() + 1])
        out = ffi.msm_g1_batch(self._handle, vecs)
        return [KZGCommitment(o) for o in out]

    def eval(self, poly: FpPolynomial, point) -> np.ndarray:
        """PolyComScheme::eval (kzg_poly_commitment.rs:295-297)."""
        return poly.eval(point)

    def prove(self, poly: FpPolynomial, x, max_degree: int) -> KZGCommitment:
        """PolyComScheme::prove (kzg_poly_commitment.rs:315-342): the commitment of (P(X) - P(x)) / (X - x).
        Evaluation, division and commitment all run on the GPU (Horner scan + MSM); the division's remainder is
        P(x), so the quotient of P itself by (X - x) equals the quotient of P - P(x)."""
        if poly.degree() > max_degree:
            raise DegreeError("DegreeError")
        q, _r = poly.div_rem_linear(x)
        return self.commit(q)

    def apply_blind_factors(self, commitment: KZGCommitment, blinds, zeroing_degree: int) -> KZGCommitment:
        """kzg_poly_commitment.rs:299-313: C + sum_i b_i * (SRS[i] - SRS[zeroing_degree + i]), as two tiny MSMs
        over the resident bases (the negation is folded into the scalars)."""
        b = ffi.as_u64(blinds, 4)
        if b.shape[0] == 0:
            return commitment
        if zeroing_degree + b.shape[0] > self.public_parameter_group_1.shape[0]:
            raise ParameterError("blind factors outside the SRS")
        neg = np.empty_like(b)
        for i in range(b.shape[0]):
            v = _limbs_to_int(b[i])
            neg[i] = _int_to_limbs((FR_MODULUS - v) % FR_MODULUS)  # -b in Montgomery form is r - b
        lo = KZGCommitment(ffi.msm_g1(self._handle, b, 0))
        hi = KZGCommitment(ffi.msm_g1(self._handle, neg, zeroing_degree))
        return commitment.add(lo).add(hi)
