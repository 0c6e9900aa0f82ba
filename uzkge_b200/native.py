"""The compiled prover behind the C ABI: uzkge_cuda_plonk_params_upload / uzkge_cuda_plonk_prove (uzkge_b200/csrc/prover.cu).

`NativeProver` is what the patched Rust body of `prover_with_lagrange` (/root/reference/uzkge/src/plonk/prover.rs:88-394) does, in
Python for this Rust-less image: it keeps the serial, cheap host work the reference also does on the host -- transcript_init_plonk
(plonk/transcript.rs:8-31), the prover's `Fr::rand` draws in the reference's order -- and hands everything that touches a polynomial
to ONE library call.  The proof is byte-identical to `plonk.prover` (the call-by-call mirror); tests/test_gpu_prover_native.py checks
that on every feature set and commitment route.  There is no CPU path: without the CUDA library every step raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import ffi
from .errors import DegreeError, ParameterError, UzkgeError
from .poly_commit import FR_MODULUS, KZGCommitment
from .rng import fr_rand_mont
from .transcript import Transcript, fr_mont_rows_to_bytes_be, transcript_init_plonk

_M64 = 0xFFFFFFFFFFFFFFFF
_FR_R_INV = pow((1 << 256) % FR_MODULUS, -1, FR_MODULUS)


class PlonkParamsDesc(C.Structure):
    """uzkge_plonk_params_desc (include/uzkge_cuda.h)."""
    _fields_ = [
        ("n", C.c_uint64), ("m", C.c_uint64), ("num_vars", C.c_uint64),
        ("wiring", C.c_void_p), ("permutation", C.c_void_p),
        ("k", (C.c_uint64 * 4) * 5),
        ("q_polys", C.c_void_p * 9), ("q_len", C.c_size_t * 9),
        ("s_polys", C.c_void_p * 5), ("s_len", C.c_size_t * 5),
        ("qb_poly", C.c_void_p), ("qb_len", C.c_size_t),
        ("q_prk_polys", C.c_void_p * 4), ("q_prk_len", C.c_size_t * 4),
        ("anemoi_generator", C.c_uint64 * 4), ("anemoi_generator_inv", C.c_uint64 * 4),
        ("public_vars_constraint_indices", C.c_void_p), ("public_vars_witness_indices", C.c_void_p), ("n_public", C.c_size_t),
        ("shuffle", C.c_int32), ("reserved", C.c_int32),
        ("q_ecc_poly", C.c_void_p), ("q_ecc_len", C.c_size_t),
        ("q_shuffle_generator_polys", C.c_void_p * 12), ("gen_len", C.c_size_t * 12),
        ("q_shuffle_public_key_polys", C.c_void_p * 12), ("pk_len", C.c_size_t * 12),
        ("edwards_a", C.c_uint64 * 4),
    ]


class PlonkProveArgs(C.Structure):
    """uzkge_plonk_prove_args."""
    _fields_ = [
        ("params", C.c_uint64), ("srs", C.c_uint64), ("lagrange_srs", C.c_uint64),
        ("lagrange_all", C.c_int32), ("witness_on_device", C.c_int32),
        ("witness", C.c_void_p), ("w_sel_evals", C.c_void_p * 3),
        ("blinds", C.c_void_p), ("n_blinds", C.c_size_t),
        ("transcript", C.c_char_p), ("transcript_len", C.c_size_t),
    ]


class PlonkProofOut(C.Structure):
    """uzkge_plonk_proof."""
    _fields_ = [
        ("cm_w", (C.c_uint64 * 8) * 5), ("cm_w_sel", (C.c_uint64 * 8) * 3), ("cm_t", (C.c_uint64 * 8) * 5), ("cm_z", C.c_uint64 * 8),
        ("prk_3_poly_eval_zeta", C.c_uint64 * 4), ("prk_4_poly_eval_zeta", C.c_uint64 * 4),
        ("w_polys_eval_zeta", (C.c_uint64 * 4) * 5), ("w_polys_eval_zeta_omega", (C.c_uint64 * 4) * 3),
        ("z_eval_zeta_omega", C.c_uint64 * 4), ("s_polys_eval_zeta", (C.c_uint64 * 4) * 4),
        ("q_ecc_poly_eval_zeta", C.c_uint64 * 4), ("w_sel_polys_eval_zeta", (C.c_uint64 * 4) * 3),
        ("opening_witness_zeta", C.c_uint64 * 8), ("opening_witness_zeta_omega", C.c_uint64 * 8),
        ("transcript_state", C.c_uint8 * 32),
        ("launches", C.c_uint32),
        ("msm", C.c_uint32), ("ifft_n", C.c_uint32), ("fft_n", C.c_uint32), ("coset_fft_m", C.c_uint32), ("coset_ifft_m", C.c_uint32),
        ("evals", C.c_uint32),
        ("rounds_ms", C.c_double * 6),
    ]


ROUND_NAMES = ("round1_wires", "round2_z", "round3_quotient", "round3_commit_t", "round4_evals_r", "round5_openings")


def _lib():
    return ffi.lib()


def _mont_limbs(x: int, modulus: int) -> np.ndarray:
    v = x % modulus * ((1 << 256) % modulus) % modulus
    return np.array([(v >> (64 * i)) & _M64 for i in range(4)], dtype=np.uint64)


def _fill4(dst, limbs) -> None:
    for i in range(4):
        dst[i] = int(limbs[i])


def _aff_to_commitment(aff) -> KZGCommitment:
    """affine (x, y) Montgomery limbs (zeros = identity) -> KZGCommitment (Jacobian with Z = 1, or arkworks' zero (1, 1, 0))."""
    a = np.array(list(aff), dtype=np.uint64)
    from .poly_commit import FQ_MODULUS

    one = _mont_limbs(1, FQ_MODULUS)
    if not a.any():
        return KZGCommitment(np.concatenate([one, one, np.zeros(4, dtype=np.uint64)]))
    return KZGCommitment(np.concatenate([a, one]))


class NativeProver:
    """One circuit's prover parameters resident in HBM behind a library handle, plus the SRS handles to prove against.

    cs, prover_params: a padded `plonk.TurboCS` and the `plonk.PlonkProverParams` `plonk.indexer` built for it (their coefficient forms
    are read back once and handed to uzkge_cuda_plonk_params_upload -- the same data the Rust struct holds on the host).
    pcs / lagrange_pcs: `KZGCommitmentSchemeBN254` over the monomial SRS / the size-n Lagrange SRS (None: monomial commitments only).
    lagrange_all: commit everything over the Lagrange bases (None = decide from the SRS: on when the monomial SRS has holes below n).
    multi: ONE proof on every GPU of the device group (ffi.init_devices first): the parameters are uploaded to every member and the
    SRS points split over them; `prove` is unchanged and returns the same bytes as the single-device prover."""

    def __init__(self, cs, prover_params, pcs, lagrange_pcs=None, lagrange_all: bool | None = None, multi: bool = False):
        from .plonk import unmont  # noqa: F401  (imported here: plonk imports torch)

        L = _lib()
        P = prover_params
        vp = P.verifier_params
        self.cs, self.P, self.vp, self.pcs = cs, P, vp, pcs
        self.n, self.m = P.n, P.m
        self.shuffle = P.q_ecc_poly is not None
        if lagrange_pcs is not None and lagrange_pcs.max_degree() + 1 != self.n:
            lagrange_pcs = None                                  # prover.rs:119-124
        if lagrange_pcs is None:
            lagrange_all = False
        elif lagrange_all is None:
            lagrange_all = not bool(np.asarray(pcs.public_parameter_group_1[: self.n]).any(axis=1).all())
        self.lagrange_all = bool(lagrange_all)
        self._keep = []

        def host(poly):
            a = np.ascontiguousarray(poly.numpy(self.n))
            self._keep.append(a)
            return a.ctypes.data, a.shape[0]

        d = PlonkParamsDesc()
        d.n, d.m, d.num_vars = self.n, self.m, cs.num_vars
        wiring = np.ascontiguousarray(cs.wiring.reshape(-1).astype(np.uint32))
        perm = np.ascontiguousarray(cs.compute_permutation().astype(np.uint64))
        self._keep += [wiring, perm]
        d.wiring, d.permutation = wiring.ctypes.data, perm.ctypes.data
        from .poly_commit import FR_MODULUS

        for i in range(5):
            _fill4(d.k[i], _mont_limbs(vp.k[i], FR_MODULUS))
        zero_ptr = P.q_prk_polys[0].ptr if not cs.anemoi_constraints_indices else None     # the shared zero polynomial's buffer

        def put(dst_ptrs, dst_lens, i, poly):
            if zero_ptr is not None and poly.ptr == zero_ptr:
                dst_ptrs[i], dst_lens[i] = None, 0                # absent selector: the library shares one zero buffer too
            else:
                dst_ptrs[i], dst_lens[i] = host(poly)

        for i in range(9):
            put(d.q_polys, d.q_len, i, P.q_polys[i])
        for i in range(5):
            put(d.s_polys, d.s_len, i, P.s_polys[i])
        if zero_ptr is not None and P.qb_poly.ptr == zero_ptr:
            d.qb_poly, d.qb_len = None, 0
        else:
            d.qb_poly, d.qb_len = host(P.qb_poly)
        for i in range(4):
            put(d.q_prk_polys, d.q_prk_len, i, P.q_prk_polys[i])
        _fill4(d.anemoi_generator, _mont_limbs(vp.anemoi_generator, FR_MODULUS))
        _fill4(d.anemoi_generator_inv, _mont_limbs(vp.anemoi_generator_inv, FR_MODULUS))
        rows = np.ascontiguousarray(np.asarray(vp.public_vars_constraint_indices, dtype=np.uint64))
        wits = np.ascontiguousarray(np.asarray(cs.public_vars_witness_indices, dtype=np.uint64))
        self._keep += [rows, wits]
        d.n_public = len(rows)
        if len(rows):
            d.public_vars_constraint_indices, d.public_vars_witness_indices = rows.ctypes.data, wits.ctypes.data
        d.shuffle = 1 if self.shuffle else 0
        if self.shuffle:
            if zero_ptr is not None and P.q_ecc_poly.ptr == zero_ptr:
                d.q_ecc_poly, d.q_ecc_len = None, 0
            else:
                d.q_ecc_poly, d.q_ecc_len = host(P.q_ecc_poly)
            for i in range(12):
                put(d.q_shuffle_generator_polys, d.gen_len, i, P.q_shuffle_generator_polys[i])
                put(d.q_shuffle_public_key_polys, d.pk_len, i, P.q_shuffle_public_key_polys[i])
            _fill4(d.edwards_a, _mont_limbs(vp.edwards_a, FR_MODULUS))
        h = C.c_uint64(0)
        self.multi = bool(multi)
        upload = L.uzkge_cuda_plonk_params_upload_multi if self.multi else L.uzkge_cuda_plonk_params_upload
        ffi.check(upload(C.cast(C.pointer(d), C.c_void_p), C.byref(h)), ParameterError)
        self.handle = int(h.value)
        self._keep = []
        self.lagrange_handle = 0
        self.srs_handle = self.pcs.handle if self.pcs is not None else 0
        if self.multi and self.pcs is not None:
            self.srs_handle = ffi.srs_upload_multi(pcs.public_parameter_group_1, ffi.MULTI_SPLIT, getattr(pcs, "window_bits", 0))
        if lagrange_pcs is not None:
            lag_pts = ffi.as_u64(lagrange_pcs.public_parameter_group_1, 8)
            mono = ffi.as_u64(pcs.public_parameter_group_1, 8)
            lh = C.c_uint64(0)
            up = L.uzkge_cuda_srs_upload_lagrange_commit_multi if self.multi else L.uzkge_cuda_srs_upload_lagrange_commit
            ffi.check(up(ffi.ptr(lag_pts), self.n, ffi.ptr(mono), mono.shape[0], getattr(lagrange_pcs, "window_bits", 0), C.byref(lh)),
                      ParameterError)
            self.lagrange_handle = int(lh.value)
        self.last_stats: dict = {}
        self._sel_cache = None
        self._tinit_prefix = None
        self._pub_idx = None

    def refresh_public_key(self) -> None:
        """After plonk.refresh_prover_params_public_key changed the 12 public-key selector polynomials of `prover_params`."""
        polys = [np.ascontiguousarray(p.numpy(self.n)) for p in self.P.q_shuffle_public_key_polys]
        ptrs = (C.c_void_p * 12)(*[a.ctypes.data for a in polys])
        lens = (C.c_size_t * 12)(*[a.shape[0] for a in polys])
        ffi.check(_lib().uzkge_cuda_plonk_params_set_public_key(self.handle, ptrs, lens), ParameterError)

    def prove(self, prng, transcript, w, timings: dict | None = None):
        """plonk/prover.rs:76-394.  `w`: the witness, (num_vars, 4) Montgomery limbs (numpy), or a plonk.DevVec already in HBM."""
        from .plonk import DevVec, PlonkProof, unmont

        cs = self.cs
        if cs.is_verifier_only():
            raise UzkgeError("FuncParamsError")
        on_device = isinstance(w, DevVec)
        if (w.len if on_device else np.asarray(w).shape[0]) != cs.num_vars:
            raise ParameterError("witness length != num_vars")
        idx = cs.public_vars_witness_indices
        if on_device:
            if idx:
                import torch

                if self._pub_idx is None or self._pub_idx.device != w.t.device:
                    self._pub_idx = torch.as_tensor(np.asarray(idx, dtype=np.int64), device=w.t.device)
                rows = w.t.view(-1, 4)[self._pub_idx].cpu().numpy().view(np.uint64)      # the public inputs only, not the witness
            else:
                rows = np.zeros((0, 4), dtype=np.uint64)
            w_ptr, keep_w = w.ptr, w
        else:
            wa = ffi.as_u64(w, 4)
            rows = wa[idx] if idx else np.zeros((0, 4), dtype=np.uint64)
            w_ptr, keep_w = wa.ctypes.data, wa
        # transcript_init_plonk (plonk/transcript.rs:8-31): everything but the public inputs is fixed per circuit
        if self._tinit_prefix is None:
            pre = Transcript(b"")
            pre.state = bytearray()
            transcript_init_plonk(pre, self.vp, [], self.P.root)
            self._tinit_prefix = bytes(pre.state)
        transcript.state.extend(self._tinit_prefix)
        transcript.state.extend(fr_mont_rows_to_bytes_be(rows))
        n_blinds = 21 + (6 if self.shuffle else 0)
        blinds = np.zeros((n_blinds, 4), dtype=np.uint64)
        for j in range(n_blinds):
            raw = fr_rand_mont(prng)
            blinds[j] = [(raw >> (64 * i)) & _M64 for i in range(4)]
        a = PlonkProveArgs()
        a.params, a.srs, a.lagrange_srs = self.handle, self.srs_handle, self.lagrange_handle
        a.lagrange_all = 1 if self.lagrange_all else 0
        a.witness_on_device = 1 if on_device else 0
        a.witness = w_ptr
        sel = None
        if self.shuffle and cs.shuffle_remark_constraints:
            if self._sel_cache is None or self._sel_cache[0] is not cs._remark_sel_codes:      # per witness: the remark traces' bits / signs
                self._sel_cache = (cs._remark_sel_codes, np.ascontiguousarray(cs.compute_witness_selectors()))
            sel = self._sel_cache[1]
            for i in range(3):
                a.w_sel_evals[i] = sel[i].ctypes.data
        a.blinds, a.n_blinds = blinds.ctypes.data, n_blinds
        state = bytes(transcript.state)
        a.transcript, a.transcript_len = state, len(state)
        out = PlonkProofOut()
        rc = _lib().uzkge_cuda_plonk_prove(C.cast(C.pointer(a), C.c_void_p), C.cast(C.pointer(out), C.c_void_p))
        del keep_w, sel
        if rc == ffi.ERR_SIZE and "DegreeError" in ffi.last_error():
            raise DegreeError("DegreeError")
        ffi.check(rc, UzkgeError)
        transcript.state = bytearray(bytes(out.transcript_state))
        self.last_stats = {"launches": out.launches, "msm": out.msm, "ifft_n": out.ifft_n, "fft_n": out.fft_n, "coset_fft_m": out.coset_fft_m,
                           "coset_ifft_m": out.coset_ifft_m, "evals": out.evals, "rounds_ms": dict(zip(ROUND_NAMES, list(out.rounds_ms)))}
        if timings is not None:
            for k, v in self.last_stats["rounds_ms"].items():
                timings[k] = timings.get(k, 0.0) + v
        sc = lambda limbs: unmont(list(limbs))
        cm = _aff_to_commitment
        return PlonkProof(
            cm_w_vec=[cm(c) for c in out.cm_w], cm_t_vec=[cm(c) for c in out.cm_t], cm_z=cm(out.cm_z),
            prk_3_poly_eval_zeta=sc(out.prk_3_poly_eval_zeta), prk_4_poly_eval_zeta=sc(out.prk_4_poly_eval_zeta),
            w_polys_eval_zeta=[sc(v) for v in out.w_polys_eval_zeta], w_polys_eval_zeta_omega=[sc(v) for v in out.w_polys_eval_zeta_omega],
            z_eval_zeta_omega=sc(out.z_eval_zeta_omega), s_polys_eval_zeta=[sc(v) for v in out.s_polys_eval_zeta],
            opening_witness_zeta=cm(out.opening_witness_zeta), opening_witness_zeta_omega=cm(out.opening_witness_zeta_omega),
            cm_w_sel_vec=[cm(c) for c in out.cm_w_sel] if self.shuffle else None,
            q_ecc_poly_eval_zeta=sc(out.q_ecc_poly_eval_zeta) if self.shuffle else None,
            w_sel_polys_eval_zeta=[sc(v) for v in out.w_sel_polys_eval_zeta] if self.shuffle else None)

    def close(self) -> None:
        if getattr(self, "handle", 0):
            _lib().uzkge_cuda_plonk_params_free(self.handle)
            self.handle = 0
        if getattr(self, "lagrange_handle", 0):
            ffi.srs_free(self.lagrange_handle)
            self.lagrange_handle = 0
        if getattr(self, "multi", False) and getattr(self, "srs_handle", 0):
            ffi.srs_free(self.srs_handle)
            self.srs_handle = 0

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
