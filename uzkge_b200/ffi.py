"""ctypes binding of the C ABI in include/uzkge_cuda.h -- the same symbols the Rust `uzkge-cuda-sys` crate binds
(INTEGRATION.md).  Arrays are numpy uint64 in the ABI layout: Fr/Fq (n, 4) Montgomery limbs, affine (n, 8),
Jacobian (12,).  Loading fails loudly when the library is missing: there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

from .errors import BackendUnavailable, CommitmentError, FFTError, ParameterError, UzkgeError

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libuzkge_cuda.so")
HEADER_PATH = os.path.join(HERE, "..", "include", "uzkge_cuda.h")

OK, ERR_NO_DEVICE, ERR_SIZE, ERR_CUDA, ERR_OOM, ERR_HANDLE, ERR_ARG, ERR_INTERNAL = range(8)

u64p = C.POINTER(C.c_uint64)


class SrsInfo(C.Structure):
    _fields_ = [
        ("window_bits", C.c_uint32),
        ("windows", C.c_uint32),
        ("n", C.c_uint64),
        ("device_bytes", C.c_uint64),
        ("batch_slots", C.c_uint32),
        ("reserved", C.c_uint32),
        ("precompute_ms", C.c_double),
    ]


class QuotientArgs(C.Structure):
    """uzkge_quotient_args (include/uzkge_cuda.h)."""

    _fields_ = [
        ("w", C.c_void_p * 5),
        ("q", C.c_void_p * 9),
        ("pi", C.c_void_p),
        ("z", C.c_void_p),
        ("s", C.c_void_p * 5),
        ("coset_quotient", C.c_void_p),
        ("l1", C.c_void_p),
        ("qb", C.c_void_p),
        ("q_prk", C.c_void_p * 4),
        ("k", (C.c_uint64 * 4) * 5),
        ("alpha", C.c_uint64 * 4),
        ("beta", C.c_uint64 * 4),
        ("gamma", C.c_uint64 * 4),
        ("anemoi_generator", C.c_uint64 * 4),
        ("anemoi_generator_inv", C.c_uint64 * 4),
        ("z_h_inv", (C.c_uint64 * 4) * 16),
        ("m", C.c_size_t),
        ("factor", C.c_size_t),
    ]


class QuotientShuffleArgs(C.Structure):
    """uzkge_quotient_shuffle_args (include/uzkge_cuda.h)."""
    _fields_ = [
        ("w_sel", C.c_void_p * 3),
        ("q_ecc", C.c_void_p),
        ("pk", C.c_void_p * 12),
        ("gen", C.c_void_p * 12),
        ("edwards_a", C.c_uint64 * 4),
    ]


_SIGNATURES = {
    "uzkge_cuda_init": (C.c_int32, [C.c_int32]),
    "uzkge_cuda_device_count": (C.c_int32, []),
    "uzkge_cuda_set_device": (C.c_int32, [C.c_int32]),
    "uzkge_cuda_get_device": (C.c_int32, []),
    "uzkge_cuda_init_devices": (C.c_int32, [C.c_int32]),
    "uzkge_cuda_group_size": (C.c_int32, []),
    "uzkge_cuda_srs_upload_multi": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_uint32, C.c_int32, u64p]),
    "uzkge_cuda_last_error": (C.c_char_p, []),
    "uzkge_cuda_version": (C.c_char_p, []),
    "uzkge_cuda_srs_upload": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_uint32, u64p]),
    "uzkge_cuda_srs_generate": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "uzkge_cuda_srs_generate_lagrange": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "uzkge_cuda_msm_g1_small_device": (C.c_int32, [C.c_uint64, C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p, C.c_void_p]),
    "uzkge_cuda_srs_lagrange_from_monomial": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_void_p]),
    "uzkge_cuda_srs_free": (C.c_int32, [C.c_uint64]),
    "uzkge_cuda_msm_g1": (C.c_int32, [C.c_uint64, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p]),
    "uzkge_cuda_msm_g1_batch": (C.c_int32, [C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_size_t, C.c_void_p]),
    "uzkge_cuda_ntt_fr": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int32, C.c_void_p]),
    "uzkge_cuda_ntt_fr_batch": (C.c_int32, [C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_size_t, C.c_size_t, C.c_int32, C.c_void_p]),
    "uzkge_cuda_fr_root_of_unity": (C.c_int32, [C.c_size_t, C.c_void_p]),
    "uzkge_cuda_msm_g1_device": (C.c_int32, [C.c_uint64, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "uzkge_cuda_msm_g1_batch_device": (
        C.c_int32,
        [C.c_uint64, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_size_t, C.c_void_p, C.c_void_p],
    ),
    "uzkge_cuda_ntt_fr_device": (
        C.c_int32,
        [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int32, C.c_void_p, C.c_void_p],
    ),
    "uzkge_cuda_ntt_fr_batch_device": (
        C.c_int32,
        [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.POINTER(C.c_size_t), C.c_size_t, C.c_size_t, C.c_int32, C.c_void_p, C.c_void_p],
    ),
    "uzkge_cuda_ntt_cross_fr_device": (
        C.c_int32,
        [C.c_void_p, C.c_void_p, C.c_uint32, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int32, C.c_void_p],
    ),
    "uzkge_cuda_ntt_cross_rows_fr_device": (
        C.c_int32,
        [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_uint32, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int32, C.c_void_p],
    ),
    "uzkge_cuda_ntt_fr_multi": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_int32, C.c_void_p]),
    "uzkge_cuda_ntt_fr_scatter_device": (C.c_int32, [C.c_void_p, C.POINTER(C.c_void_p), C.c_void_p, C.c_size_t, C.c_int32, C.c_uint32, C.c_uint32, C.c_void_p]),
    "uzkge_cuda_dev_alloc": (C.c_int32, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "uzkge_cuda_dev_free": (C.c_int32, [C.c_void_p]),
    "uzkge_cuda_dev_copy_in": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "uzkge_cuda_dev_copy_out": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t]),
    "uzkge_cuda_ipc_export": (C.c_int32, [C.c_void_p, C.c_char_p]),
    "uzkge_cuda_ipc_open": (C.c_int32, [C.c_char_p, C.POINTER(C.c_void_p)]),
    "uzkge_cuda_ipc_close": (C.c_int32, [C.c_void_p]),
    "uzkge_cuda_poly_eval_fr": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "uzkge_cuda_poly_div_linear_fr": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]),
    "uzkge_cuda_poly_horner_fr_device": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]),
    "uzkge_cuda_poly_eval_batch_fr_device": (
        C.c_int32,
        [C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(C.c_uint32), C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p],
    ),
    "uzkge_cuda_grand_product_fr": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "uzkge_cuda_plonk_quotient_fr_device": (C.c_int32, [C.POINTER(QuotientArgs), C.c_void_p, C.c_void_p]),
    "uzkge_cuda_plonk_quotient_shuffle_fr_device": (C.c_int32, [C.POINTER(QuotientArgs), C.POINTER(QuotientShuffleArgs), C.c_void_p, C.c_void_p]),
    "uzkge_cuda_fr_lincomb_device": (
        C.c_int32,
        [C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p],
    ),
    "uzkge_cuda_fr_add_sparse_device": (C.c_int32, [C.c_void_p, C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t, C.c_void_p]),
    "uzkge_cuda_fr_add_sparse_multi_device": (C.c_int32, [C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.c_void_p, C.c_size_t, C.c_void_p]),
    "uzkge_cuda_fr_powers_device": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "uzkge_cuda_fr_gather_device": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "uzkge_cuda_fr_gather_scatter_device": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p]),
    "uzkge_cuda_fr_mul_device": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p]),
    "uzkge_cuda_fr_trimmed_len_device": (C.c_int32, [C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t), C.c_void_p]),
    "uzkge_cuda_grand_product_fr_device": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]),
    "uzkge_cuda_plonk_z_evals_fr_device": (
        C.c_int32,
        [C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p,
         C.c_void_p, C.c_void_p],
    ),
    # the compiled prover (csrc/prover.cu); native.py declares the structures behind the void pointers
    "uzkge_cuda_plonk_params_upload": (C.c_int32, [C.c_void_p, u64p]),
    "uzkge_cuda_plonk_params_upload_multi": (C.c_int32, [C.c_void_p, u64p]),
    "uzkge_cuda_srs_upload_lagrange_commit_multi": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint32, u64p]),
    "uzkge_cuda_plonk_quotient_range_fr_device": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p]),
    "uzkge_cuda_plonk_coset_combine_fr_device": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p]),
    "uzkge_cuda_fr_strided_copy_device": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p]),
    "uzkge_cuda_plonk_params_set_public_key": (C.c_int32, [C.c_uint64, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t)]),
    "uzkge_cuda_plonk_params_free": (C.c_int32, [C.c_uint64]),
    "uzkge_cuda_srs_upload_lagrange_commit": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_uint32, u64p]),
    "uzkge_cuda_plonk_prove": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "uzkge_cuda_g1_add": (C.c_int32, [C.c_void_p, C.c_void_p, C.c_void_p]),
    "uzkge_cuda_g1_sum_device": (C.c_int32, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p, C.c_void_p]),
    "uzkge_cuda_g1_to_affine": (C.c_int32, [C.c_void_p, C.c_void_p]),
    "uzkge_cuda_host_alloc": (C.c_int32, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "uzkge_cuda_host_free": (C.c_int32, [C.c_void_p]),
    "uzkge_cuda_host_register": (C.c_int32, [C.c_void_p, C.c_size_t]),
    "uzkge_cuda_host_unregister": (C.c_int32, [C.c_void_p]),
    "uzkge_cuda_srs_info": (C.c_int32, [C.c_uint64, C.POINTER(SrsInfo)]),
    "uzkge_cuda_launch_count": (C.c_uint64, []),
    "uzkge_cuda_profile_enable": (C.c_int32, [C.c_int32]),
    "uzkge_cuda_profile_read": (C.c_int32, [C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]),
    "uzkge_cuda_configure": (C.c_int32, [C.c_char_p, C.c_uint64]),
    "uzkge_cuda_field_mul": (C.c_int32, [C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]),
    "uzkge_cuda_bench_field_mul": (C.c_int32, [C.c_int32, C.c_uint32, C.POINTER(C.c_double)]),
}


def header_symbols() -> list[str]:
    """Every entry point include/uzkge_cuda.h declares (used by the CPU test-suite's export check)."""
    txt = open(HEADER_PATH).read()
    return sorted(set(re.findall(r"UZKGE_API\s+[\w\s\*]+?\b(uzkge_cuda_\w+)\s*\(", txt)))


_lib = None


def lib() -> C.CDLL:
    """Load libuzkge_cuda.so and declare every prototype.  No GPU is touched until an entry point is called."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BackendUnavailable(
                f"{LIB_PATH} is missing: build it with `python -m uzkge_b200.build` (nvcc, sm_100a). "
                "uzkge_b200 has no CPU fallback."
            )
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def last_error() -> str:
    return (lib().uzkge_cuda_last_error() or b"").decode()


def check(rc: int, exc=UzkgeError):
    if rc == OK:
        return
    msg = f"uzkge_cuda error {rc}: {last_error()}"
    if rc == ERR_NO_DEVICE:
        raise BackendUnavailable(msg)
    if rc in (ERR_SIZE, ERR_ARG, ERR_HANDLE) and exc is UzkgeError:
        raise ParameterError(msg)
    raise exc(msg)


def as_u64(a, cols: int | None = None) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=np.uint64)
    if cols is not None:
        a = a.reshape(-1, cols)
    return a


def ptr(a: np.ndarray) -> C.c_void_p:
    return C.c_void_p(a.ctypes.data)


def init(device: int = -1) -> None:
    check(lib().uzkge_cuda_init(device))


def device_count() -> int:
    return int(lib().uzkge_cuda_device_count())


def set_device(device: int) -> None:
    check(lib().uzkge_cuda_set_device(device))


def get_device() -> int:
    return int(lib().uzkge_cuda_get_device())


def init_devices(device_count: int = 0) -> int:
    """One process, several GPUs: initialise devices 0 .. device_count - 1 (0 = all visible) as the device group; returns its size."""
    check(lib().uzkge_cuda_init_devices(device_count))
    return int(lib().uzkge_cuda_group_size())


MULTI_SPLIT, MULTI_REPLICATED = 0, 1


def srs_upload_multi(affine_xy: np.ndarray, mode: int = MULTI_SPLIT, window_bits: int = 0) -> int:
    """The SRS over the device group (init_devices): points split per GPU, or replicated for dealing a round's commitments."""
    pts = as_u64(affine_xy, 8)
    h = C.c_uint64(0)
    check(lib().uzkge_cuda_srs_upload_multi(ptr(pts), pts.shape[0], window_bits, mode, C.byref(h)), CommitmentError)
    return int(h.value)


def version() -> str:
    return lib().uzkge_cuda_version().decode()


def launch_count() -> int:
    return int(lib().uzkge_cuda_launch_count())


MSM_PHASES = ("recode", "sort", "offsets", "accumulate", "large", "reduce")
NTT_PHASES = ("radix3", "pass0", "pass1", "pass2")


def profile_enable(on: bool) -> None:
    check(lib().uzkge_cuda_profile_enable(1 if on else 0))


def profile_read(kind: str) -> dict:
    """Average per-phase device milliseconds per engine run since the last read."""
    ms = (C.c_double * 8)()
    runs = C.c_uint64(0)
    check(lib().uzkge_cuda_profile_read(0 if kind == "msm" else 1, ms, C.byref(runs)))
    names = MSM_PHASES if kind == "msm" else NTT_PHASES
    r = max(1, int(runs.value))
    return {"runs": int(runs.value), "ms": {n: ms[i] / r for i, n in enumerate(names)}}


def configure(key: str, value: int) -> None:
    check(lib().uzkge_cuda_configure(key.encode(), value))


def srs_upload(affine_xy: np.ndarray, window_bits: int = 0) -> int:
    pts = as_u64(affine_xy, 8)
    h = C.c_uint64(0)
    check(lib().uzkge_cuda_srs_upload(ptr(pts), pts.shape[0], window_bits, C.byref(h)), CommitmentError)
    return int(h.value)


def srs_generate(tau, n: int) -> np.ndarray:
    """(n, 8) affine powers-of-tau SRS: tau^i * G."""
    t = as_u64(tau).reshape(4)
    out = np.zeros((n, 8), dtype=np.uint64)
    check(lib().uzkge_cuda_srs_generate(ptr(t), n, ptr(out)), CommitmentError)
    return out


def srs_generate_lagrange(tau, n: int) -> np.ndarray:
    """L_i(tau) * G, i < n (n a power of two), affine Montgomery limbs, built on the GPU."""
    t = as_u64(tau).reshape(4)
    out = np.zeros((n, 8), dtype=np.uint64)
    check(lib().uzkge_cuda_srs_generate_lagrange(ptr(t), n, ptr(out)))
    return out


def srs_lagrange_from_monomial(monomial_affine_xy, n: int) -> np.ndarray:
    """(1 / n) sum_j w^(-i j) monomial[j]: the Lagrange SRS of the size-n domain from the first n monomial points, no trapdoor."""
    pts = as_u64(monomial_affine_xy, 8)
    assert pts.shape[0] >= n
    src = np.ascontiguousarray(pts[:n])
    out = np.zeros((n, 8), dtype=np.uint64)
    check(lib().uzkge_cuda_srs_lagrange_from_monomial(ptr(src), n, ptr(out)))
    return out


def msm_g1_small_device(handle: int, idx, scalars, d_out: int, accumulate: bool = False, stream: int = 0) -> None:
    """*d_out (+)= sum_j scalars[j] * srs[idx[j]] for a handful of terms (blind factors)."""
    k = len(idx)
    s = as_u64(scalars, 4)
    assert s.shape[0] == k
    ii = (C.c_size_t * k)(*[int(x) for x in idx])
    check(lib().uzkge_cuda_msm_g1_small_device(handle, ii, ptr(s), k, 1 if accumulate else 0, d_out, stream), CommitmentError)


def srs_free(handle: int) -> None:
    check(lib().uzkge_cuda_srs_free(handle))


def srs_info(handle: int) -> dict:
    info = SrsInfo()
    check(lib().uzkge_cuda_srs_info(handle, C.byref(info)))
    return {k: getattr(info, k) for k, _ in SrsInfo._fields_}


def msm_g1(handle: int, scalars: np.ndarray, base_offset: int = 0) -> np.ndarray:
    s = as_u64(scalars, 4)
    out = np.zeros(12, dtype=np.uint64)
    check(lib().uzkge_cuda_msm_g1(handle, base_offset, ptr(s), s.shape[0], ptr(out)), CommitmentError)
    return out


def msm_g1_batch(handle: int, scalar_vectors) -> np.ndarray:
    vecs = [as_u64(v, 4) for v in scalar_vectors]
    k = len(vecs)
    out = np.zeros((k, 12), dtype=np.uint64)
    if k == 0:
        return out
    ptrs = (C.c_void_p * k)(*[v.ctypes.data for v in vecs])
    lens = (C.c_size_t * k)(*[v.shape[0] for v in vecs])
    check(lib().uzkge_cuda_msm_g1_batch(handle, ptrs, lens, k, ptr(out)), CommitmentError)
    return out


def ntt_fr(data: np.ndarray, domain_size: int, inverse: bool = False, coset_shift=None) -> np.ndarray:
    """Returns a new (domain_size, 4) array; `data` holds the first len_in elements."""
    d = as_u64(data, 4)
    len_in = d.shape[0]
    if len_in > domain_size:
        raise ParameterError("input longer than the domain")
    buf = np.empty((domain_size, 4), dtype=np.uint64)
    buf[:len_in] = d
    shift = as_u64(coset_shift, 4) if coset_shift is not None else None
    check(
        lib().uzkge_cuda_ntt_fr(ptr(buf), len_in, domain_size, 1 if inverse else 0, ptr(shift) if shift is not None else None),
        FFTError,
    )
    return buf


def ntt_fr_multi_inplace(buf: np.ndarray, len_in: int, domain_size: int, inverse: bool = False, coset_shift=None) -> None:
    """uzkge_cuda_ntt_fr_multi: the same contract as ntt_fr_inplace, ONE transform over the whole device group (init_devices)."""
    assert buf.dtype == np.uint64 and buf.flags["C_CONTIGUOUS"] and buf.size >= 4 * domain_size
    shift = as_u64(coset_shift, 4) if coset_shift is not None else None
    check(
        lib().uzkge_cuda_ntt_fr_multi(ptr(buf), len_in, domain_size, 1 if inverse else 0, ptr(shift) if shift is not None else None),
        FFTError,
    )


def ntt_fr_inplace(buf: np.ndarray, len_in: int, domain_size: int, inverse: bool = False, coset_shift=None) -> None:
    """The raw call: `buf` (capacity domain_size, any host memory incl. pinned) is transformed in place."""
    assert buf.dtype == np.uint64 and buf.flags["C_CONTIGUOUS"] and buf.size >= 4 * domain_size
    shift = as_u64(coset_shift, 4) if coset_shift is not None else None
    check(
        lib().uzkge_cuda_ntt_fr(ptr(buf), len_in, domain_size, 1 if inverse else 0, ptr(shift) if shift is not None else None),
        FFTError,
    )


def ntt_fr_batch_inplace(bufs, lens_in, domain_size: int, inverse: bool = False, coset_shift=None) -> None:
    """In-place transforms of k host buffers (each (domain_size, 4) uint64, C-contiguous, ideally pinned) in one pipelined call."""
    k = len(bufs)
    for b in bufs:
        assert b.dtype == np.uint64 and b.flags["C_CONTIGUOUS"] and b.size == 4 * domain_size
    pp = (C.c_void_p * k)(*[b.ctypes.data for b in bufs])
    ll = (C.c_size_t * k)(*[int(x) for x in lens_in])
    shift = as_u64(coset_shift, 4) if coset_shift is not None else None
    check(lib().uzkge_cuda_ntt_fr_batch(pp, ll, k, domain_size, 1 if inverse else 0, ptr(shift) if shift is not None else None), FFTError)


def fr_root_of_unity(n: int) -> np.ndarray:
    out = np.zeros(4, dtype=np.uint64)
    check(lib().uzkge_cuda_fr_root_of_unity(n, ptr(out)))
    return out


def msm_g1_device(handle: int, d_scalars: int, n: int, d_out: int, stream: int = 0, base_offset: int = 0) -> None:
    check(lib().uzkge_cuda_msm_g1_device(handle, base_offset, d_scalars, n, d_out, stream), CommitmentError)


def msm_g1_batch_device(handle: int, d_scalar_ptrs, ns, d_out: int, stream: int = 0, base_offset: int = 0) -> None:
    k = len(d_scalar_ptrs)
    ptrs = (C.c_void_p * k)(*d_scalar_ptrs)
    lens = (C.c_size_t * k)(*ns)
    check(lib().uzkge_cuda_msm_g1_batch_device(handle, base_offset, ptrs, lens, k, d_out, stream), CommitmentError)


def ntt_fr_device(d_in: int, d_out: int, d_scratch: int, len_in: int, domain_size: int, inverse: bool = False,
                  coset_shift=None, stream: int = 0) -> None:
    shift = as_u64(coset_shift, 4) if coset_shift is not None else None
    check(
        lib().uzkge_cuda_ntt_fr_device(d_in, d_out, d_scratch, len_in, domain_size, 1 if inverse else 0,
                                       ptr(shift) if shift is not None else None, stream),
        FFTError,
    )


def ntt_fr_batch_device(d_ins, d_outs, d_scratch: int, len_in, domain_size: int, inverse: bool = False, coset_shift=None,
                        stream: int = 0) -> None:
    """k transforms over one domain in one launch per pass; d_scratch holds k * domain_size elements."""
    k = len(d_ins)
    ins = (C.c_void_p * k)(*d_ins)
    outs = (C.c_void_p * k)(*d_outs)
    lens = (C.c_size_t * k)(*len_in)
    shift = as_u64(coset_shift, 4) if coset_shift is not None else None
    check(lib().uzkge_cuda_ntt_fr_batch_device(ins, outs, d_scratch, lens, k, domain_size, 1 if inverse else 0,
                                               ptr(shift) if shift is not None else None, stream), FFTError)


def ntt_cross_fr_device(d_in: int, d_out: int, log_ranks: int, cols: int, col_offset: int, n_total: int,
                        inverse: bool = False, stream: int = 0) -> None:
    check(lib().uzkge_cuda_ntt_cross_fr_device(d_in, d_out, log_ranks, cols, col_offset, n_total, 1 if inverse else 0, stream),
          FFTError)


def ntt_fr_scatter_device(d_in: int, d_out_rows, d_scratch: int, n: int, inverse: bool, log_ranks: int, rank: int, stream: int = 0) -> None:
    """The local transform of a distributed four-step whose last pass stores into the owners' natural slices (peer memory)."""
    rows = (C.c_void_p * len(d_out_rows))(*d_out_rows)
    check(lib().uzkge_cuda_ntt_fr_scatter_device(d_in, rows, d_scratch, n, 1 if inverse else 0, log_ranks, rank, stream), FFTError)


def ntt_cross_rows_fr_device(d_in_rows, d_out_rows, log_ranks: int, cols: int, col_offset: int, n_total: int,
                             inverse: bool = False, stream: int = 0) -> None:
    g = 1 << log_ranks
    ii = (C.c_void_p * g)(*[int(x) for x in d_in_rows])
    oo = (C.c_void_p * g)(*[int(x) for x in d_out_rows])
    check(lib().uzkge_cuda_ntt_cross_rows_fr_device(ii, oo, log_ranks, cols, col_offset, n_total, 1 if inverse else 0, stream), FFTError)


def dev_alloc(nbytes: int) -> int:
    p = C.c_void_p()
    check(lib().uzkge_cuda_dev_alloc(nbytes, C.byref(p)))
    return int(p.value)


def dev_free(d_ptr: int) -> None:
    check(lib().uzkge_cuda_dev_free(d_ptr))


def ipc_export(d_ptr: int) -> bytes:
    buf = C.create_string_buffer(64)
    check(lib().uzkge_cuda_ipc_export(d_ptr, buf))
    return buf.raw


def ipc_open(handle: bytes) -> int:
    p = C.c_void_p()
    check(lib().uzkge_cuda_ipc_open(handle, C.byref(p)))
    return int(p.value)


def ipc_close(d_ptr: int) -> None:
    check(lib().uzkge_cuda_ipc_close(d_ptr))


def poly_eval_fr(coefs, x) -> np.ndarray:
    c = as_u64(coefs, 4)
    out = np.zeros(4, dtype=np.uint64)
    check(lib().uzkge_cuda_poly_eval_fr(ptr(c), c.shape[0], ptr(as_u64(x).reshape(4)), ptr(out)))
    return out


def poly_div_linear_fr(coefs, z):
    """(quotient (n - 1, 4), remainder (4,)) of p / (X - z)."""
    c = as_u64(coefs, 4)
    q = np.zeros((max(c.shape[0] - 1, 0), 4), dtype=np.uint64)
    rem = np.zeros(4, dtype=np.uint64)
    check(lib().uzkge_cuda_poly_div_linear_fr(ptr(c), c.shape[0], ptr(as_u64(z).reshape(4)), ptr(q) if q.size else None, ptr(rem)))
    return q, rem


def poly_horner_fr_device(d_coefs: int, n: int, z, d_quotient: int, d_value: int, stream: int = 0) -> None:
    check(lib().uzkge_cuda_poly_horner_fr_device(d_coefs, n, ptr(as_u64(z).reshape(4)), d_quotient or None, d_value, stream))


EVAL_BATCH_MAX = 32


def poly_eval_batch_fr_device(d_polys, lens, point_index, points, d_values: int, stream: int = 0) -> None:
    """d_values[j] = polys[j](points[point_index[j]]); points: (1 or 2, 4) Montgomery limbs on the host."""
    k = len(d_polys)
    pts = as_u64(points, 4)
    pp = (C.c_void_p * k)(*[int(x) for x in d_polys])
    ll = (C.c_size_t * k)(*[int(x) for x in lens])
    ii = (C.c_uint32 * k)(*[int(x) for x in point_index])
    check(lib().uzkge_cuda_poly_eval_batch_fr_device(pp, ll, ii, k, ptr(pts), pts.shape[0], d_values, stream))


def grand_product_fr(num, den) -> np.ndarray:
    a, b = as_u64(num, 4), as_u64(den, 4)
    assert a.shape == b.shape
    out = np.zeros((a.shape[0] + 1, 4), dtype=np.uint64)
    check(lib().uzkge_cuda_grand_product_fr(ptr(a), ptr(b), a.shape[0], ptr(out)))
    return out


def plonk_quotient_fr_device(w, q, pi, z, s, coset_quotient, l1, qb, q_prk, k, alpha, beta, gamma, anemoi_g, anemoi_g_inv,
                             z_h_inv, m: int, factor: int, d_out: int, stream: int = 0, shuffle=None, point_range=None) -> None:
    """Device pointers (ints) for the arrays, numpy Montgomery limbs for the scalars; see uzkge_quotient_args.
    shuffle: None, or a dict {w_sel: 3 pointers, q_ecc: pointer, pk: 12 pointers, gen: 12 pointers, edwards_a: limbs} for the
    `shuffle` feature set (uzkge_cuda_plonk_quotient_shuffle_fr_device).
    point_range: None (all m points) or (start, step, count): only the points start + step * i (uzkge_cuda_plonk_quotient_range_fr_device)."""
    a = QuotientArgs()
    for j in range(5):
        a.w[j], a.s[j] = w[j], s[j]
        a.k[j][:] = [int(v) for v in as_u64(k[j]).reshape(4)]
    for j in range(9):
        a.q[j] = q[j]
    for j in range(4):
        a.q_prk[j] = q_prk[j]
    a.pi, a.z, a.coset_quotient, a.l1, a.qb = pi, z, coset_quotient, l1, qb
    for name, v in (("alpha", alpha), ("beta", beta), ("gamma", gamma), ("anemoi_generator", anemoi_g),
                    ("anemoi_generator_inv", anemoi_g_inv)):
        getattr(a, name)[:] = [int(x) for x in as_u64(v).reshape(4)]
    zh = as_u64(z_h_inv, 4)
    for i in range(zh.shape[0]):
        a.z_h_inv[i][:] = [int(x) for x in zh[i]]
    a.m, a.factor = m, factor
    if shuffle is None:
        if point_range is not None:
            check(lib().uzkge_cuda_plonk_quotient_range_fr_device(C.byref(a), None, *[int(x) for x in point_range], d_out, stream))
        else:
            check(lib().uzkge_cuda_plonk_quotient_fr_device(C.byref(a), d_out, stream))
        return
    b = QuotientShuffleArgs()
    for j in range(3):
        b.w_sel[j] = shuffle["w_sel"][j]
    b.q_ecc = shuffle["q_ecc"]
    for j in range(12):
        b.pk[j], b.gen[j] = shuffle["pk"][j], shuffle["gen"][j]
    b.edwards_a[:] = [int(x) for x in as_u64(shuffle["edwards_a"]).reshape(4)]
    if point_range is not None:
        check(lib().uzkge_cuda_plonk_quotient_range_fr_device(C.byref(a), C.byref(b), *[int(x) for x in point_range], d_out, stream))
        return
    check(lib().uzkge_cuda_plonk_quotient_shuffle_fr_device(C.byref(a), C.byref(b), d_out, stream))


def plonk_coset_combine_fr_device(d_u: int, n: int, factor: int, k1, d_out: int, stream: int = 0) -> None:
    """t's factor * n coefficients from the per-coset inverse transforms u_j (compact, factor x n); k1: Montgomery limbs."""
    check(lib().uzkge_cuda_plonk_coset_combine_fr_device(d_u, n, factor, ptr(as_u64(k1).reshape(4)), d_out, stream))


def fr_strided_copy_device(d_src: int, src_start: int, src_step: int, d_dst: int, dst_start: int, dst_step: int, count: int,
                           stream: int = 0) -> None:
    """dst[dst_start + dst_step * i] = src[src_start + src_step * i], i < count."""
    check(lib().uzkge_cuda_fr_strided_copy_device(d_src, src_start, src_step, d_dst, dst_start, dst_step, count, stream))


LINCOMB_MAX = 24
SPARSE_MAX = 16


def fr_lincomb_device(d_polys, lens, coefs, d_out: int, out_len: int, stream: int = 0) -> None:
    """out[i] = sum_j coefs[j] * polys[j][i]  (device pointers as ints; coefs: (k, 4) Montgomery limbs on the host)."""
    k = len(d_polys)
    c = as_u64(coefs, 4)
    assert c.shape[0] == k == len(lens)
    pp = (C.c_void_p * k)(*[int(x) for x in d_polys])
    ll = (C.c_size_t * k)(*[int(x) for x in lens])
    check(lib().uzkge_cuda_fr_lincomb_device(pp, ll, ptr(c), k, d_out, out_len, stream))


def fr_add_sparse_device(d_poly: int, idx, vals, stream: int = 0) -> None:
    """poly[idx[j]] += vals[j] in order."""
    k = len(idx)
    if k == 0:
        return
    v = as_u64(vals, 4)
    assert v.shape[0] == k
    ii = (C.c_size_t * k)(*[int(x) for x in idx])
    check(lib().uzkge_cuda_fr_add_sparse_device(d_poly, ii, ptr(v), k, stream))


def fr_powers_device(base, n: int, d_out: int, scale=None, stream: int = 0) -> None:
    """out[i] = scale * base^i."""
    b = as_u64(base).reshape(4)
    sc = None if scale is None else as_u64(scale).reshape(4)
    check(lib().uzkge_cuda_fr_powers_device(ptr(b), None if sc is None else ptr(sc), n, d_out, stream))


def fr_gather_device(d_src: int, d_idx_u32: int, n: int, d_out: int, stream: int = 0) -> None:
    check(lib().uzkge_cuda_fr_gather_device(d_src, d_idx_u32, n, d_out, stream))


def fr_mul_device(d_a: int, d_b: int, n: int, d_out: int, stream: int = 0) -> None:
    check(lib().uzkge_cuda_fr_mul_device(d_a, d_b, n, d_out, stream))


def fr_trimmed_len_device(d_poly: int, n: int, stream: int = 0) -> int:
    """Length after FpPolynomial::from_coefs' trim (0 for the zero vector).  Synchronises the stream."""
    v = C.c_size_t(0)
    check(lib().uzkge_cuda_fr_trimmed_len_device(d_poly, n, C.byref(v), stream))
    return int(v.value)


def grand_product_fr_device(d_num: int, d_den: int, n: int, d_out: int, d_tmp: int, stream: int = 0) -> None:
    check(lib().uzkge_cuda_grand_product_fr_device(d_num, d_den, n, d_out, d_tmp, stream))


def plonk_z_evals_fr_device(d_w, d_sigma, d_group: int, k, beta, gamma, n: int, d_z: int, d_tmp: int, stream: int = 0) -> None:
    ww = (C.c_void_p * 5)(*[int(x) for x in d_w])
    ss = (C.c_void_p * 5)(*[int(x) for x in d_sigma])
    kk = as_u64(k, 4)
    assert kk.shape[0] == 5
    check(lib().uzkge_cuda_plonk_z_evals_fr_device(ww, ss, d_group, ptr(kk), ptr(as_u64(beta).reshape(4)), ptr(as_u64(gamma).reshape(4)),
                                                   n, d_z, d_tmp, stream))


def g1_sum_device(d_parts: int, count: int, d_out: int, stream: int = 0, k: int = 1, stride: int | None = None) -> None:
    """d_out[j] = sum_r d_parts[r * stride + j], j < k (device pointers to Jacobian points): the combine of gathered per-GPU partial sums."""
    check(lib().uzkge_cuda_g1_sum_device(d_parts, count, k if stride is None else stride, k, d_out, stream), CommitmentError)


def g1_add(a_jac, b_jac) -> np.ndarray:
    a, b = as_u64(a_jac).reshape(12), as_u64(b_jac).reshape(12)
    out = np.zeros(12, dtype=np.uint64)
    check(lib().uzkge_cuda_g1_add(ptr(a), ptr(b), ptr(out)), CommitmentError)
    return out


def g1_to_affine(jac) -> np.ndarray:
    a = as_u64(jac).reshape(12)
    out = np.zeros(8, dtype=np.uint64)
    check(lib().uzkge_cuda_g1_to_affine(ptr(a), ptr(out)), CommitmentError)
    return out


def field_mul(a, b, field: str = "fr") -> np.ndarray:
    a, b = as_u64(a, 4), as_u64(b, 4)
    assert a.shape == b.shape
    out = np.empty_like(a)
    check(lib().uzkge_cuda_field_mul(0 if field == "fr" else 1, ptr(a), ptr(b), ptr(out), a.shape[0]))
    return out


def bench_field_mul(field: str = "fr", iters: int = 2000) -> float:
    v = C.c_double(0)
    check(lib().uzkge_cuda_bench_field_mul(0 if field == "fr" else 1, iters, C.byref(v)))
    return float(v.value)


class PinnedArray:
    """A page-locked uint64 host buffer from uzkge_cuda_host_alloc, viewed as a numpy array."""

    def __init__(self, shape):
        self.shape = tuple(shape)
        n = int(np.prod(self.shape))
        p = C.c_void_p()
        check(lib().uzkge_cuda_host_alloc(n * 8, C.byref(p)))
        self._p = p
        self.array = np.ctypeslib.as_array(C.cast(p, u64p), shape=(n,)).reshape(self.shape)

    def free(self):
        if self._p:
            lib().uzkge_cuda_host_free(self._p)
            self._p = None
            self.array = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
