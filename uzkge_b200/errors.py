"""Error type of the host layer: the variants of `UzkgeError` (/root/reference/uzkge/src/errors.rs:6-45) that the
hot path can raise, plus the backend failure the Rust wrapper maps CUDA errors to."""
from __future__ import annotations


class UzkgeError(Exception):
    """Base class; `kind` is the reference's enum variant name."""

    kind = "Message"


class DegreeError(UzkgeError):
    """PolyComScheme: the degree of the polynomial is higher than the maximum supported
    (kzg_poly_commitment.rs:283-285)."""

    kind = "DegreeError"


class CommitmentError(UzkgeError):
    """Plonk: commitment error -- what a failing `uzkge_cuda_msm_g1*` call maps to."""

    kind = "CommitmentError"


class FFTError(UzkgeError):
    """Plonk: FFT error -- what a failing `uzkge_cuda_ntt_fr*` call maps to."""

    kind = "FFTError"


class ParameterError(UzkgeError):
    kind = "ParameterError"


class BackendUnavailable(UzkgeError):
    """The CUDA library is missing or no sm_100 device is usable.  There is no CPU fallback."""

    kind = "Message"
