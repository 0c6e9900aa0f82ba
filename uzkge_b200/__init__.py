"""uzkge_b200 -- B200-native (sm_100a) MSM / NTT proving backend for zypher-game/uzkge.

Layout: `csrc/` CUDA kernels + the C ABI (include/uzkge_cuda.h), `ffi` the ctypes binding of that ABI,
`poly_commit` the host-side mirror of the reference's FpPolynomial / KZG commit interface, `plonk` (+ `transcript`, `rng`) the
device-resident TurboPlonK indexer / prover mirroring the reference's `plonk` module (imported on demand: it needs torch),
`dist` the one-process-per-GPU sharding (split MSMs, the prover's commitment / transform service, four-step NTT with NCCL or
fused over peer memory).  There is no CPU fallback anywhere in this package.
"""
from . import errors, ffi  # noqa: F401
from .errors import BackendUnavailable, CommitmentError, DegreeError, FFTError, UzkgeError  # noqa: F401
from .poly_commit import (  # noqa: F401
    FpPolynomial,
    KZGCommitment,
    KZGCommitmentSchemeBN254,
    MixedRadixEvaluationDomain,
    Radix2EvaluationDomain,
)

__version__ = "0.1.0"
