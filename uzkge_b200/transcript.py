"""Fiat-Shamir transcript of the reference, host side (serial, O(1) per round: SURVEY 8e "replicas only").

  Transcript               /root/reference/uzkge/src/utils/transcript.rs:8-69   (Keccak-256 over a byte state of 32-byte slots;
                           a challenge is the digest read as a big-endian integer mod r, and becomes the new state)
  transcript_init_plonk    /root/reference/uzkge/src/plonk/transcript.rs:8-31
  init_pcs_batch_eval...   /root/reference/uzkge/src/poly_commit/pcs.rs:220-240

Keccak-256 is the original Keccak padding (0x01), as in the `sha3` crate's `Keccak256` -- not NIST SHA3-256, so `hashlib`
does not provide it.  Field elements here are canonical Python integers.
"""
from __future__ import annotations

FR_MODULUS = 21888242871839275222246405745257275088548364400416034343698204186575808495617
SLOT_SIZE = 32

_MASK = (1 << 64) - 1
_RC = [
    0x0000000000000001, 0x0000000000008082, 0x800000000000808A, 0x8000000080008000, 0x000000000000808B, 0x0000000080000001,
    0x8000000080008081, 0x8000000000008009, 0x000000000000008A, 0x0000000000000088, 0x0000000080008009, 0x000000008000000A,
    0x000000008000808B, 0x800000000000008B, 0x8000000000008089, 0x8000000000008003, 0x8000000000008002, 0x8000000000000080,
    0x000000000000800A, 0x800000008000000A, 0x8000000080008081, 0x8000000000008080, 0x0000000080000001, 0x8000000080008008,
]
# rotation offsets r[x + 5 y]
_ROT = [0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14]


def _rol(v: int, r: int) -> int:
    return ((v << r) | (v >> (64 - r))) & _MASK if r else v


def _keccak_f(a: list[int]) -> None:
    for rc in _RC:
        c = [a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20] for x in range(5)]
        d = [c[(x + 4) % 5] ^ _rol(c[(x + 1) % 5], 1) for x in range(5)]
        for i in range(25):
            a[i] ^= d[i % 5]
        b = [0] * 25
        for x in range(5):
            for y in range(5):
                b[y + 5 * ((2 * x + 3 * y) % 5)] = _rol(a[x + 5 * y], _ROT[x + 5 * y])
        for y in range(0, 25, 5):
            for x in range(5):
                a[y + x] = b[y + x] ^ (~b[y + (x + 1) % 5] & _MASK & b[y + (x + 2) % 5])
        a[0] ^= rc


def keccak256_py(data: bytes) -> bytes:
    """Pure-Python Keccak-256: the readable statement of what uzkge_host_keccak256 (csrc/hostutil.c) computes; the tests compare them."""
    rate = 136
    msg = bytearray(data)
    msg.append(0x01)
    msg.extend(b"\x00" * (-len(msg) % rate))
    msg[-1] |= 0x80
    a = [0] * 25
    for off in range(0, len(msg), rate):
        for i in range(rate // 8):
            a[i] ^= int.from_bytes(msg[off + 8 * i: off + 8 * i + 8], "little")
        _keccak_f(a)
    return b"".join(a[i].to_bytes(8, "little") for i in range(4))


_host = None


def host_lib():
    """libuzkge_host.so (csrc/hostutil.c, gcc): Keccak-256 and the ChaCha20 block, so the serial host part of a small proof stays small."""
    global _host
    if _host is None:
        import ctypes as C
        import os

        path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "lib", "libuzkge_host.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is missing: build it with `python -m uzkge_b200.build`")
        L = C.CDLL(path)
        L.uzkge_host_keccak256.restype = None
        L.uzkge_host_keccak256.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
        L.uzkge_host_chacha20_block.restype = None
        L.uzkge_host_chacha20_block.argtypes = [C.POINTER(C.c_uint32), C.c_uint64, C.c_uint64, C.POINTER(C.c_uint32)]
        L.uzkge_host_fr_mont_to_be.restype = None
        L.uzkge_host_fr_mont_to_be.argtypes = [C.c_void_p, C.c_size_t, C.c_char_p]
        _host = L
    return _host


def keccak256(data: bytes) -> bytes:
    import ctypes as C

    out = C.create_string_buffer(32)
    data = bytes(data)
    host_lib().uzkge_host_keccak256(data, len(data), out)
    return out.raw


def fr_mont_rows_to_bytes_be(rows) -> bytes:
    """(k, 4) uint64 Montgomery limbs -> k canonical 32-byte big-endian strings, concatenated (one C call: a circuit's hundreds of
    public inputs enter the transcript in this form, plonk/transcript.rs:27-30)."""
    import ctypes as C

    import numpy as np

    a = np.ascontiguousarray(rows, dtype=np.uint64).reshape(-1, 4)
    out = C.create_string_buffer(32 * a.shape[0]) if a.shape[0] else None
    if a.shape[0] == 0:
        return b""
    host_lib().uzkge_host_fr_mont_to_be(a.ctypes.data, a.shape[0], out)
    return out.raw


def fr_to_bytes_be(x: int) -> bytes:
    """`into_bigint().to_bytes_be()` of a BN254 Fr element: 32 bytes."""
    return int(x).to_bytes(32, "big")


class Transcript:
    """utils/transcript.rs:8-69.  Labels are ignored by the reference ("omitted for efficiency") and are not taken here."""

    def __init__(self, msg: bytes):
        self.state = bytearray()
        self.append_message(msg)

    def append_message(self, msg: bytes) -> None:
        if len(msg) < SLOT_SIZE:
            self.state.extend(b"\x00" * (SLOT_SIZE - len(msg)) + bytes(msg))
        else:
            assert len(msg) % SLOT_SIZE == 0
            self.state.extend(msg)

    def append_u64(self, a: int) -> None:
        self.state.extend(b"\x00" * (SLOT_SIZE - 8) + int(a).to_bytes(8, "big"))

    def append_single_byte(self, b: int) -> None:
        self.state.append(b)

    def append_commitment(self, comm) -> None:
        """`comm.to_transcript_bytes()`: affine x BE || y BE, 64 zero bytes for the identity (kzg_poly_commitment.rs:37-53)."""
        self.append_message(comm.to_transcript_bytes())

    def append_challenge(self, challenge: int) -> None:
        self.append_message(fr_to_bytes_be(challenge))

    def get_challenge_field_elem(self) -> int:
        digest = keccak256(bytes(self.state))
        challenge = int.from_bytes(digest, "big") % FR_MODULUS   # buf.reverse(); from_le_bytes_mod_order
        self.state = bytearray(fr_to_bytes_be(challenge))
        return challenge


def transcript_init_plonk(transcript: Transcript, params, pi_values, root: int) -> None:
    """plonk/transcript.rs:8-31."""
    transcript.append_message(b"PLONK")
    transcript.append_u64(params.cs_size)
    transcript.append_message(fr_to_bytes_be(FR_MODULUS))
    for q in params.cm_q_vec:
        transcript.append_commitment(q)
    for p in params.cm_s_vec:
        transcript.append_commitment(p)
    transcript.append_challenge(root)
    for k in params.k:
        transcript.append_challenge(k)
    for v in pi_values:
        transcript.append_challenge(v)


def init_pcs_batch_eval_transcript(transcript: Transcript, max_degree: int, point: int) -> None:
    """poly_commit/pcs.rs:220-240."""
    transcript.append_message(b"New PCS-Batch-Eval Protocol")
    transcript.append_message(fr_to_bytes_be(FR_MODULUS))
    transcript.append_u64(max_degree)
    transcript.append_challenge(point)
