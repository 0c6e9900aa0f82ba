"""The prover's randomness source, host side: `rand_chacha::ChaChaRng` (ChaCha20, 64-bit block counter, stream 0) and
arkworks' `Fr::rand` on top of it.

  ChaChaRng::from_seed([0u8; 32])   /root/reference/uzkge/src/plonk/indexer.rs:258 (choose_ks), every test's prover RNG
  F::rand(prng)                     /root/reference/uzkge/src/plonk/helpers.rs:147 (hide_polynomial), :1349 (split_t_and_commit),
                                    plonk/indexer.rs:224 (choose_ks)

`Fr::rand` (ark-ff 0.4, `Distribution<Fp> for Standard`): draw 4 x next_u64 as the RAW Montgomery limbs (little-endian limb
order), clear the top 64 * 4 - 254 = 2 bits of the last limb, reject and redraw while the value is >= r.  The golden k[1..4] of the
reference's verifier keys (tests/golden/domain_kat.json, produced by choose_ks with this generator at seed 0) pin this.
"""
from __future__ import annotations

FR_MODULUS = 21888242871839275222246405745257275088548364400416034343698204186575808495617
_R_INV = pow(1 << 256, -1, FR_MODULUS)
_M32 = 0xFFFFFFFF


def _rotl(v: int, r: int) -> int:
    return ((v << r) | (v >> (32 - r))) & _M32


def _quarter(s, a, b, c, d):
    s[a] = (s[a] + s[b]) & _M32; s[d] = _rotl(s[d] ^ s[a], 16)
    s[c] = (s[c] + s[d]) & _M32; s[b] = _rotl(s[b] ^ s[c], 12)
    s[a] = (s[a] + s[b]) & _M32; s[d] = _rotl(s[d] ^ s[a], 8)
    s[c] = (s[c] + s[d]) & _M32; s[b] = _rotl(s[b] ^ s[c], 7)


class ChaChaRng:
    """ChaCha20 keystream as a sequence of little-endian u32 words; next_u64 = low word then high word (rand_core BlockRng)."""

    def __init__(self, seed: bytes = b"\x00" * 32, stream: int = 0):
        assert len(seed) == 32
        self._key = [int.from_bytes(seed[4 * i: 4 * i + 4], "little") for i in range(8)]
        self._stream = stream
        self._counter = 0
        self._buf: list[int] = []
        self._pos = 0

    @classmethod
    def from_seed(cls, seed: bytes) -> "ChaChaRng":
        return cls(seed)

    def _block(self) -> list[int]:
        import ctypes as C

        from .transcript import host_lib

        key = (C.c_uint32 * 8)(*self._key)
        out = (C.c_uint32 * 16)()
        host_lib().uzkge_host_chacha20_block(key, self._counter, self._stream, out)
        self._counter += 1
        return list(out)

    def _block_py(self) -> list[int]:
        """Pure-Python statement of uzkge_host_chacha20_block (csrc/hostutil.c); the tests compare them."""
        init = [0x61707865, 0x3320646E, 0x79622D32, 0x6B206574] + self._key + [
            self._counter & _M32, (self._counter >> 32) & _M32, self._stream & _M32, (self._stream >> 32) & _M32]
        s = list(init)
        for _ in range(10):
            _quarter(s, 0, 4, 8, 12); _quarter(s, 1, 5, 9, 13); _quarter(s, 2, 6, 10, 14); _quarter(s, 3, 7, 11, 15)
            _quarter(s, 0, 5, 10, 15); _quarter(s, 1, 6, 11, 12); _quarter(s, 2, 7, 8, 13); _quarter(s, 3, 4, 9, 14)
        self._counter += 1
        return [(s[i] + init[i]) & _M32 for i in range(16)]

    def next_u32(self) -> int:
        if self._pos >= len(self._buf):
            self._buf = self._block()
            self._pos = 0
        v = self._buf[self._pos]
        self._pos += 1
        return v

    def next_u64(self) -> int:
        lo = self.next_u32()
        return lo | (self.next_u32() << 32)


def fr_rand_mont(prng) -> int:
    """The raw (Montgomery) 256-bit value arkworks stores for `Fr::rand(prng)`."""
    while True:
        limbs = [prng.next_u64() for _ in range(4)]
        limbs[3] &= (1 << 62) - 1
        raw = limbs[0] | (limbs[1] << 64) | (limbs[2] << 128) | (limbs[3] << 192)
        if raw < FR_MODULUS:
            return raw


def fr_rand(prng) -> int:
    """`Fr::rand(prng)` as a canonical integer."""
    return fr_rand_mont(prng) * _R_INV % FR_MODULUS


def choose_ks(prng, n_wires_per_gate: int) -> list[int]:
    """plonk/indexer.rs:211-235: k[0] = 1, then distinct quadratic non-residues drawn from `prng`."""
    k = [1]
    exp = (FR_MODULUS - 1) // 2
    while len(k) < n_wires_per_gate:
        ki = fr_rand(prng)
        if ki == 0:
            continue
        if ki not in k and pow(ki, exp, FR_MODULUS) != 1:
            k.append(ki)
    return k
