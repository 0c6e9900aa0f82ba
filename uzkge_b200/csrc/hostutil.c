/* hostutil.c -- host-side primitives of the Fiat-Shamir transcript and of the prover's RNG, for the Python mirror of the
 * reference's prover (uzkge_b200/plonk.py).  In production these stay in Rust (`sha3::Keccak256`, `rand_chacha::ChaChaRng`);
 * here they are C so that the serial host part of a small proof does not dominate it.  Built with gcc into
 * uzkge_b200/lib/libuzkge_host.so; no CUDA.
 *
 *   uzkge_host_keccak256      Keccak-256 with the original 0x01 padding   (utils/transcript.rs:60-62)
 *   uzkge_host_chacha20_block one 64-byte ChaCha20 block, 64-bit counter  (rand_chacha 0.3 `ChaCha20Rng`, stream 0)
 *   uzkge_host_fr_mont_to_be  n Montgomery Fr elements -> canonical 32-byte big-endian strings, the form in which public inputs
 *                             enter the transcript (`into_bigint().to_bytes_be()`, plonk/transcript.rs:27-30)
 */
#include <stddef.h>
#include <stdint.h>
#include <string.h>

#define API __attribute__((visibility("default")))

static const uint64_t RC[24] = {
    0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808AULL, 0x8000000080008000ULL, 0x000000000000808BULL,
    0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL, 0x000000000000008AULL, 0x0000000000000088ULL,
    0x0000000080008009ULL, 0x000000008000000AULL, 0x000000008000808BULL, 0x800000000000008BULL, 0x8000000000008089ULL,
    0x8000000000008003ULL, 0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800AULL, 0x800000008000000AULL,
    0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};

static inline uint64_t rol64(uint64_t v, int r) { return r ? (v << r) | (v >> (64 - r)) : v; }

static void keccak_f(uint64_t a[25]) {
    uint64_t b[25], c[5], d[5];
    for (int round = 0; round < 24; round++) {
        for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
        for (int x = 0; x < 5; x++)
            for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(a[x + 5 * y], ROT[x + 5 * y]);
        for (int y = 0; y < 25; y += 5)
            for (int x = 0; x < 5; x++) a[y + x] = b[y + x] ^ (~b[y + (x + 1) % 5] & b[y + (x + 2) % 5]);
        a[0] ^= RC[round];
    }
}

API void uzkge_host_keccak256(const uint8_t* data, size_t len, uint8_t out[32]) {
    enum { RATE = 136 };
    uint64_t a[25];
    uint8_t block[RATE];
    memset(a, 0, sizeof a);
    size_t off = 0;
    for (;;) {
        size_t take = len - off < RATE ? len - off : RATE;
        int last = take < RATE;
        memset(block, 0, RATE);
        if (take) memcpy(block, data + off, take);
        if (last) {
            block[take] ^= 0x01;
            block[RATE - 1] ^= 0x80;
        }
        for (int i = 0; i < RATE / 8; i++) {
            uint64_t lane = 0;
            for (int k = 7; k >= 0; k--) lane = (lane << 8) | block[8 * i + k];
            a[i] ^= lane;
        }
        keccak_f(a);
        off += take;
        if (last) break;
    }
    for (int i = 0; i < 4; i++)
        for (int k = 0; k < 8; k++) out[8 * i + k] = (uint8_t)(a[i] >> (8 * k));
}

static inline uint32_t rol32(uint32_t v, int r) { return (v << r) | (v >> (32 - r)); }
#define QR(a, b, c, d)                 \
    do {                               \
        a += b; d = rol32(d ^ a, 16);  \
        c += d; b = rol32(b ^ c, 12);  \
        a += b; d = rol32(d ^ a, 8);   \
        c += d; b = rol32(b ^ c, 7);   \
    } while (0)

API void uzkge_host_chacha20_block(const uint32_t key[8], uint64_t counter, uint64_t stream, uint32_t out[16]) {
    uint32_t in[16] = {0x61707865u, 0x3320646Eu, 0x79622D32u, 0x6B206574u, key[0], key[1], key[2], key[3], key[4], key[5], key[6], key[7],
                       (uint32_t)counter, (uint32_t)(counter >> 32), (uint32_t)stream, (uint32_t)(stream >> 32)};
    uint32_t s[16];
    memcpy(s, in, sizeof s);
    for (int i = 0; i < 10; i++) {
        QR(s[0], s[4], s[8], s[12]); QR(s[1], s[5], s[9], s[13]); QR(s[2], s[6], s[10], s[14]); QR(s[3], s[7], s[11], s[15]);
        QR(s[0], s[5], s[10], s[15]); QR(s[1], s[6], s[11], s[12]); QR(s[2], s[7], s[8], s[13]); QR(s[3], s[4], s[9], s[14]);
    }
    for (int i = 0; i < 16; i++) out[i] = s[i] + in[i];
}

/* ---- BN254 Fr out of Montgomery form: one REDC pass per element (multiplication by 1) */
typedef unsigned __int128 u128;
static const uint64_t FR_P[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
static const uint64_t FR_INV = 0xc2e1f593efffffffULL;   /* -p^-1 mod 2^64 */

API void uzkge_host_fr_mont_to_be(const uint64_t* limbs, size_t n, uint8_t* out) {
    for (size_t e = 0; e < n; e++) {
        uint64_t t[5] = {limbs[4 * e], limbs[4 * e + 1], limbs[4 * e + 2], limbs[4 * e + 3], 0};
        for (int i = 0; i < 4; i++) {
            const uint64_t m = t[0] * FR_INV;
            u128 c = (u128)m * FR_P[0] + t[0];
            c >>= 64;
            for (int j = 1; j < 4; j++) {
                c += (u128)m * FR_P[j] + t[j];
                t[j - 1] = (uint64_t)c;
                c >>= 64;
            }
            c += t[4];
            t[3] = (uint64_t)c;
            t[4] = (uint64_t)(c >> 64);
        }
        /* t < 2p: one conditional subtraction */
        uint64_t d[4];
        u128 b = 0;
        int ge = 1;
        for (int j = 3; j >= 0; j--)
            if (t[j] != FR_P[j]) { ge = t[j] > FR_P[j]; break; }
        if (t[4] || ge) {
            for (int j = 0; j < 4; j++) {
                const u128 x = (u128)t[j] - FR_P[j] - (uint64_t)b;
                d[j] = (uint64_t)x;
                b = (x >> 64) & 1;
            }
        } else {
            memcpy(d, t, sizeof d);
        }
        for (int j = 0; j < 4; j++)
            for (int k = 0; k < 8; k++) out[32 * e + 8 * (3 - j) + (7 - k)] = (uint8_t)(d[j] >> (8 * k));
    }
}
