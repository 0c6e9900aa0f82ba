// ntt_plan.h -- decomposition of one Fr transform into tile passes.
//
// Contract (what the passes must add up to): out[i] = sum_j in[j] * w_N^(i*j), natural order on both sides,
// input zero-padded from len_in to N -- FpPolynomial::fft_with_domain / ifft_with_domain,
// /root/reference/uzkge/src/poly_commit/field_polynomial.rs:583-597, N = 2^b (Radix2EvaluationDomain) or
// 3 * 2^b (MixedRadixEvaluationDomain, :561-567).
//
// Decomposition (Cooley-Tukey / "four-step", no bit-reversal pass, no transposes in HBM):
//   N = 3^a * M,  M = N_0 * N_1 * ... * N_{P-1}  (P <= 3 power-of-two factors, each <= one shared-memory tile)
//   a = 1: a radix-3 pre-pass computes y[k*M + n] = w_N^(n*k) * sum_{m<3} x[m*M + n] * w_3^(m*k); the three rows
//          are then independent size-M transforms with root w_N^3 whose outputs interleave (out[k + 3*i]).
//   pass i < P-1: sub-transforms of size N_i over the slowest remaining input digit, columns are contiguous in
//          memory; results are multiplied by the inter-pass twiddle w_M^(A_i * k_i * n_rest) and written in place
//          (A_i = N_0*...*N_{i-1}, n_rest = index of the untransformed digits).
//   last pass: sub-transforms over the contiguous digit; the tile holds C adjacent k_0 values so that the
//          digit-reversed output index sum_j k_j * A_j is written in runs of C contiguous elements.
//   inverse: same passes with the output index i -> (N - i) mod N and a factor 1/N  (w^-1 = conj. index).
//
// The index helpers below are shared by the CUDA kernels (ntt.cu) and by the host simulator in tests/host.
#pragma once
#include <stdint.h>

#include "ff.cuh"

namespace uz {

struct NttPass {
    uint32_t logR, logC;    // tile: R rows (the transformed digit) x C columns
    uint32_t inner_tiles;   // columns / C
    uint32_t outer;         // outer repetitions
    uint32_t batch;         // 1, or 3 rows for mixed radix
    uint64_t in_rs, in_cs, in_os, in_bs;
    uint64_t out_rs, out_cs, out_os, out_bs;
    uint32_t in_r_contig;   // load mapping: rows fastest (in_rs == 1) or columns fastest
    uint32_t stage_stride;  // stage-twiddle table stride: w_R^j = stage_tab[j * stage_stride]
    uint64_t tw_mul;        // inter-pass twiddle exponent = k * col * tw_mul (0: none)
    uint32_t first, last;
};

struct NttPlan {
    uint64_t n;        // domain size
    uint64_t m;        // power-of-two part
    uint32_t logm;
    uint32_t mixed;    // 1 if n == 3 * m
    uint32_t npass;
    NttPass pass[3];
};

constexpr uint32_t NTT_LOG_TWLO = 12;  // two-level power tables: x^e = hi[e >> 12] * lo[e & 4095]

UZ_HD uint64_t ntt_in_index(const NttPass& p, uint32_t b, uint32_t o, uint32_t t, uint32_t r, uint32_t c) {
    return (uint64_t)b * p.in_bs + (uint64_t)o * p.in_os + ((uint64_t)t * (1u << p.logC) + c) * p.in_cs + (uint64_t)r * p.in_rs;
}
UZ_HD uint64_t ntt_out_index(const NttPass& p, uint32_t b, uint32_t o, uint32_t t, uint32_t k, uint32_t c) {
    return (uint64_t)b * p.out_bs + (uint64_t)o * p.out_os + ((uint64_t)t * (1u << p.logC) + c) * p.out_cs + (uint64_t)k * p.out_rs;
}
UZ_HD uint64_t ntt_tw_exponent(const NttPass& p, uint32_t t, uint32_t k, uint32_t c) {
    return (uint64_t)k * ((uint64_t)t * (1u << p.logC) + c) * p.tw_mul;
}
UZ_HD uint32_t ntt_bitrev(uint32_t x, uint32_t bits) {
#if defined(__CUDA_ARCH__)
    return bits ? (__brev(x) >> (32 - bits)) : 0;
#else
    uint32_t r = 0;
    for (uint32_t i = 0; i < bits; i++) {
        r = (r << 1) | (x & 1);
        x >>= 1;
    }
    return r;
#endif
}

// max_log_tile: log2 of the elements one CTA can stage (shared memory); max_log_r: largest sub-transform.
// Returns false if the size is unsupported.
inline bool ntt_make_plan(uint64_t n, uint32_t max_log_tile, uint32_t max_log_r, uint32_t two_pass_max_log, NttPlan* plan) {
    if (n == 0) return false;
    uint64_t m = n;
    uint32_t mixed = 0;
    if (m % 3 == 0) {
        m /= 3;
        mixed = 1;
    }
    if (m & (m - 1)) return false;
    uint32_t logm = 0;
    while ((1ull << logm) < m) logm++;
    if (logm > 28) return false;  // Fr two-adicity
    if (max_log_r > max_log_tile) max_log_r = max_log_tile;
    uint32_t np;
    if (logm <= max_log_r && logm <= 10)
        np = 1;
    else if (logm <= two_pass_max_log && (logm + 1) / 2 <= max_log_r)
        np = 2;
    else
        np = 3;
    if ((logm + np - 1) / np > max_log_r) return false;
    plan->n = n;
    plan->m = m;
    plan->logm = logm;
    plan->mixed = mixed;
    plan->npass = np;
    uint32_t lg[3] = {0, 0, 0};
    for (uint32_t i = 0; i < np; i++) lg[i] = logm / np + (i < logm % np ? 1 : 0);
    const uint64_t scale = n / m;  // w_M = w_N^scale
    const uint32_t log_rtab = logm < 12 ? logm : 12;
    uint64_t A = 1;
    for (uint32_t i = 0; i < np; i++) {
        NttPass& p = plan->pass[i];
        p.logR = lg[i];
        p.batch = mixed ? 3 : 1;
        p.first = (i == 0);
        p.last = (i == np - 1);
        p.stage_stride = 1u << (log_rtab - (lg[i] < log_rtab ? lg[i] : log_rtab));
        uint64_t Z = 1;
        for (uint32_t j = i + 1; j < np; j++) Z <<= lg[j];
        if (!p.last) {
            uint32_t logC = max_log_tile - p.logR;
            uint32_t logZ = 0;
            while ((1ull << logZ) < Z) logZ++;
            if (logC > logZ) logC = logZ;
            if (logC > 4) logC = 4;
            p.logC = logC;
            p.inner_tiles = (uint32_t)(Z >> logC);
            p.outer = (uint32_t)A;
            p.in_rs = Z;
            p.in_cs = 1;
            p.in_os = (Z << p.logR);
            p.in_bs = m;
            p.out_rs = p.in_rs;
            p.out_cs = p.in_cs;
            p.out_os = p.in_os;
            p.out_bs = m;
            p.in_r_contig = 0;
            p.tw_mul = A * scale;
        } else {
            const uint64_t mul = mixed ? 3 : 1;
            p.tw_mul = 0;
            p.in_rs = 1;
            p.in_r_contig = 1;
            p.in_bs = m;
            p.out_bs = mixed ? 1 : 0;
            if (np == 1) {
                p.logC = 0;
                p.inner_tiles = 1;
                p.outer = 1;
                p.in_cs = 0;
                p.in_os = 0;
                p.out_rs = mul;
                p.out_cs = 0;
                p.out_os = 0;
            } else {
                uint32_t logC = max_log_tile - p.logR;
                if (logC > lg[0]) logC = lg[0];
                if (logC > 3) logC = 3;
                p.logC = logC;
                p.inner_tiles = (1u << lg[0]) >> logC;   // columns are k_0
                p.in_cs = m >> lg[0];                     // stride of k_0 in the partially transformed array
                p.out_cs = mul;                           // A_0 = 1
                p.out_rs = A * mul;                       // A_{P-1}
                if (np == 3) {
                    p.outer = 1u << lg[1];
                    p.in_os = 1ull << lg[2];
                    p.out_os = (1ull << lg[0]) * mul;
                } else {
                    p.outer = 1;
                    p.in_os = 0;
                    p.out_os = 0;
                }
            }
        }
        A <<= lg[i];
    }
    return true;
}

// LARGE = 5^((r-1)/(2^28 * 9)) in Montgomery form: generator of the 2^28 * 3^2 subgroup;
// w_N = LARGE^(3^(2-a) * 2^(28-b)) for N = 3^a 2^b.  Pinned by tests/golden/domain_kat.json.
inline fe fr_large_subgroup_root() {
    fe g = fe_zero();
    g.l[0] = 5;
    g = fe_to_mont<FrP>(g);
    // exponent (r - 1) / (2^28 * 9), little-endian 32-bit limbs
    static const uint32_t e[8] = {0x2358d107u, 0x9f828f3bu, 0xd5b19cf2u, 0x58026433u,
                                  0x2b395f7du, 0xacca004au, 0x5607a7e8u, 0x00000000u};
    fe acc = fe_one<FrP>();
    for (int i = 255; i >= 0; i--) {
        acc = fe_sqr<FrP>(acc);
        if ((e[i >> 5] >> (i & 31)) & 1) acc = fe_mul<FrP>(acc, g);
    }
    return acc;
}

inline fe ntt_root_of_unity(uint64_t n, bool* ok) {
    uint32_t a = 0, b = 0;
    uint64_t m = n;
    *ok = false;
    if (n == 0) return fe_zero();
    while (m % 3 == 0) { m /= 3; a++; }
    while (m % 2 == 0) { m /= 2; b++; }
    if (m != 1 || a > 2 || b > 28) return fe_zero();
    static const fe large = fr_large_subgroup_root();
    fe w = large;
    for (uint32_t i = 0; i < 2 - a; i++) w = fe_mul<FrP>(fe_sqr<FrP>(w), w);
    for (uint32_t i = 0; i < 28 - b; i++) w = fe_sqr<FrP>(w);
    *ok = true;
    return w;
}


}  // namespace uz
