// ntt.cu -- Fr NTT / iNTT / coset-FFT kernels (radix-2 and 3 * 2^k domains) for sm_100a.
//
// Replaces domain.fft / domain.ifft behind FpPolynomial::{fft,ifft,coset_fft,coset_ifft}_with_domain
// (/root/reference/uzkge/src/poly_commit/field_polynomial.rs:583-607) and the serial coset scaling of
// mul_var_assign (:470-477).  See ntt_plan.h for the decomposition.
//
// One CTA transforms one R x C tile staged in shared memory: the tile is loaded with 128-bit accesses,
// log2(R) decimation-in-frequency butterfly stages run on it, and it is written back with the
// digit-reversed row index, the inter-pass twiddle, and (last pass) the inverse-transform index flip, the
// 1/N factor and the coset post-scale all fused into the store.  Twiddles come from small HBM/L2-resident
// tables: w_R^j for the butterfly stages and a two-level table x^e = hi[e >> 12] * lo[e & 4095] for the
// inter-pass twiddles and the coset powers.
//
// Tried and rejected on B200 (measured): running three DIF stages per shared-memory round trip in registers
// (radix-8 items, 156 registers, 128-thread CTAs) -- 1159 us for 2^22 against 1065 us for this radix-2 version with
// 58 registers and 4x the resident warps: for these carry-chain kernels thread-level parallelism hides the
// multiplier latency better than instruction-level parallelism does.
#include <cuda_runtime.h>

#include <cstring>

#include "devmem.cuh"
#include "internal.h"
#include "ntt_plan.h"

namespace uz {

// shared-memory tile: two 16-byte planes so that a warp's 128-bit accesses to consecutive elements are
// conflict-free; columns are rotated by the row index so that row-major and column-major walks both spread
// over the banks.
struct TileView {
    uint4* lo;
    uint4* hi;
    uint32_t logC, cmask;
    // A quarter-warp (8 lanes x 16 bytes = one 128-byte wavefront) must hit 8 distinct slots mod 8.  With C >= 8 it stays inside
    // one row and the rotation by r suffices.  With C = 4 it spans two rows, which the stage / store walks pick 2h or R/2 apart
    // (same low row bit): bit 2 of the slot is therefore the low row bit XOR the parity of the remaining row bits, so that two
    // rows differing in exactly one bit never share a half.
    __device__ __forceinline__ uint32_t pos(uint32_t r, uint32_t c) const {
        uint32_t p = (r << logC) | ((c + r) & cmask);
        if (logC == 2) p ^= (__popc(r >> 1) & 1u) << 2;
        return p;
    }
    __device__ __forceinline__ fe ld(uint32_t p) const {
        fe x;
        uint4 a = lo[p], b = hi[p];
        x.l[0] = a.x; x.l[1] = a.y; x.l[2] = a.z; x.l[3] = a.w;
        x.l[4] = b.x; x.l[5] = b.y; x.l[6] = b.z; x.l[7] = b.w;
        return x;
    }
    __device__ __forceinline__ void st(uint32_t p, const fe& x) const {
        lo[p] = make_uint4(x.l[0], x.l[1], x.l[2], x.l[3]);
        hi[p] = make_uint4(x.l[4], x.l[5], x.l[6], x.l[7]);
    }
};

__device__ __forceinline__ fe pow2l(const fe* lo, const fe* hi, uint64_t e) {
    fe a = ldg_fe(lo + (e & ((1u << NTT_LOG_TWLO) - 1)));
    uint64_t h = e >> NTT_LOG_TWLO;
    if (h == 0) return a;
    return fe_mul<FrP>(a, ldg_fe(hi + h));
}

// up to NTT_MAX_BATCH independent vectors of one domain travel through a pass in ONE launch: blockIdx.y is the vector
struct NttJobs {
    const fe* in[NTT_MAX_BATCH];
    fe* out[NTT_MAX_BATCH];
    uint64_t len_in[NTT_MAX_BATCH];
};
struct NttKernelArgs {
    NttPass p;
    NttJobs jobs;
    uint64_t n;
    const fe* stage_tab;
    const fe* w_lo;
    const fe* w_hi;
    const fe* g_lo;  // coset powers (forward: g^j, inverse: g^j / N) or null
    const fe* g_hi;
    const fe* g_full;   // g_full[j] = the same powers, one entry per element (resident in HBM)
    const fe* tw_pass;  // inter-pass twiddles of this pass, one entry per output position (resident in HBM)
    fe scale;        // 1/N (inverse without coset)
    uint32_t inverse, pre_coset, post_coset, has_scale, zero_pad;
    // SCATTER instantiation only (the local transform of a distributed four-step, rank k1 of G = 2^scatter_log_g): the last pass stores
    // output k2 of this size-L transform -- X[k1 + G k2] of the whole transform -- straight into its owner's natural slice over peer
    // memory: rank k2 >> scatter_shift (scatter_shift = log2(L / G)), position k1 + G (k2 mod L / G).  No third exchange.
    fe* scatter_rows[8];
    uint32_t scatter_log_g, scatter_shift, scatter_k1;
};

// R4: two butterfly stages per shared-memory round trip (radix-4 items held in registers): half the LDS / STS / index arithmetic and
// half the barriers of the radix-2 walk, the same products.  Held to 64 registers (4 resident CTAs per SM; at its natural 80 registers
// and 3 CTAs it only ties the radix-2 walk).  Measured: 2^22 883 -> 864 us, 2^24 3711 -> 3637 us, 3 * 2^21 1457 -> 1427 us; slower below
// 2^18 (2^14: 23.3 -> 26.9 us), so the engine uses it from 2^19 points on (uzkge_cuda_configure "ntt_radix4": minimum log2 size, 0 = off).
template <int NT, bool SCATTER, bool R4 = false>
__global__ void __launch_bounds__(NT, (R4 && NT == 256) ? 4 : 1) ntt_pass_kernel(const NttKernelArgs a) {
    extern __shared__ uint4 ntt_smem[];
    const NttPass& p = a.p;
    const uint32_t logR = p.logR, logC = p.logC;
    const uint32_t R = 1u << logR, C = 1u << logC, T = R << logC;
    TileView tv{ntt_smem, ntt_smem + T, logC, C - 1};

    uint32_t tile = blockIdx.x;
    const uint32_t t = tile % p.inner_tiles;
    tile /= p.inner_tiles;
    const uint32_t o = tile % p.outer;
    const uint32_t b = tile / p.outer;
    const fe* const in = a.jobs.in[blockIdx.y];
    fe* const out = a.jobs.out[blockIdx.y];
    const uint64_t len_in = a.jobs.len_in[blockIdx.y];

    // ---- load (+ zero padding, + coset pre-scale on the very first read of the input)
    for (uint32_t idx = threadIdx.x; idx < T; idx += NT) {
        uint32_t r, c;
        if (p.in_r_contig) {
            r = idx & (R - 1);
            c = idx >> logR;
        } else {
            c = idx & (C - 1);
            r = idx >> logC;
        }
        const uint64_t g = ntt_in_index(p, b, o, t, r, c);
        fe x;
        if (a.zero_pad && g >= len_in) {
            x = fe_zero();
        } else {
            x = ld_fe(in + g);
            if (a.pre_coset) x = fe_mul<FrP>(x, ldg_fe(a.g_full + g));
        }
        tv.st(tv.pos(r, c), x);
    }
    __syncthreads();

    // ---- decimation-in-frequency stages: (u, v) -> (u + v, (u - v) * w_{2h}^j); result row = bitrev(k)
    int s_first = (int)logR - 1;
    if (R4) {
        // stages s (distance h) and s - 1 (distance q = h / 2) on the four rows r0, r0 + q, r0 + h, r0 + h + q, r0 = blk * 2h + j, j < q:
        //   stage s    : (x0, x2) with twiddle w_{2h}^j, (x1, x3) with w_{2h}^(j + q)
        //   stage s - 1: (a0, a1) and (a2, a3), both with w_h^j
        for (; s_first >= 1; s_first -= 2) {
            const int s = s_first;
            const uint32_t h = 1u << s, q = h >> 1;
            const uint32_t sh1 = logR - 1 - s, sh2 = logR - s;      // twiddle index shifts of the two stages
            const uint32_t log_nblk = logR - 1 - s;
            const bool by_block = logC >= 2 && logC + log_nblk >= 5;
            for (uint32_t bb = threadIdx.x; bb < (T >> 2); bb += NT) {
                const uint32_t c = bb & (C - 1), pr = bb >> logC;
                uint32_t j, blk;
                if (by_block) {
                    j = pr >> log_nblk;
                    blk = pr & ((1u << log_nblk) - 1);
                } else {
                    j = pr & (q - 1);
                    blk = pr >> (s - 1);
                }
                const uint32_t r0 = (blk << (s + 1)) | j;
                const uint32_t p0 = tv.pos(r0, c), p1 = tv.pos(r0 + q, c), p2 = tv.pos(r0 + h, c), p3 = tv.pos(r0 + h + q, c);
                fe a0, a1, a2, a3;
                {
                    const fe x0 = tv.ld(p0), x2 = tv.ld(p2);
                    a0 = fe_add<FrP>(x0, x2);
                    a2 = fe_sub<FrP>(x0, x2);
                    if (j) a2 = fe_mul<FrP>(a2, ldg_fe(a.stage_tab + (size_t)(j << sh1) * p.stage_stride));
                }
                {
                    const fe x1 = tv.ld(p1), x3 = tv.ld(p3);
                    a1 = fe_add<FrP>(x1, x3);
                    a3 = fe_mul<FrP>(fe_sub<FrP>(x1, x3), ldg_fe(a.stage_tab + (size_t)((j + q) << sh1) * p.stage_stride));
                }
                tv.st(p0, fe_add<FrP>(a0, a1));
                tv.st(p2, fe_add<FrP>(a2, a3));
                fe b1 = fe_sub<FrP>(a0, a1), b3 = fe_sub<FrP>(a2, a3);
                if (j) {
                    const fe w2 = ldg_fe(a.stage_tab + (size_t)(j << sh2) * p.stage_stride);
                    b1 = fe_mul<FrP>(b1, w2);
                    b3 = fe_mul<FrP>(b3, w2);
                }
                tv.st(p1, b1);
                tv.st(p3, b3);
            }
            __syncthreads();
        }
    }
    for (int s = s_first; s >= 0; s--) {
        const uint32_t h = 1u << s;
        const uint32_t tw_shift = logR - 1 - s;
        // butterfly -> thread mapping: where a warp would otherwise mix twiddle indices, walk the 2h-row blocks first so that
        // all 32 lanes share j -- the twiddle load is a broadcast, and the warps with j == 0 (w = 1: half of the butterflies
        // at h = 2, a quarter at h = 4, ...) skip the product altogether: ~N/2 of the (N/2)(log2 R - 1) products of a pass
        const uint32_t log_nblk = logR - 1 - s;
        const bool by_block = logC >= 2 && logC + log_nblk >= 5;
        for (uint32_t bb = threadIdx.x; bb < (T >> 1); bb += NT) {
            const uint32_t c = bb & (C - 1), pr = bb >> logC;
            uint32_t j, r;
            if (by_block) {
                j = pr >> log_nblk;
                r = ((pr & ((1u << log_nblk) - 1)) << (s + 1)) | j;
            } else {
                j = pr & (h - 1);
                r = ((pr >> s) << (s + 1)) | j;
            }
            const uint32_t p0 = tv.pos(r, c), p1 = tv.pos(r + h, c);
            const fe u = tv.ld(p0), v = tv.ld(p1);
            fe d = fe_sub<FrP>(u, v);
            if (j) d = fe_mul<FrP>(d, ldg_fe(a.stage_tab + (size_t)(j << tw_shift) * p.stage_stride));
            tv.st(p0, fe_add<FrP>(u, v));
            tv.st(p1, d);
        }
        __syncthreads();
    }

    // ---- store: natural row index k lives at row bitrev(k)
    for (uint32_t idx = threadIdx.x; idx < T; idx += NT) {
        const uint32_t c = idx & (C - 1), k = idx >> logC;
        fe x = tv.ld(tv.pos(ntt_bitrev(k, logR), c));
        uint64_t g = ntt_out_index(p, b, o, t, k, c);
        if (p.tw_mul) x = fe_mul<FrP>(x, ldg_fe(a.tw_pass + g));
        if (p.last) {
            if (a.inverse && g) g = a.n - g;
            if (a.post_coset)
                x = fe_mul<FrP>(x, ldg_fe(a.g_full + g));
            else if (a.has_scale)
                x = fe_mul<FrP>(x, a.scale);
        }
        if (SCATTER && p.last) {
            const uint64_t owner = g >> a.scatter_shift;
            const uint64_t pos = a.scatter_k1 + ((g & ((1ull << a.scatter_shift) - 1)) << a.scatter_log_g);
            st_fe(a.scatter_rows[owner] + pos, x);
        } else {
            st_fe(out + g, x);
        }
    }
}

// radix-3 pre-pass for N = 3 * M:  y[k*M + n] = w_N^(n*k) * sum_{m<3} x[m*M + n] * w_3^(m*k)
struct Radix3Args {
    NttJobs jobs;
    uint64_t m;
    const fe* w_lo;
    const fe* w_hi;
    const fe* g_full;
    const fe* tw3;   // tw3[n] = w^n, tw3[m + n] = w^(2n)
    fe w3, w3sq;
    uint32_t pre_coset;
};
__global__ void __launch_bounds__(256) ntt_radix3_kernel(const Radix3Args a) {
    const uint64_t n = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= a.m) return;
    const fe* const in = a.jobs.in[blockIdx.y];
    fe* const out = a.jobs.out[blockIdx.y];
    const uint64_t len_in = a.jobs.len_in[blockIdx.y];
    fe x[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const uint64_t g = (uint64_t)i * a.m + n;
        if (g < len_in) {
            x[i] = ld_fe(in + g);
            if (a.pre_coset) x[i] = fe_mul<FrP>(x[i], ldg_fe(a.g_full + g));
        } else {
            x[i] = fe_zero();
        }
    }
    const fe y0 = fe_add<FrP>(fe_add<FrP>(x[0], x[1]), x[2]);
    const fe b1 = fe_mul<FrP>(x[1], a.w3), c1 = fe_mul<FrP>(x[2], a.w3sq);
    const fe b2 = fe_mul<FrP>(x[1], a.w3sq), c2 = fe_mul<FrP>(x[2], a.w3);
    fe y1 = fe_add<FrP>(fe_add<FrP>(x[0], b1), c1);
    fe y2 = fe_add<FrP>(fe_add<FrP>(x[0], b2), c2);
    y1 = fe_mul<FrP>(y1, ldg_fe(a.tw3 + n));
    y2 = fe_mul<FrP>(y2, ldg_fe(a.tw3 + a.m + n));
    st_fe(out + n, y0);
    st_fe(out + a.m + n, y1);
    st_fe(out + 2 * a.m + n, y2);
}

// ---- cross-GPU step of a distributed transform of size N = G * L (G = 2, 4, 8 ranks, L = N / G per rank).
// After the first all-to-all a rank holds, for its `cols` columns n2 = col_offset + t, the G elements
// x[n1 * L + n2] (row n1 came from rank n1).  This kernel does the G-point transforms over n1 in registers and
// applies the four-step twiddle:   out[k1][t] = w_N^(n2 * k1) * sum_{n1} in[n1][t] * w_G^(n1 * k1)
// (inverse: conjugate roots and a factor 1/G).  Row k1 then travels to rank k1, which runs the size-L transform.
struct CrossArgs {
    const fe* in_rows[8];   // row n1: `cols` elements -- local memory after an all-to-all, or rank n1's slice itself (peer memory)
    fe* out_rows[8];        // row k1: local memory before an all-to-all, or rank k1's receive buffer itself (peer memory)
    uint64_t cols, col_offset, n_total;
    const fe* w_lo;   // power table of w_N
    const fe* w_hi;
    fe wg[4];         // w_G^j (conjugated for the inverse), j < G / 2
    fe scale;         // 1 / G
    uint32_t inverse;
};
template <int LOGG>
__global__ void __launch_bounds__(128) ntt_cross_kernel(const CrossArgs a) {
    constexpr int G = 1 << LOGG;
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.cols) return;
    fe x[G];
#pragma unroll
    for (int i = 0; i < G; i++) x[i] = ld_fe(a.in_rows[i] + t);
#pragma unroll
    for (int s = LOGG - 1; s >= 0; s--) {
        const int h = 1 << s;
#pragma unroll
        for (int pr = 0; pr < G / 2; pr++) {
            const int j = pr & (h - 1);
            const int r = ((pr >> s) << (s + 1)) | j;
            const fe u = x[r], v = x[r + h];
            fe d = fe_sub<FrP>(u, v);
            if (j) d = fe_mul<FrP>(d, a.wg[j << (LOGG - 1 - s)]);
            x[r] = fe_add<FrP>(u, v);
            x[r + h] = d;
        }
    }
    const uint64_t c = a.col_offset + t;
#pragma unroll
    for (int k = 0; k < G; k++) {
        int br = 0;
#pragma unroll
        for (int b = 0; b < LOGG; b++) br |= ((k >> b) & 1) << (LOGG - 1 - b);
        fe y = x[br];
        uint64_t e = (c * (uint64_t)k) % a.n_total;
        if (a.inverse) {
            if (e) e = a.n_total - e;
            y = fe_mul<FrP>(y, a.scale);
        }
        if (e) y = fe_mul<FrP>(y, pow2l(a.w_lo, a.w_hi, e));
        st_fe(a.out_rows[k] + t, y);
    }
}

// lo[i] = premul * base^i (i < 2^12);  hi[i] = base^(i << 12) (i < n_hi)
__global__ void ntt_build_pow_tables(fe base, fe premul, fe* lo, fe* hi, uint32_t n_hi) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nlo = 1u << NTT_LOG_TWLO;
    if (i < nlo) {
        st_fe(lo + i, fe_mul<FrP>(fe_pow_u64<FrP>(base, i), premul));
    } else if (i - nlo < n_hi) {
        st_fe(hi + (i - nlo), fe_pow_u64<FrP>(base, (uint64_t)(i - nlo) << NTT_LOG_TWLO));
    }
}
// full[j] = lo[j & 4095] * hi[j >> 12]  (= premul * base^j), j < n
__global__ void ntt_expand_pow_table(const fe* lo, const fe* hi, fe* full, uint64_t n) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j < n) st_fe(full + j, pow2l(lo, hi, j));
}
// tw3[n] = w^n, tw3[m + n] = w^(2n), n < m
__global__ void ntt_build_radix3_twiddles(const fe* lo, const fe* hi, fe* tw3, uint64_t m) {
    const uint64_t n = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= m) return;
    st_fe(tw3 + n, pow2l(lo, hi, n));
    st_fe(tw3 + m + n, pow2l(lo, hi, 2 * n));
}
// inter-pass twiddles of one pass, stored at the position the pass writes its result to
__global__ void __launch_bounds__(256) ntt_build_pass_twiddles(const NttPass p, const fe* lo, const fe* hi, fe* tw) {
    const uint32_t logR = p.logR, logC = p.logC, C = 1u << logC, T = (1u << logR) << logC;
    uint32_t tile = blockIdx.x;
    const uint32_t t = tile % p.inner_tiles;
    tile /= p.inner_tiles;
    const uint32_t o = tile % p.outer;
    const uint32_t b = tile / p.outer;
    for (uint32_t idx = threadIdx.x; idx < T; idx += blockDim.x) {
        const uint32_t c = idx & (C - 1), k = idx >> logC;
        const uint64_t e = ntt_tw_exponent(p, t, k, c);
        st_fe(tw + ntt_out_index(p, b, o, t, k, c), pow2l(lo, hi, e));
    }
}
// tab[j] = base^j, j < cnt
__global__ void ntt_build_stage_table(fe base, fe* tab, uint32_t cnt) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cnt) st_fe(tab + i, fe_pow_u64<FrP>(base, i));
}

// ------------------------------------------------------------------ host side
static fe host_pow(fe a, uint64_t e) { return fe_pow_u64<FrP>(a, e); }

const NttDomain* NttEngine::domain(uint64_t n, cudaStream_t st) {
    auto it = domains_.find(n);
    if (it != domains_.end()) return &it->second;
    NttDomain d;
    // small transforms: smaller tiles give more CTAs (measured: 2^16 takes 28 us with 2^9 tiles, 42 us with 2^10)
    uint32_t lg = 0;
    while ((1ull << (lg + 1)) <= n) lg++;
    uint32_t tile = lg > 16 ? lg - 7 : 9;
    if (tile > cfg_log_tile_) tile = cfg_log_tile_;
    const uint32_t max_r = cfg_max_log_r_ < tile ? cfg_max_log_r_ : tile;
    if (!ntt_make_plan(n, tile, max_r, cfg_two_pass_max_, &d.plan)) return nullptr;
    // small transforms: shrink the column count until there are enough CTAs to cover the SMs
    for (uint32_t i = 0; i < d.plan.npass; i++) {
        NttPass& p = d.plan.pass[i];
        while (p.logC > 2 && (uint64_t)p.inner_tiles * p.outer * p.batch < 2ull * sm_count_) {
            p.logC--;
            p.inner_tiles <<= 1;
        }
    }
    bool ok;
    d.omega = ntt_root_of_unity(n, &ok);
    if (!ok) return nullptr;
    fe nf = fe_zero();
    nf.l[0] = (uint32_t)n;
    nf.l[1] = (uint32_t)(n >> 32);
    d.n_inv = fe_inv<FrP>(fe_to_mont<FrP>(nf));
    d.w3 = host_pow(d.omega, d.plan.m);  // only meaningful when mixed
    d.w3sq = fe_sqr<FrP>(d.w3);
    const uint32_t nlo = 1u << NTT_LOG_TWLO;
    d.n_hi = (uint32_t)((n >> NTT_LOG_TWLO) + 1);
    const uint32_t log_rtab = d.plan.logm < 12 ? d.plan.logm : 12;
    d.n_stage = log_rtab ? (1u << (log_rtab - 1)) : 1;
    if (cudaMalloc(&d.w_lo, sizeof(fe) * ((size_t)nlo + d.n_hi + d.n_stage)) != cudaSuccess) return nullptr;
    d.w_hi = d.w_lo + nlo;
    d.stage = d.w_hi + d.n_hi;
    const uint32_t tot = nlo + d.n_hi;
    ntt_build_pow_tables<<<(tot + 127) / 128, 128, 0, st>>>(d.omega, fe_one<FrP>(), d.w_lo, d.w_hi, d.n_hi);
    const fe w_rtab = host_pow(d.omega, n >> log_rtab);
    ntt_build_stage_table<<<(d.n_stage + 127) / 128, 128, 0, st>>>(w_rtab, d.stage, d.n_stage);
    UZ_COUNT_LAUNCH(2);
    // inter-pass twiddles (and the radix-3 pre-pass twiddles), one entry per element: resident in HBM
    {
        size_t elems = 0;
        for (uint32_t i = 0; i < d.plan.npass; i++)
            if (d.plan.pass[i].tw_mul) elems += n;
        if (d.plan.mixed) elems += 2 * d.plan.m;
        if (elems) {
            if (cudaMalloc(&d.tw_all, sizeof(fe) * elems) != cudaSuccess) {
                cudaStreamSynchronize(st);   // the power tables are still being written
                cudaFree(d.w_lo);
                return nullptr;
            }
            fe* cur = d.tw_all;
            for (uint32_t i = 0; i < d.plan.npass; i++) {
                const NttPass& p = d.plan.pass[i];
                if (!p.tw_mul) continue;
                d.tw_pass[i] = cur;
                cur += n;
                ntt_build_pass_twiddles<<<p.inner_tiles * p.outer * p.batch, 256, 0, st>>>(p, d.w_lo, d.w_hi, d.tw_pass[i]);
                UZ_COUNT_LAUNCH(1);
            }
            if (d.plan.mixed) {
                d.tw3 = cur;
                ntt_build_radix3_twiddles<<<(unsigned)((d.plan.m + 255) / 256), 256, 0, st>>>(d.w_lo, d.w_hi, d.tw3, d.plan.m);
                UZ_COUNT_LAUNCH(1);
            }
        }
    }
    // tables are built once; later calls may come on other streams
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
        cudaFree(d.w_lo);
        if (d.tw_all) cudaFree(d.tw_all);
        return nullptr;
    }
    auto res = domains_.emplace(n, d);
    return &res.first->second;
}

const NttCoset* NttEngine::coset(uint64_t n, const fe& g, bool with_ninv, const NttDomain* d, cudaStream_t st) {
    for (auto& c : cosets_)
        if (c.n == n && c.with_ninv == with_ninv && fe_eq(c.g, g)) return &c;
    // FIFO cache: the prover uses one shift (k[1]) and its inverse per domain; the coset-by-coset quotient round six more.  At most
    // 16 tables and 8 GiB are kept (a table is n field elements)
    auto table_bytes = [](uint64_t m) { return sizeof(fe) * ((size_t)(1u << NTT_LOG_TWLO) + (m >> NTT_LOG_TWLO) + 1 + m); };
    size_t held = 0;
    for (auto& c : cosets_) held += table_bytes(c.n);
    while (!cosets_.empty() && (cosets_.size() >= 16 || held + table_bytes(n) > ((size_t)8 << 30))) {
        cudaDeviceSynchronize();   // users of the evicted table may be on any stream
        held -= table_bytes(cosets_.front().n);
        cudaFree(cosets_.front().g_lo);
        cosets_.erase(cosets_.begin());
    }
    NttCoset c;
    c.n = n;
    c.g = g;
    c.with_ninv = with_ninv;
    const uint32_t nlo = 1u << NTT_LOG_TWLO;
    c.n_hi = (uint32_t)((n >> NTT_LOG_TWLO) + 1);
    if (cudaMalloc(&c.g_lo, sizeof(fe) * ((size_t)nlo + c.n_hi + n)) != cudaSuccess) return nullptr;
    c.g_hi = c.g_lo + nlo;
    c.g_full = c.g_hi + c.n_hi;
    const uint32_t tot = nlo + c.n_hi;
    ntt_build_pow_tables<<<(tot + 127) / 128, 128, 0, st>>>(g, with_ninv ? d->n_inv : fe_one<FrP>(), c.g_lo, c.g_hi, c.n_hi);
    ntt_expand_pow_table<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(c.g_lo, c.g_hi, c.g_full, n);
    UZ_COUNT_LAUNCH(2);
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(st) != cudaSuccess) {
        cudaFree(c.g_lo);
        return nullptr;
    }
    cosets_.push_back(c);
    return &cosets_.back();
}

template <int NT, bool SCATTER = false, bool R4 = false>
static cudaError_t launch_pass(const NttKernelArgs& ka, uint32_t k, cudaStream_t st) {
    const NttPass& p = ka.p;
    const size_t T = (size_t)1 << (p.logR + p.logC);
    const size_t smem = T * 32;
    static size_t configured[UZ_MAX_DEVICES] = {};   // function attributes are per device (and per instantiation: the static is too)
    int dev = 0;
    cudaGetDevice(&dev);
    if (smem > 48 * 1024 && smem > configured[dev]) {
        cudaError_t e = cudaFuncSetAttribute(ntt_pass_kernel<NT, SCATTER, R4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(200 * 1024));
        if (e != cudaSuccess) return e;
        configured[dev] = 200 * 1024;
    }
    const uint32_t grid = p.inner_tiles * p.outer * p.batch;
    ntt_pass_kernel<NT, SCATTER, R4><<<dim3(grid, k), NT, smem, st>>>(ka);
    UZ_COUNT_LAUNCH(1);
    return cudaGetLastError();
}

// d_in: len_in elements readable; d_out: n elements; d_scratch: n elements (may alias neither).
// d_in == d_out is allowed.
int NttEngine::run(const fe* d_in, fe* d_out, fe* d_scratch, uint64_t len_in, uint64_t n, bool inverse,
                   const fe* coset_shift /* host, Montgomery, or null */, cudaStream_t st) {
    return run_batch(&d_in, &d_out, d_scratch, &len_in, 1, n, inverse, coset_shift, st);
}

// k <= NTT_MAX_BATCH independent transforms over the same domain (and coset), one launch per pass; d_scratch: k * n elements.
// The prover's rounds transform 5-8 polynomials of 2^13..2^14 coefficients at a time: one such vector is 16 tiles, i.e. a tenth of
// the GPU's SMs, and its passes are launch-bound.
int NttEngine::run_batch(const fe* const* d_in, fe* const* d_out, fe* d_scratch, const uint64_t* len_in, uint32_t k, uint64_t n, bool inverse,
                         const fe* coset_shift, cudaStream_t st, const NttScatter* scatter) {
    if (k == 0) return UZKGE_OK;
    if (k > NTT_MAX_BATCH) return UZKGE_ERR_SIZE;
    if (scatter && (k != 1 || scatter->log_g < 1 || scatter->log_g > 3 || (n & (n - 1)) || (n >> scatter->log_g) == 0 || scatter->k1 >= (1u << scatter->log_g)))
        return UZKGE_ERR_SIZE;
    const NttDomain* d = domain(n, st);
    if (!d) return UZKGE_ERR_SIZE;
    bool any_pad = false;
    for (uint32_t j = 0; j < k; j++) {
        if (len_in[j] > n) return UZKGE_ERR_SIZE;
        if (!d_in[j] || !d_out[j]) return UZKGE_ERR_ARG;
        any_pad = any_pad || len_in[j] < n;
    }
    const NttCoset* cs = nullptr;
    if (coset_shift) {
        cs = coset(n, *coset_shift, inverse, d, st);
        if (!cs) return UZKGE_ERR_CUDA;
    }
    const NttPlan& pl = d->plan;
    const int prof = g_prof.begin(Profiler::NTT, st);
    NttJobs cur;          // where each vector currently lives
    memset(&cur, 0, sizeof cur);
    for (uint32_t j = 0; j < k; j++) {
        cur.in[j] = d_in[j];
        cur.len_in[j] = len_in[j];
    }
    bool input_consumed = false;
    if (pl.mixed) {
        Radix3Args ra;
        ra.jobs = cur;
        for (uint32_t j = 0; j < k; j++) ra.jobs.out[j] = d_scratch + (uint64_t)j * n;
        ra.m = pl.m;
        ra.w_lo = d->w_lo;
        ra.w_hi = d->w_hi;
        ra.g_full = cs ? cs->g_full : nullptr;
        ra.tw3 = d->tw3;
        ra.w3 = d->w3;
        ra.w3sq = d->w3sq;
        ra.pre_coset = (cs && !inverse) ? 1 : 0;
        ntt_radix3_kernel<<<dim3((unsigned)((pl.m + 255) / 256), k), 256, 0, st>>>(ra);
        UZ_COUNT_LAUNCH(1);
        if (cudaGetLastError() != cudaSuccess) return UZKGE_ERR_CUDA;
        g_prof.mark(prof, NTT_PH_RADIX3, st);
        for (uint32_t j = 0; j < k; j++) cur.in[j] = ra.jobs.out[j];
        input_consumed = true;
    }
    for (uint32_t i = 0; i < pl.npass; i++) {
        NttKernelArgs ka;
        ka.p = pl.pass[i];
        ka.jobs = cur;
        // non-last passes write into scratch (in place once the data lives there); the last pass writes d_out
        for (uint32_t j = 0; j < k; j++) ka.jobs.out[j] = ka.p.last ? d_out[j] : d_scratch + (uint64_t)j * n;
        ka.n = n;
        ka.stage_tab = d->stage;
        ka.w_lo = d->w_lo;
        ka.w_hi = d->w_hi;
        ka.g_lo = cs ? cs->g_lo : nullptr;
        ka.g_hi = cs ? cs->g_hi : nullptr;
        ka.g_full = cs ? cs->g_full : nullptr;
        ka.tw_pass = d->tw_pass[i];
        ka.scale = d->n_inv;
        ka.inverse = inverse ? 1 : 0;
        ka.zero_pad = (!input_consumed && any_pad) ? 1 : 0;
        ka.pre_coset = (!input_consumed && cs && !inverse) ? 1 : 0;
        ka.post_coset = (ka.p.last && cs && inverse) ? 1 : 0;
        ka.has_scale = (ka.p.last && inverse && !cs) ? 1 : 0;
        const bool scat = scatter && ka.p.last;
        if (scat) {
            uint32_t lg = 0;
            while ((1ull << lg) < n) lg++;
            for (uint32_t r = 0; r < 8; r++) ka.scatter_rows[r] = r < (1u << scatter->log_g) ? scatter->rows[r] : nullptr;
            ka.scatter_log_g = scatter->log_g;
            ka.scatter_shift = lg - scatter->log_g;
            ka.scatter_k1 = scatter->k1;
        }
        // a single-tile last pass may run in place; a multi-tile last pass transposes and must not alias
        if (!scat && ka.p.last && (ka.p.inner_tiles * ka.p.outer * ka.p.batch) > 1)
            for (uint32_t j = 0; j < k; j++)
                if (ka.jobs.in[j] == ka.jobs.out[j]) return UZKGE_ERR_INTERNAL;
        const uint32_t T = 1u << (ka.p.logR + ka.p.logC);
        cudaError_t e;
        if (scat) {
            if (T >= 4096 && cfg_big_threads_ == 1024)
                e = launch_pass<1024, true>(ka, k, st);
            else if (T >= 2048)
                e = launch_pass<512, true>(ka, k, st);
            else if (T >= 512)
                e = launch_pass<256, true>(ka, k, st);
            else
                e = launch_pass<64, true>(ka, k, st);
        } else if (cfg_radix4_ && n >= (1ull << cfg_radix4_) && T >= 512 && T < 2048) {
            e = launch_pass<256, false, true>(ka, k, st);
        } else if (T >= 4096 && cfg_big_threads_ == 1024)
            e = launch_pass<1024>(ka, k, st);
        else if (T >= 2048)
            e = launch_pass<512>(ka, k, st);
        else if (T >= 512)
            e = launch_pass<256>(ka, k, st);
        else
            e = launch_pass<64>(ka, k, st);
        if (e != cudaSuccess) return UZKGE_ERR_CUDA;
        g_prof.mark(prof, NTT_PH_PASS0 + (int)i, st);
        for (uint32_t j = 0; j < k; j++) cur.in[j] = ka.jobs.out[j];
        input_consumed = true;
    }
    return UZKGE_OK;
}

int NttEngine::cross(const fe* d_in, fe* d_out, uint32_t log_g, uint64_t cols, uint64_t col_offset, uint64_t n_total,
                     bool inverse, cudaStream_t st) {
    if (log_g < 1 || log_g > 3) return UZKGE_ERR_SIZE;
    const fe* in_rows[8];
    fe* out_rows[8];
    for (uint32_t i = 0; i < (1u << log_g); i++) {
        in_rows[i] = d_in + (uint64_t)i * cols;
        out_rows[i] = d_out + (uint64_t)i * cols;
    }
    return cross_rows(in_rows, out_rows, log_g, cols, col_offset, n_total, inverse, st);
}

// the same step with one base pointer per row: rows may live in the memory of peer GPUs (cudaIpc mappings), which fuses
// the two exchanges of the four-step transform into the kernel's own loads and stores over NVLink
int NttEngine::cross_rows(const fe* const* in_rows, fe* const* out_rows, uint32_t log_g, uint64_t cols, uint64_t col_offset,
                          uint64_t n_total, bool inverse, cudaStream_t st) {
    if (log_g < 1 || log_g > 3 || cols == 0 || (n_total >> log_g) == 0 || col_offset + cols > (n_total >> log_g))
        return UZKGE_ERR_SIZE;
    if (n_total & (n_total - 1)) return UZKGE_ERR_SIZE;
    const NttDomain* d = domain(n_total, st);
    if (!d) return UZKGE_ERR_SIZE;
    CrossArgs ca;
    const uint32_t g = 1u << log_g;
    for (uint32_t i = 0; i < 8; i++) {
        ca.in_rows[i] = i < g ? in_rows[i] : nullptr;
        ca.out_rows[i] = i < g ? out_rows[i] : nullptr;
        if (i < g && (!ca.in_rows[i] || !ca.out_rows[i])) return UZKGE_ERR_ARG;
    }
    ca.cols = cols;
    ca.col_offset = col_offset;
    ca.n_total = n_total;
    ca.w_lo = d->w_lo;
    ca.w_hi = d->w_hi;
    ca.inverse = inverse ? 1 : 0;
    fe wg = host_pow(d->omega, n_total >> log_g);
    if (inverse) wg = fe_inv<FrP>(wg);
    fe acc = fe_one<FrP>();
    for (uint32_t j = 0; j < 4; j++) {
        ca.wg[j] = acc;
        acc = fe_mul<FrP>(acc, wg);
    }
    fe gf = fe_zero();
    gf.l[0] = g;
    ca.scale = fe_inv<FrP>(fe_to_mont<FrP>(gf));
    const unsigned grid = (unsigned)((cols + 127) / 128);
    if (log_g == 1)
        ntt_cross_kernel<1><<<grid, 128, 0, st>>>(ca);
    else if (log_g == 2)
        ntt_cross_kernel<2><<<grid, 128, 0, st>>>(ca);
    else
        ntt_cross_kernel<3><<<grid, 128, 0, st>>>(ca);
    UZ_COUNT_LAUNCH(1);
    return cudaGetLastError() == cudaSuccess ? UZKGE_OK : UZKGE_ERR_CUDA;
}

NttEngine::~NttEngine() {
    for (auto& kv : domains_) {
        cudaFree(kv.second.w_lo);
        if (kv.second.tw_all) cudaFree(kv.second.tw_all);
    }
    for (auto& c : cosets_) cudaFree(c.g_lo);
}

}  // namespace uz
