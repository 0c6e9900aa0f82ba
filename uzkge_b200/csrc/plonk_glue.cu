// plonk_glue.cu -- the elementwise pieces of a device-resident TurboPlonK prover: everything between the
// transforms, the MSMs and the scans that the reference does with per-coefficient loops on the CPU.
//
//   fr_lincomb        out = sum_j c_j * p_j      r_poly_or_comm's mul / add_assign chain (plonk/helpers.rs:716-745, 986-993)
//                                                and batch_prove's  h += (p - p(x)) * alpha^j  (poly_commit/pcs.rs:124-131)
//   fr_add_sparse     p[idx_j] += v_j            hide_polynomial (helpers.rs:139-154), split_t_and_commit's blinds
//                                                (helpers.rs:1351-1361), the "- eval" of batch_prove (pcs.rs:127)
//   fr_powers         out[i] = s * b^i           domain.elements() -> group / coset_quotient (plonk/indexer.rs:276-282)
//   fr_gather         out[i] = src[idx[i]]       ConstraintSystem::extend_witness (constraint_system/mod.rs:103-111),
//                                                encode_perm_to_group (indexer.rs:195-208)
//   plonk_perm_terms  numerator / denominator of z_poly's running product (helpers.rs:184-199)
//
// All are streaming kernels, one element per thread, 128-bit loads and stores; HBM-bound except perm_terms (18 products
// per 416 B: multiplier-bound like the quotient map).
#include <cuda_runtime.h>

#include "devmem.cuh"
#include "internal.h"

namespace uz {

struct LincombArgs {
    const fe* p[UZKGE_LINCOMB_MAX];
    uint64_t len[UZKGE_LINCOMB_MAX];
    fe c[UZKGE_LINCOMB_MAX];
    uint32_t k;
    uint64_t out_len;
    fe* out;
};

__global__ void __launch_bounds__(256) fr_lincomb_kernel(const __grid_constant__ LincombArgs a) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.out_len) return;
    fe acc = fe_zero();
    for (uint32_t j = 0; j < a.k; j++)
        if (i < a.len[j]) acc = fe_add<FrP>(acc, fe_mul<FrP>(ld_fe(a.p[j] + i), a.c[j]));
    st_fe(a.out + i, acc);
}

struct SparseArgs {
    uint64_t idx[UZKGE_SPARSE_MAX];
    fe v[UZKGE_SPARSE_MAX];
    uint32_t k;
};

// one thread: the indices may repeat (split_t_and_commit touches coefficient 0 and n of neighbouring pieces)
__global__ void fr_add_sparse_kernel(fe* __restrict__ p, const __grid_constant__ SparseArgs a) {
    if (threadIdx.x || blockIdx.x) return;
    for (uint32_t j = 0; j < a.k; j++) st_fe(p + a.idx[j], fe_add<FrP>(ld_fe(p + a.idx[j]), a.v[j]));
}

struct PowTable {
    fe p[40];  // p[s] = b^(2^s)
};

// 4 consecutive powers per thread: one table walk, then 3 products
__global__ void __launch_bounds__(256) fr_powers_kernel(const __grid_constant__ PowTable t, const fe scale, uint64_t n, fe* __restrict__ out) {
    const uint64_t i0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i0 >= n) return;
    fe r = scale;
    uint64_t e = i0;
    for (int s = 0; e; s++, e >>= 1)
        if (e & 1) r = fe_mul<FrP>(r, t.p[s]);
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (i0 + j < n) st_fe(out + i0 + j, r);
        r = fe_mul<FrP>(r, t.p[0]);
    }
}

__global__ void __launch_bounds__(256) fr_gather_kernel(const fe* __restrict__ src, const uint32_t* __restrict__ idx, uint64_t n, fe* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    st_fe(out + i, ld_fe(src + idx[i]));
}

// polys[j][idx[j]] += v[j] for k <= UZKGE_SPARSE_MULTI_MAX entries addressing DISTINCT elements (one thread each): the blinds of all the
// polynomials of a prover round in one launch
struct SparseMultiArgs {
    fe* p[UZKGE_SPARSE_MULTI_MAX];
    uint64_t idx[UZKGE_SPARSE_MULTI_MAX];
    fe v[UZKGE_SPARSE_MULTI_MAX];
    uint32_t k;
};
__global__ void __launch_bounds__(64) fr_add_sparse_multi_kernel(const __grid_constant__ SparseMultiArgs a) {
    const uint32_t j = threadIdx.x;
    if (j >= a.k) return;
    fe* q = a.p[j] + a.idx[j];
    st_fe(q, fe_add<FrP>(ld_fe(q), a.v[j]));
}

// dst[dst_idx[j]] = src[src_idx[j]]: the public-input rows of pi_poly's evaluation vector, straight from the witness (helpers.rs:111-131)
__global__ void __launch_bounds__(256) fr_gather_scatter_kernel(const fe* __restrict__ src, const uint32_t* __restrict__ src_idx,
                                                                fe* __restrict__ dst, const uint32_t* __restrict__ dst_idx, uint64_t k) {
    const uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    st_fe(dst + dst_idx[j], ld_fe(src + src_idx[j]));
}

__global__ void __launch_bounds__(256) fr_mul_kernel(const fe* __restrict__ a, const fe* __restrict__ b, uint64_t n, fe* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    st_fe(out + i, fe_mul<FrP>(ld_fe(a + i), ld_fe(b + i)));
}

// *out = max over non-zero elements of (index + 1); *out must be 0 on entry
__global__ void __launch_bounds__(256) fr_trimmed_len_kernel(const fe* __restrict__ p, uint64_t n, unsigned long long* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long v = 0;
    if (i < n && !fe_is_zero(ld_fe(p + i))) v = i + 1;
    for (int d = 16; d; d >>= 1) v = max(v, __shfl_xor_sync(0xffffffffu, v, d));
    if ((threadIdx.x & 31) == 0 && v) atomicMax(out, v);
}

struct PermArgs {
    const fe* w[5];      // extended witness, wire j at gate i
    const fe* sigma[5];  // encoded permutation k_{j'} * group[i'] of wire j at gate i
    const fe* group;     // omega^i
    fe beta_k[5];        // beta * k_j
    fe beta, gamma;
    uint64_t n;          // number of (num, den) pairs
    fe *num, *den;
};

__global__ void __launch_bounds__(128) plonk_perm_terms_kernel(const __grid_constant__ PermArgs a) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    const fe g = ld_fe(a.group + i);
    fe num = fe_one<FrP>(), den = fe_one<FrP>();
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const fe f = fe_add<FrP>(ld_fe(a.w[j] + i), a.gamma);
        num = fe_mul<FrP>(num, fe_add<FrP>(f, fe_mul<FrP>(a.beta_k[j], g)));
        den = fe_mul<FrP>(den, fe_add<FrP>(f, fe_mul<FrP>(a.beta, ld_fe(a.sigma[j] + i))));
    }
    st_fe(a.num + i, num);
    st_fe(a.den + i, den);
}

// ------------------------------------------------------------------ host side
static inline unsigned blocks_for(uint64_t n, unsigned per) { return (unsigned)((n + per - 1) / per); }

int fr_lincomb_run(const void* const* d_polys, const size_t* lens, const uint64_t* coefs, size_t k, void* d_out, size_t out_len, cudaStream_t st) {
    if (k == 0 || k > UZKGE_LINCOMB_MAX) return UZKGE_ERR_SIZE;
    if (!d_polys || !lens || !coefs || (!d_out && out_len)) return UZKGE_ERR_ARG;
    if (out_len == 0) return UZKGE_OK;
    LincombArgs a;
    memset(&a, 0, sizeof(a));
    for (size_t j = 0; j < k; j++) {
        if (!d_polys[j] && lens[j]) return UZKGE_ERR_ARG;
        a.p[j] = (const fe*)d_polys[j];
        a.len[j] = lens[j];
        memcpy(&a.c[j], coefs + 4 * j, sizeof(fe));
    }
    a.k = (uint32_t)k;
    a.out_len = out_len;
    a.out = (fe*)d_out;
    fr_lincomb_kernel<<<blocks_for(out_len, 256), 256, 0, st>>>(a);
    UZ_COUNT_LAUNCH(1);
    return cudaGetLastError() == cudaSuccess ? UZKGE_OK : UZKGE_ERR_CUDA;
}

int fr_add_sparse_run(void* d_poly, const size_t* idx, const uint64_t* vals, size_t k, cudaStream_t st) {
    if (k == 0) return UZKGE_OK;
    if (k > UZKGE_SPARSE_MAX) return UZKGE_ERR_SIZE;
    if (!d_poly || !idx || !vals) return UZKGE_ERR_ARG;
    SparseArgs a;
    memset(&a, 0, sizeof(a));
    for (size_t j = 0; j < k; j++) {
        a.idx[j] = idx[j];
        memcpy(&a.v[j], vals + 4 * j, sizeof(fe));
    }
    a.k = (uint32_t)k;
    fr_add_sparse_kernel<<<1, 32, 0, st>>>((fe*)d_poly, a);
    UZ_COUNT_LAUNCH(1);
    return cudaGetLastError() == cudaSuccess ? UZKGE_OK : UZKGE_ERR_CUDA;
}

int fr_add_sparse_multi_run(void* const* d_polys, const size_t* idx, const uint64_t* vals, size_t k, cudaStream_t st) {
    if (k == 0) return UZKGE_OK;
    if (k > UZKGE_SPARSE_MULTI_MAX) return UZKGE_ERR_SIZE;
    if (!d_polys || !idx || !vals) return UZKGE_ERR_ARG;
    SparseMultiArgs a;
    memset(&a, 0, sizeof(a));
    for (size_t j = 0; j < k; j++) {
        if (!d_polys[j]) return UZKGE_ERR_ARG;
        a.p[j] = (fe*)d_polys[j];
        a.idx[j] = idx[j];
        memcpy(&a.v[j], vals + 4 * j, sizeof(fe));
    }
    a.k = (uint32_t)k;
    fr_add_sparse_multi_kernel<<<1, 64, 0, st>>>(a);
    UZ_COUNT_LAUNCH(1);
    return cudaGetLastError() == cudaSuccess ? UZKGE_OK : UZKGE_ERR_CUDA;
}

int fr_powers_run(const uint64_t* base, const uint64_t* scale, size_t n, void* d_out, cudaStream_t st) {
    if (!base || (!d_out && n)) return UZKGE_ERR_ARG;
    if (n == 0) return UZKGE_OK;
    PowTable t;
    fe b;
    memcpy(&b, base, sizeof(fe));
    for (int s = 0; s < 40; s++) {
        t.p[s] = b;
        b = fe_sqr<FrP>(b);
    }
    fe sc = fe_one<FrP>();
    if (scale) memcpy(&sc, scale, sizeof(fe));
    fr_powers_kernel<<<blocks_for(n, 1024), 256, 0, st>>>(t, sc, n, (fe*)d_out);
    UZ_COUNT_LAUNCH(1);
    return cudaGetLastError() == cudaSuccess ? UZKGE_OK : UZKGE_ERR_CUDA;
}

int fr_gather_run(const void* d_src, const void* d_idx, size_t n, void* d_out, cudaStream_t st) {
    if (n == 0) return UZKGE_OK;
    if (!d_src || !d_idx || !d_out) return UZKGE_ERR_ARG;
    fr_gather_kernel<<<blocks_for(n, 256), 256, 0, st>>>((const fe*)d_src, (const uint32_t*)d_idx, n, (fe*)d_out);
    UZ_COUNT_LAUNCH(1);
    return cudaGetLastError() == cudaSuccess ? UZKGE_OK : UZKGE_ERR_CUDA;
}

int fr_gather_scatter_run(const void* d_src, const void* d_src_idx, void* d_dst, const void* d_dst_idx, size_t k, cudaStream_t st) {
    if (k == 0) return UZKGE_OK;
    if (!d_src || !d_src_idx || !d_dst || !d_dst_idx) return UZKGE_ERR_ARG;
    fr_gather_scatter_kernel<<<blocks_for(k, 256), 256, 0, st>>>((const fe*)d_src, (const uint32_t*)d_src_idx, (fe*)d_dst, (const uint32_t*)d_dst_idx, k);
    UZ_COUNT_LAUNCH(1);
    return cudaGetLastError() == cudaSuccess ? UZKGE_OK : UZKGE_ERR_CUDA;
}

// one coset of an interleaved domain <-> a compact vector (the quotient round split by cosets: prover.cu)
__global__ void __launch_bounds__(256) fr_strided_copy_kernel(const fe* src, size_t src_start, size_t src_step, fe* dst, size_t dst_start,
                                                              size_t dst_step, size_t count) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) st_fe(dst + dst_start + dst_step * i, ld_fe(src + src_start + src_step * i));
}

int fr_strided_copy_run(const void* d_src, size_t src_start, size_t src_step, void* d_dst, size_t dst_start, size_t dst_step, size_t count,
                        cudaStream_t st) {
    if (count == 0) return UZKGE_OK;
    if (!d_src || !d_dst) return UZKGE_ERR_ARG;
    fr_strided_copy_kernel<<<blocks_for(count, 256), 256, 0, st>>>((const fe*)d_src, src_start, src_step, (fe*)d_dst, dst_start, dst_step, count);
    UZ_COUNT_LAUNCH(1);
    return cudaGetLastError() == cudaSuccess ? UZKGE_OK : UZKGE_ERR_CUDA;
}

// ---- the quotient's coefficients from its per-coset inverse transforms (the quotient round split by cosets, prover.cu).
// t(X) = sum_r X^r T_r(X^n), T_r(Y) = sum_{c < f} t[r + n c] Y^c.  On the coset g_j <w_n> X^n is the constant K e_f^j (K = k1^n, e_f the
// primitive f-th root w_m^n), so the size-n coset iFFT of t's values on coset j is u_j[r] = T_r(K e_f^j), and per r an f-point
// inverse DFT recovers t[r + n c] = K^-c / f * sum_j u_j[r] e_f^(-j c).
struct CosetCombineArgs {
    const fe* u;      // f x n: u_j at [j n, (j + 1) n)
    fe* out;          // f n coefficients
    uint64_t n;
    uint32_t f;
    fe e_pow[16];     // e^k, e = e_f^-1
    fe scale[16];     // K^-c / f
};
// f = 6: e is a primitive 6th root of unity, e^2 = e - 1 and e^3 = -1, so e^k v is one of v, ev, ev - v and their negatives: ONE product
// per input, six for the scales -- 11 per r instead of the 36 of the general form below
__global__ void __launch_bounds__(128) plonk_coset_combine6_kernel(const __grid_constant__ CosetCombineArgs a) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    fe S[6];
    const fe v0 = ld_fe(a.u + r);
#pragma unroll
    for (int c = 0; c < 6; c++) S[c] = v0;
#pragma unroll
    for (int j = 1; j < 6; j++) {
        const fe v = ld_fe(a.u + (uint64_t)j * a.n + r);
        const fe b = fe_mul<FrP>(v, a.e_pow[1]);
        const fe d = fe_sub<FrP>(b, v);          // e^2 v
#pragma unroll
        for (int c = 0; c < 6; c++) {
            const int k = (j * c) % 6;           // compile-time after unrolling
            if (k == 0) S[c] = fe_add<FrP>(S[c], v);
            else if (k == 1) S[c] = fe_add<FrP>(S[c], b);
            else if (k == 2) S[c] = fe_add<FrP>(S[c], d);
            else if (k == 3) S[c] = fe_sub<FrP>(S[c], v);
            else if (k == 4) S[c] = fe_sub<FrP>(S[c], b);
            else S[c] = fe_sub<FrP>(S[c], d);
        }
    }
#pragma unroll
    for (int c = 0; c < 6; c++) st_fe(a.out + (uint64_t)c * a.n + r, fe_mul<FrP>(S[c], a.scale[c]));
}
__global__ void __launch_bounds__(128) plonk_coset_combine_kernel(const __grid_constant__ CosetCombineArgs a) {
    const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= a.n) return;
    for (uint32_t c = 0; c < a.f; c++) {
        fe acc = fe_zero();
        for (uint32_t j = 0; j < a.f; j++) {
            const fe v = ld_fe(a.u + (uint64_t)j * a.n + r);
            const uint32_t k = (j * c) % a.f;
            acc = fe_add<FrP>(acc, k ? fe_mul<FrP>(v, a.e_pow[k]) : v);
        }
        st_fe(a.out + (uint64_t)c * a.n + r, fe_mul<FrP>(acc, a.scale[c]));
    }
}

int plonk_coset_combine_run(const void* d_u, size_t n, size_t factor, const uint64_t* k1, void* d_out, cudaStream_t st) {
    if (!d_u || !d_out || !k1) return UZKGE_ERR_ARG;
    if (n == 0 || (n & (n - 1)) || factor == 0 || factor > 16) return UZKGE_ERR_SIZE;
    bool ok = false;
    const fe w_m = ntt_root_of_unity((uint64_t)n * factor, &ok);
    if (!ok) return UZKGE_ERR_SIZE;
    CosetCombineArgs a;
    memset(&a, 0, sizeof a);
    a.u = (const fe*)d_u;
    a.out = (fe*)d_out;
    a.n = n;
    a.f = (uint32_t)factor;
    const fe e = fe_inv<FrP>(fe_pow_u64<FrP>(w_m, n));        // e_f^-1
    fe kk;
    memcpy(&kk, k1, sizeof(fe));
    const fe K_inv = fe_inv<FrP>(fe_pow_u64<FrP>(kk, n));
    fe ff = fe_zero();
    ff.l[0] = (uint32_t)factor;
    fe sc = fe_inv<FrP>(fe_to_mont<FrP>(ff));
    fe ep = fe_one<FrP>();
    for (uint32_t c = 0; c < factor; c++) {
        a.e_pow[c] = ep;
        a.scale[c] = sc;
        ep = fe_mul<FrP>(ep, e);
        sc = fe_mul<FrP>(sc, K_inv);
    }
    if (factor == 6)
        plonk_coset_combine6_kernel<<<blocks_for(n, 128), 128, 0, st>>>(a);
    else
        plonk_coset_combine_kernel<<<blocks_for(n, 128), 128, 0, st>>>(a);
    UZ_COUNT_LAUNCH(1);
    return cudaGetLastError() == cudaSuccess ? UZKGE_OK : UZKGE_ERR_CUDA;
}

int fr_mul_run(const void* d_a, const void* d_b, size_t n, void* d_out, cudaStream_t st) {
    if (n == 0) return UZKGE_OK;
    if (!d_a || !d_b || !d_out) return UZKGE_ERR_ARG;
    fr_mul_kernel<<<blocks_for(n, 256), 256, 0, st>>>((const fe*)d_a, (const fe*)d_b, n, (fe*)d_out);
    UZ_COUNT_LAUNCH(1);
    return cudaGetLastError() == cudaSuccess ? UZKGE_OK : UZKGE_ERR_CUDA;
}

int fr_trimmed_len_run(const void* d_poly, size_t n, unsigned long long* d_scratch, size_t* len_out, cudaStream_t st) {
    if (!len_out || !d_scratch || (!d_poly && n)) return UZKGE_ERR_ARG;
    *len_out = 0;
    if (n == 0) return UZKGE_OK;
    if (cudaMemsetAsync(d_scratch, 0, sizeof(unsigned long long), st) != cudaSuccess) return UZKGE_ERR_CUDA;
    fr_trimmed_len_kernel<<<blocks_for(n, 256), 256, 0, st>>>((const fe*)d_poly, n, d_scratch);
    UZ_COUNT_LAUNCH(1);
    unsigned long long v = 0;
    if (cudaMemcpyAsync(&v, d_scratch, sizeof v, cudaMemcpyDeviceToHost, st) != cudaSuccess) return UZKGE_ERR_CUDA;
    if (cudaStreamSynchronize(st) != cudaSuccess) return UZKGE_ERR_CUDA;
    *len_out = (size_t)v;
    return UZKGE_OK;
}

// d_z: n elements (z[0] = 1, z[i + 1] = z[i] * num_i / den_i, i < n - 1);  d_tmp: 4 n elements of scratch
int plonk_z_evals_run(PolyEngine* poly, const void* const d_w[5], const void* const d_sigma[5], const void* d_group, const uint64_t* k,
                      const uint64_t* beta, const uint64_t* gamma, size_t n, void* d_z, void* d_tmp, cudaStream_t st) {
    if (!d_w || !d_sigma || !d_group || !k || !beta || !gamma || !d_z || !d_tmp) return UZKGE_ERR_ARG;
    if (n < 2) return UZKGE_ERR_SIZE;
    PermArgs a;
    memset(&a, 0, sizeof(a));
    memcpy(&a.beta, beta, sizeof(fe));
    memcpy(&a.gamma, gamma, sizeof(fe));
    for (int j = 0; j < 5; j++) {
        if (!d_w[j] || !d_sigma[j]) return UZKGE_ERR_ARG;
        a.w[j] = (const fe*)d_w[j];
        a.sigma[j] = (const fe*)d_sigma[j];
        fe kj;
        memcpy(&kj, k + 4 * j, sizeof(fe));
        a.beta_k[j] = fe_mul<FrP>(a.beta, kj);
    }
    a.group = (const fe*)d_group;
    a.n = n - 1;
    fe* tmp = (fe*)d_tmp;
    a.num = tmp;
    a.den = tmp + n;
    plonk_perm_terms_kernel<<<blocks_for(n - 1, 128), 128, 0, st>>>(a);
    UZ_COUNT_LAUNCH(1);
    if (cudaGetLastError() != cudaSuccess) return UZKGE_ERR_CUDA;
    return poly->grand_product(a.num, a.den, n - 1, (fe*)d_z, tmp + 2 * n, st);
}

}  // namespace uz
