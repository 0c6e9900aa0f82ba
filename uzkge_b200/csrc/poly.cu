// poly.cu -- the O(n) serial polynomial glue of the PlonK prover as scans and reductions over Fr.
//
//   Horner suffix scan   T_k = c_k + z * T_{k+1}          -> FpPolynomial::eval   (field_polynomial.rs:198-209): T_0
//                                                          -> div_rem by (X - z)  (field_polynomial.rs:519-550, as called from
//                                                             KZG `prove`, kzg_poly_commitment.rs:331-335): q_k = T_{k+1}, r = T_0
//   product scans        P_i = prod_{j<=i} a_j, S_i = prod_{j>=i} a_j
//                                                          -> z_poly's grand product (plonk/helpers.rs:204-217): batch inversion +
//                                                             running product, z_{i+1} = prod_{j<=i} num_j / den_j
//
// All three are memory-bound (32 B read + 32 B written per coefficient and pass).  A tile is 1024 elements (256 threads x 4);
// scans are three-phase: per-tile aggregates, one small scan over the aggregates, per-tile apply.  Powers z^(2^s) come from
// the host (ff.cuh's host twins).
#include <cuda_runtime.h>

#include "devmem.cuh"
#include "internal.h"

namespace uz {

static constexpr int PT = 256;              // threads per tile CTA
static constexpr int PE = 4;                // elements per thread
static constexpr uint32_t TILE = PT * PE;   // 1024
static constexpr uint32_t LOG_TILE = 10;

struct Pow2Table {
    fe p[40];  // p[s] = z^(2^s)
};

__device__ __forceinline__ fe z_pow(const Pow2Table& t, uint64_t e) {
    fe r = fe_one<FrP>();
    for (int s = 0; e; s++, e >>= 1)
        if (e & 1) r = fe_mul<FrP>(r, t.p[s]);
    return r;
}

// ------------------------------------------------------------------ Horner suffix scan
// tile b covers [b * TILE, min(n, (b + 1) * TILE)).  agg[b] = sum_j c_j z^(j - lo_b)
__global__ void __launch_bounds__(PT) horner_tile_agg_kernel(const fe* __restrict__ c, uint64_t n, const Pow2Table zt, fe* __restrict__ agg) {
    __shared__ fe sh[PT];
    const uint64_t lo = (uint64_t)blockIdx.x * TILE + (uint64_t)threadIdx.x * PE;
    fe a = fe_zero();
#pragma unroll
    for (int i = PE - 1; i >= 0; i--) {
        a = fe_mul<FrP>(a, zt.p[0]);
        if (lo + i < n) a = fe_add<FrP>(a, ld_fe(c + lo + i));
    }
    sh[threadIdx.x] = a;
    __syncthreads();
    // tree: the upper half is shifted by z^(PE * stride)
    for (uint32_t s = 0; (1u << s) < PT; s++) {
        const uint32_t stride = 1u << s;
        if ((threadIdx.x & (2 * stride - 1)) == 0)
            sh[threadIdx.x] = fe_add<FrP>(sh[threadIdx.x], fe_mul<FrP>(sh[threadIdx.x + stride], zt.p[s + 2]));  // z^(4 * 2^s)
        __syncthreads();
    }
    if (threadIdx.x == 0) st_fe(agg + blockIdx.x, sh[0]);
}

// carry[b] = T at index (b + 1) * TILE = sum_{b' > b} agg[b'] z^(TILE * (b' - b - 1));  total = T_0.  One CTA.
__global__ void __launch_bounds__(1024) horner_carry_kernel(const fe* __restrict__ agg, uint32_t nt, const Pow2Table zt, fe* __restrict__ carry,
                                                            fe* __restrict__ total) {
    __shared__ fe sh[1024];
    const uint32_t NTH = blockDim.x;          // a power of two <= 1024, sized to the tile count by the host (a 1024-thread scan
                                              // over 16 tiles spends 10 rounds of 32 warps on one SM: 44 us at n = 2^14)
    const uint32_t per = (nt + NTH - 1) / NTH;  // tiles per thread
    const uint32_t lo = threadIdx.x * per, hi = min(lo + per, nt);
    const fe zt_tile = zt.p[LOG_TILE];
    // thread aggregate over its tiles: A = sum_{b in [lo, hi)} agg[b] z^(TILE (b - lo))
    fe a = fe_zero();
    for (uint32_t b = hi; b > lo; b--) a = fe_add<FrP>(fe_mul<FrP>(a, zt_tile), ld_fe(agg + b - 1));
    sh[threadIdx.x] = a;
    __syncthreads();
    // Hillis-Steele suffix scan over threads: after step s, sh[t] covers threads t .. t + 2^(s+1) - 1
    const fe zper = z_pow(zt, (uint64_t)per << LOG_TILE);  // z^(TILE * per)
    fe zs = zper;
    fe mine = a;
    for (uint32_t stride = 1; stride < NTH; stride <<= 1) {
        fe other = fe_zero();
        const bool has = threadIdx.x + stride < NTH;
        if (has) other = sh[threadIdx.x + stride];
        __syncthreads();
        if (has) mine = fe_add<FrP>(mine, fe_mul<FrP>(other, zs));
        sh[threadIdx.x] = mine;
        __syncthreads();
        zs = fe_sqr<FrP>(zs);
    }
    // carry entering thread t's range from above = inclusive suffix of thread t + 1
    fe cin = threadIdx.x + 1 < NTH ? sh[threadIdx.x + 1] : fe_zero();
    // walk down the thread's tiles
    for (uint32_t b = hi; b > lo; b--) {
        st_fe(carry + b - 1, cin);
        cin = fe_add<FrP>(fe_mul<FrP>(cin, zt_tile), ld_fe(agg + b - 1));
    }
    if (threadIdx.x == 0) st_fe(total, sh[0]);
}

// out[k - 1] = T_k for 1 <= k < n  (the quotient by X - z)
__global__ void __launch_bounds__(PT) horner_apply_kernel(const fe* __restrict__ c, uint64_t n, const Pow2Table zt, const fe* __restrict__ carry,
                                                          fe* __restrict__ out) {
    __shared__ fe sh[PT];
    const uint64_t lo = (uint64_t)blockIdx.x * TILE + (uint64_t)threadIdx.x * PE;
    fe x[PE];
    fe a = fe_zero();
#pragma unroll
    for (int i = PE - 1; i >= 0; i--) {
        a = fe_mul<FrP>(a, zt.p[0]);
        if (lo + i < n) a = fe_add<FrP>(a, ld_fe(c + lo + i));
        x[i] = a;  // local suffix value with zero carry
    }
    sh[threadIdx.x] = a;
    __syncthreads();
    fe mine = a;
    fe zs = zt.p[2];  // z^PE
    for (uint32_t stride = 1; stride < PT; stride <<= 1) {
        fe other = fe_zero();
        const bool has = threadIdx.x + stride < PT;
        if (has) other = sh[threadIdx.x + stride];
        __syncthreads();
        if (has) mine = fe_add<FrP>(mine, fe_mul<FrP>(other, zs));
        sh[threadIdx.x] = mine;
        __syncthreads();
        zs = fe_sqr<FrP>(zs);
    }
    // carry into this thread's chunk: suffix of the next thread, plus the tile's carry shifted to the thread boundary
    fe cin = threadIdx.x + 1 < PT ? sh[threadIdx.x + 1] : fe_zero();
    const fe tile_carry = ld_fe(carry + blockIdx.x);
    const uint32_t above = (PT - 1 - threadIdx.x) * PE;  // elements between this chunk's end and the tile's end
    cin = fe_add<FrP>(cin, fe_mul<FrP>(tile_carry, z_pow(zt, above)));
    fe zk = zt.p[0];
#pragma unroll
    for (int i = PE - 1; i >= 0; i--) {
        // T_{lo+i} = x[i] + z^(PE - i) * cin
        const fe t = fe_add<FrP>(x[i], fe_mul<FrP>(cin, zk));
        zk = fe_mul<FrP>(zk, zt.p[0]);
        const uint64_t k = lo + i;
        if (k >= 1 && k < n) st_fe(out + k - 1, t);
    }
}

// ------------------------------------------------------------------ batched evaluation: k (polynomial, point) pairs in two launches
// (the prover's 15 openings at zeta / zeta * omega, plonk/prover.rs:217-244: at n = 2^14 one pair is two latency-bound launches)
struct EvalBatch {
    const fe* c[UZKGE_EVAL_BATCH_MAX];
    uint64_t n[UZKGE_EVAL_BATCH_MAX];
    uint32_t tile0[UZKGE_EVAL_BATCH_MAX + 1];  // first tile of pair j in the aggregate array
    uint32_t point[UZKGE_EVAL_BATCH_MAX];      // index into zt[]
    Pow2Table zt[2];                           // at most two distinct points per batch
    uint32_t k;
};

__global__ void __launch_bounds__(PT) horner_batch_agg_kernel(const __grid_constant__ EvalBatch b, fe* __restrict__ agg) {
    // blockIdx.x enumerates the tiles of all pairs: find the pair by a linear walk (k <= 32)
    uint32_t j = 0;
    while (j + 1 < b.k && blockIdx.x >= b.tile0[j + 1]) j++;
    const uint32_t tile = blockIdx.x - b.tile0[j];
    const Pow2Table& zt = b.zt[b.point[j]];
    const fe* c = b.c[j];
    const uint64_t n = b.n[j];
    __shared__ fe sh[PT];
    const uint64_t lo = (uint64_t)tile * TILE + (uint64_t)threadIdx.x * PE;
    fe a = fe_zero();
#pragma unroll
    for (int i = PE - 1; i >= 0; i--) {
        a = fe_mul<FrP>(a, zt.p[0]);
        if (lo + i < n) a = fe_add<FrP>(a, ld_fe(c + lo + i));
    }
    sh[threadIdx.x] = a;
    __syncthreads();
    for (uint32_t s = 0; (1u << s) < PT; s++) {
        const uint32_t stride = 1u << s;
        if ((threadIdx.x & (2 * stride - 1)) == 0)
            sh[threadIdx.x] = fe_add<FrP>(sh[threadIdx.x], fe_mul<FrP>(sh[threadIdx.x + stride], zt.p[s + 2]));
        __syncthreads();
    }
    if (threadIdx.x == 0) st_fe(agg + blockIdx.x, sh[0]);
}

// one CTA per pair: value = sum_t agg[t] z^(TILE t), a tree over the tile aggregates
__global__ void __launch_bounds__(256) horner_batch_total_kernel(const __grid_constant__ EvalBatch b, const fe* __restrict__ agg, fe* __restrict__ values) {
    const uint32_t j = blockIdx.x;
    const Pow2Table& zt = b.zt[b.point[j]];
    const uint32_t nt = b.tile0[j + 1] - b.tile0[j];
    const fe* a = agg + b.tile0[j];
    __shared__ fe sh[256];
    // thread t folds the tiles t, t + 256, ... (Horner in z^(256 TILE)), then a 256-wide tree in z^TILE
    const fe zbig = zt.p[LOG_TILE + 8];
    fe acc = fe_zero();
    if (threadIdx.x < nt) {
        uint32_t last = threadIdx.x + ((nt - 1 - threadIdx.x) / 256) * 256;
        for (int64_t t = last; t >= (int64_t)threadIdx.x; t -= 256) acc = fe_add<FrP>(fe_mul<FrP>(acc, zbig), ld_fe(a + t));
    }
    sh[threadIdx.x] = acc;
    __syncthreads();
    for (uint32_t s = 0; s < 8; s++) {
        const uint32_t stride = 1u << s;
        if ((threadIdx.x & (2 * stride - 1)) == 0)
            sh[threadIdx.x] = fe_add<FrP>(sh[threadIdx.x], fe_mul<FrP>(sh[threadIdx.x + stride], zt.p[LOG_TILE + s]));
        __syncthreads();
    }
    if (threadIdx.x == 0) st_fe(values + j, sh[0]);
}

// ------------------------------------------------------------------ product scans
// REVERSE = false: inclusive prefix products P_i;  true: inclusive suffix products S_i
template <bool REVERSE>
__device__ __forceinline__ uint64_t scan_index(uint64_t i, uint64_t n) {
    return REVERSE ? n - 1 - i : i;
}

template <bool REVERSE>
__global__ void __launch_bounds__(PT) prod_tile_agg_kernel(const fe* __restrict__ a, uint64_t n, fe* __restrict__ agg) {
    __shared__ fe sh[PT];
    const uint64_t lo = (uint64_t)blockIdx.x * TILE + (uint64_t)threadIdx.x * PE;
    fe p = fe_one<FrP>();
#pragma unroll
    for (int i = 0; i < PE; i++)
        if (lo + i < n) p = fe_mul<FrP>(p, ld_fe(a + scan_index<REVERSE>(lo + i, n)));
    sh[threadIdx.x] = p;
    __syncthreads();
    for (uint32_t stride = PT / 2; stride > 0; stride >>= 1) {
        if (threadIdx.x < stride) sh[threadIdx.x] = fe_mul<FrP>(sh[threadIdx.x], sh[threadIdx.x + stride]);
        __syncthreads();
    }
    if (threadIdx.x == 0) st_fe(agg + blockIdx.x, sh[0]);
}

// exclusive scan of the tile aggregates (one CTA): carry[b] = prod_{b' < b} agg[b'];  total = prod of all
__global__ void __launch_bounds__(1024) prod_carry_kernel(const fe* __restrict__ agg, uint32_t nt, fe* __restrict__ carry, fe* __restrict__ total) {
    __shared__ fe sh[1024];
    const uint32_t NTH = blockDim.x;   // a power of two <= 1024 (see horner_carry_kernel)
    const uint32_t per = (nt + NTH - 1) / NTH;
    const uint32_t lo = threadIdx.x * per, hi = min(lo + per, nt);
    fe p = fe_one<FrP>();
    for (uint32_t b = lo; b < hi; b++) p = fe_mul<FrP>(p, ld_fe(agg + b));
    sh[threadIdx.x] = p;
    __syncthreads();
    fe mine = p;
    for (uint32_t stride = 1; stride < NTH; stride <<= 1) {
        fe other = fe_one<FrP>();
        const bool has = threadIdx.x >= stride;
        if (has) other = sh[threadIdx.x - stride];
        __syncthreads();
        if (has) mine = fe_mul<FrP>(mine, other);
        sh[threadIdx.x] = mine;
        __syncthreads();
    }
    fe cin = threadIdx.x ? sh[threadIdx.x - 1] : fe_one<FrP>();
    for (uint32_t b = lo; b < hi; b++) {
        st_fe(carry + b, cin);
        cin = fe_mul<FrP>(cin, ld_fe(agg + b));
    }
    if (threadIdx.x == NTH - 1) st_fe(total, sh[NTH - 1]);
}

template <bool REVERSE>
__global__ void __launch_bounds__(PT) prod_apply_kernel(const fe* __restrict__ a, uint64_t n, const fe* __restrict__ carry, fe* __restrict__ out) {
    __shared__ fe sh[PT];
    const uint64_t lo = (uint64_t)blockIdx.x * TILE + (uint64_t)threadIdx.x * PE;
    fe x[PE];
    fe p = fe_one<FrP>();
#pragma unroll
    for (int i = 0; i < PE; i++) {
        if (lo + i < n) p = fe_mul<FrP>(p, ld_fe(a + scan_index<REVERSE>(lo + i, n)));
        x[i] = p;
    }
    sh[threadIdx.x] = p;
    __syncthreads();
    fe mine = p;
    for (uint32_t stride = 1; stride < PT; stride <<= 1) {
        fe other = fe_one<FrP>();
        const bool has = threadIdx.x >= stride;
        if (has) other = sh[threadIdx.x - stride];
        __syncthreads();
        if (has) mine = fe_mul<FrP>(mine, other);
        sh[threadIdx.x] = mine;
        __syncthreads();
    }
    fe cin = threadIdx.x ? sh[threadIdx.x - 1] : fe_one<FrP>();
    cin = fe_mul<FrP>(cin, ld_fe(carry + blockIdx.x));
#pragma unroll
    for (int i = 0; i < PE; i++)
        if (lo + i < n) st_fe(out + scan_index<REVERSE>(lo + i, n), fe_mul<FrP>(x[i], cin));
}

// out[0] = 1;  out[i + 1] = PN_i * SD_{i+1} / SD_0   (PN = prefix products of num, SD = suffix products of den, SD_n = 1)
__global__ void __launch_bounds__(256) grand_product_combine_kernel(const fe* __restrict__ pn, const fe* __restrict__ sd, fe inv_total, uint64_t n,
                                                                    fe* __restrict__ out) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    if (i == 0) {
        st_fe(out, fe_one<FrP>());
        return;
    }
    fe v = fe_mul<FrP>(ld_fe(pn + i - 1), inv_total);
    if (i < n) v = fe_mul<FrP>(v, ld_fe(sd + i));
    st_fe(out + i, v);
}

// ------------------------------------------------------------------ host side
// threads of the one-CTA carry scans: the tile count rounded up to a power of two, between one warp and 1024
static unsigned carry_threads(uint32_t nt) {
    unsigned t = 32;
    while (t < nt && t < 1024) t <<= 1;
    return t;
}

static Pow2Table make_pow2(const fe& z) {
    Pow2Table t;
    fe p = z;
    for (int s = 0; s < 40; s++) {
        t.p[s] = p;
        p = fe_sqr<FrP>(p);
    }
    return t;
}

static cudaError_t reserve(void** p, size_t* cap, size_t bytes) {
    if (bytes <= *cap) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc(p, bytes);
    if (e == cudaSuccess) *cap = bytes;
    return e;
}

PolyEngine::~PolyEngine() {
    if (ws_) cudaFree(ws_);
}

// d_c: n coefficients; d_quot: n - 1 coefficients or null (evaluation only); d_value: T_0
int PolyEngine::horner(const fe* d_c, uint64_t n, const fe& z, fe* d_quot, fe* d_value, cudaStream_t st) {
    if (n == 0 || n > (1ull << 32)) return UZKGE_ERR_SIZE;
    const uint32_t nt = (uint32_t)((n + TILE - 1) / TILE);
    if (reserve(&ws_, &ws_cap_, sizeof(fe) * (2ull * nt + 8)) != cudaSuccess) return UZKGE_ERR_OOM;
    fe* agg = (fe*)ws_;
    fe* carry = agg + nt;
    const Pow2Table zt = make_pow2(z);
    horner_tile_agg_kernel<<<nt, PT, 0, st>>>(d_c, n, zt, agg);
    horner_carry_kernel<<<1, carry_threads(nt), 0, st>>>(agg, nt, zt, carry, d_value);
    if (d_quot && n > 1) horner_apply_kernel<<<nt, PT, 0, st>>>(d_c, n, zt, carry, d_quot);
    UZ_COUNT_LAUNCH(d_quot && n > 1 ? 3 : 2);
    return cudaGetLastError() == cudaSuccess ? UZKGE_OK : UZKGE_ERR_CUDA;
}

// values[j] = p_j(x_{point[j]}): k <= UZKGE_EVAL_BATCH_MAX pairs, at most two distinct points
int PolyEngine::eval_batch(const fe* const* d_c, const uint64_t* n, const uint32_t* point, uint32_t k, const fe* points, uint32_t npoints,
                           fe* d_values, cudaStream_t st) {
    if (k == 0) return UZKGE_OK;
    if (k > UZKGE_EVAL_BATCH_MAX || npoints == 0 || npoints > 2) return UZKGE_ERR_SIZE;
    EvalBatch b;
    memset(&b, 0, sizeof(b));
    uint32_t tiles = 0;
    for (uint32_t j = 0; j < k; j++) {
        if (n[j] == 0 || n[j] > (1ull << 32) || point[j] >= npoints) return UZKGE_ERR_SIZE;
        if (!d_c[j]) return UZKGE_ERR_ARG;
        b.c[j] = d_c[j];
        b.n[j] = n[j];
        b.point[j] = point[j];
        b.tile0[j] = tiles;
        tiles += (uint32_t)((n[j] + TILE - 1) / TILE);
    }
    b.tile0[k] = tiles;
    b.k = k;
    for (uint32_t i = 0; i < npoints; i++) b.zt[i] = make_pow2(points[i]);
    if (reserve(&ws_, &ws_cap_, sizeof(fe) * ((size_t)tiles + 8)) != cudaSuccess) return UZKGE_ERR_OOM;
    fe* agg = (fe*)ws_;
    horner_batch_agg_kernel<<<tiles, PT, 0, st>>>(b, agg);
    horner_batch_total_kernel<<<k, 256, 0, st>>>(b, agg, d_values);
    UZ_COUNT_LAUNCH(2);
    return cudaGetLastError() == cudaSuccess ? UZKGE_OK : UZKGE_ERR_CUDA;
}

template <bool REVERSE>
static int product_scan(void** ws, size_t* cap, const fe* d_a, uint64_t n, fe* d_out, fe* d_total, cudaStream_t st) {
    const uint32_t nt = (uint32_t)((n + TILE - 1) / TILE);
    if (reserve(ws, cap, sizeof(fe) * (2ull * nt + 8)) != cudaSuccess) return UZKGE_ERR_OOM;
    fe* agg = (fe*)*ws;
    fe* carry = agg + nt;
    prod_tile_agg_kernel<REVERSE><<<nt, PT, 0, st>>>(d_a, n, agg);
    prod_carry_kernel<<<1, carry_threads(nt), 0, st>>>(agg, nt, carry, d_total);
    prod_apply_kernel<REVERSE><<<nt, PT, 0, st>>>(d_a, n, carry, d_out);
    UZ_COUNT_LAUNCH(3);
    return cudaGetLastError() == cudaSuccess ? UZKGE_OK : UZKGE_ERR_CUDA;
}

// d_out: n + 1 elements;  d_tmp: 2 n + 2 elements of scratch.  Returns UZKGE_ERR_ARG if a denominator is zero.
int PolyEngine::grand_product(const fe* d_num, const fe* d_den, uint64_t n, fe* d_out, fe* d_tmp, cudaStream_t st) {
    if (n == 0 || n > (1ull << 32)) return UZKGE_ERR_SIZE;
    fe* pn = d_tmp;
    fe* sd = d_tmp + n;
    fe* totals = d_tmp + 2 * n;  // [0] = prod num, [1] = prod den
    int rc = product_scan<false>(&ws_, &ws_cap_, d_num, n, pn, totals, st);
    if (rc != UZKGE_OK) return rc;
    rc = product_scan<true>(&ws_, &ws_cap_, d_den, n, sd, totals + 1, st);
    if (rc != UZKGE_OK) return rc;
    fe total_den;
    if (cudaMemcpyAsync(&total_den, totals + 1, sizeof(fe), cudaMemcpyDeviceToHost, st) != cudaSuccess) return UZKGE_ERR_CUDA;
    if (cudaStreamSynchronize(st) != cudaSuccess) return UZKGE_ERR_CUDA;
    if (fe_is_zero(total_den)) return UZKGE_ERR_ARG;
    const fe inv_total = fe_inv<FrP>(total_den);  // one inversion, on the host twin of the field code
    grand_product_combine_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(pn, sd, inv_total, n, d_out);
    UZ_COUNT_LAUNCH(1);
    return cudaGetLastError() == cudaSuccess ? UZKGE_OK : UZKGE_ERR_CUDA;
}

}  // namespace uz
