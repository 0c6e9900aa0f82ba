// msm.cu -- BN254 G1 variable-base MSM for sm_100a: fixed-base window tables + single-pass bucket method.
//
// Replaces `G1Projective::normalize_batch` + `G1Projective::msm(&points_raw, &coefs)` inside
// KZGCommitmentSchemeBN254::commit (/root/reference/uzkge/src/poly_commit/kzg_poly_commitment.rs:278-293).
//
// Design (KZG bases are fixed per SRS, HBM is 180 GB):
//   upload   : table f holds 2^(c*f) * P_i in affine form, f < W = ceil(255 / c)  (built on the device, once)
//   recode   : scalar -> canonical integer -> W signed radix-2^c digits d_f in [-2^(c-1), 2^(c-1)];
//              emits (key = |d_f|, val = sign | f*n + i) for every window
//   sort     : cub::DeviceRadixSort on the c-bit keys                      -> all windows share ONE bucket set
//   offsets  : bucket boundaries of the sorted key array
//   accumulate: G lanes per bucket walk the bucket's segment, mixed XYZZ additions (8M + 2S), points gathered
//              with cp.async through per-thread shared-memory slots (double buffered), then a warp-shuffle
//              reduction over the G lanes.  Buckets far above the mean (skewed witness scalars: 0/1/small
//              values) go to a list handled by whole CTAs (slices of LARGE_SLICE entries).
//   reduce   : sum_b b * B_b (msm_reduce.cu): recursive row / column marginal sums of the bucket matrix, leaves
//              finished by bit decomposition, additions split over teams of 4 lanes where parallelism is scarce.
//              No doubling ladder over windows: the tables removed the window combine.
#include <cuda_runtime.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "devmem.cuh"
#include "ec_compact.cuh"
#include "internal.h"

namespace uz {

static constexpr int ACC_NT = 256;            // threads per CTA of the accumulate kernel
static constexpr uint32_t LARGE_SLICE = 1024; // entries per warp slice of an oversized bucket
static constexpr int LARGE_NT = 256;

// ------------------------------------------------------------------ recode
struct RecodeArgs {
    const fe* scalars;
    uint32_t n;             // scalars in this MSM
    uint32_t c, windows;
    uint32_t table_stride;  // SRS length (distance between tables, in points)
    uint32_t base_offset;
    uint32_t* keys;
    uint32_t* vals;
};

__global__ void __launch_bounds__(256) msm_recode_kernel(const RecodeArgs a) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    fe s = fe_from_mont<FrP>(ld_fe(a.scalars + i));
    const uint32_t c = a.c, mask = (1u << c) - 1, half = 1u << (c - 1);
    uint32_t carry = 0;
    const uint32_t point = a.base_offset + i;
    for (uint32_t f = 0; f < a.windows; f++) {
        uint32_t d = (s.l[0] & mask) + carry;
#pragma unroll
        for (int k = 0; k < 7; k++) s.l[k] = __funnelshift_r(s.l[k], s.l[k + 1], c);
        s.l[7] >>= c;
        uint32_t neg = 0;
        carry = 0;
        if (d > half) {  // d in (2^(c-1), 2^c]  ->  d - 2^c in (-2^(c-1), 0]
            d = (1u << c) - d;
            neg = d ? 0x80000000u : 0u;
            carry = 1;
        }
        const size_t o = (size_t)f * a.n + i;
        a.keys[o] = d;
        a.vals[o] = (f * a.table_stride + point) | neg;
    }
}

// ---- counting sort by bucket (alternative to the radix sort; zero digits are dropped, order inside a bucket is
// arbitrary).  Pass 1 counts, a scan turns counts into offsets, pass 2 recomputes the digits and scatters.
template <bool SCATTER>
__global__ void __launch_bounds__(256) msm_count_scatter_kernel(const RecodeArgs a, uint32_t* __restrict__ counts,
                                                                const uint32_t* __restrict__ offsets, uint32_t* __restrict__ sorted_vals) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n) return;
    fe s = fe_from_mont<FrP>(ld_fe(a.scalars + i));
    const uint32_t c = a.c, mask = (1u << c) - 1, half = 1u << (c - 1);
    uint32_t carry = 0;
    const uint32_t point = a.base_offset + i;
    for (uint32_t f = 0; f < a.windows; f++) {
        uint32_t d = (s.l[0] & mask) + carry;
#pragma unroll
        for (int k = 0; k < 7; k++) s.l[k] = __funnelshift_r(s.l[k], s.l[k + 1], c);
        s.l[7] >>= c;
        uint32_t neg = 0;
        carry = 0;
        if (d > half) {
            d = (1u << c) - d;
            neg = 0x80000000u;
            carry = 1;
        }
        if (d == 0) continue;
        if (SCATTER) {
            const uint32_t pos = offsets[d] + atomicAdd(counts + d, 1u);
            sorted_vals[pos] = (f * a.table_stride + point) | neg;
        } else {
            atomicAdd(counts + d, 1u);
        }
    }
}

// offsets[b] = first sorted position whose key is >= b, for b in [0, nbuckets]; offsets[nbuckets] = m
__global__ void __launch_bounds__(256) msm_offsets_kernel(const uint32_t* __restrict__ keys, uint32_t m, uint32_t nbuckets,
                                                          uint32_t* __restrict__ offsets) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > m) return;
    const uint32_t cur = (i == m) ? nbuckets : min(keys[i], nbuckets);
    const uint32_t lo = (i == 0) ? 0 : min(keys[i - 1], nbuckets) + 1;
    for (uint32_t b = lo; b <= cur; b++) offsets[b] = i;
}

// key = cap - min(size, cap): an ascending sort visits the largest segments first
__global__ void __launch_bounds__(256) msm_sizes_kernel(const uint32_t* __restrict__ offsets, uint32_t nbuckets, uint32_t cap,
                                                        uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nbuckets) return;
    const uint32_t size = b ? offsets[b + 1] - offsets[b] : 0;
    keys[b] = cap - min(size, cap);
    vals[b] = b;
}

// ------------------------------------------------------------------ accumulate
struct AccArgs {
    const affine* tables;
    const uint32_t* vals;     // sorted
    const uint32_t* offsets;  // nbuckets + 1
    const uint32_t* order;    // bucket ids, largest segment first
    xyzz* buckets;            // nb_padded
    uint32_t nbuckets;        // valid bucket ids are 1 .. nbuckets-1
    uint32_t nb_padded;       // rows * cols of the reduction matrix
    uint32_t large_threshold;
    uint32_t* large_list;     // [0] = count, [1 + k] = bucket id
    uint32_t large_cap;
};

__device__ __forceinline__ affine lds_point(const uint4* slot, uint32_t stride) {
    affine p;
    p.x = fe_from_u4(slot[0], slot[stride]);
    p.y = fe_from_u4(slot[2 * stride], slot[3 * stride]);
    return p;
}
__device__ __forceinline__ void cp_async_point(uint4* slot, uint32_t stride, const affine* src) {
    const uint4* g = reinterpret_cast<const uint4*>(src);
    cp_async16(slot, g);
    cp_async16(slot + stride, g + 1);
    cp_async16(slot + 2 * stride, g + 2);
    cp_async16(slot + 3 * stride, g + 3);
}

// Walk sorted entries first, first + step, ... < end and add the referenced table points into acc.
// smem: per-thread slots, stage s chunk k of thread t lives at sm[(s * 4 + k) * NT + t].
template <int NT>
__device__ __forceinline__ void accumulate_segment(xyzz& acc, const affine* __restrict__ tables, const uint32_t* __restrict__ vals,
                                                   uint32_t first, uint32_t end, uint32_t step, uint4* sm) {
    uint4* slot0 = sm + threadIdx.x;
    uint4* slot1 = sm + 4 * NT + threadIdx.x;
    uint32_t j = first;
    uint32_t v = 0;
    if (j < end) {
        v = vals[j];
        cp_async_point(slot0, NT, tables + (v & 0x7fffffffu));
    }
    cp_async_commit();
    uint32_t stage = 0;
#pragma unroll 1
    while (j < end) {
        const uint32_t jn = j + step;
        uint32_t vn = 0;
        if (jn < end) {
            vn = vals[jn];
            cp_async_point(stage ? slot0 : slot1, NT, tables + (vn & 0x7fffffffu));
        }
        cp_async_commit();
        cp_async_wait<1>();  // everything but the newest group has landed: the current stage is readable
        affine p = lds_point(stage ? slot1 : slot0, NT);
        if (v >> 31) p.y = fe_neg<FqP>(p.y);
        xyzz_madd(acc, p);
        j = jn;
        v = vn;
        stage ^= 1;
    }
    cp_async_wait<0>();
}

template <int G>
__global__ void __launch_bounds__(ACC_NT, 2) msm_accumulate_kernel(const AccArgs a) {
    extern __shared__ uint4 acc_smem[];
    const uint32_t gtid = blockIdx.x * ACC_NT + threadIdx.x;
    const uint32_t lane = gtid & (G - 1);
    const uint32_t slot = gtid / G;
    // buckets are visited in decreasing size: the lanes of a warp get equal work and the longest segments start first
    const uint32_t b = slot < a.nb_padded ? a.order[slot] : a.nb_padded;
    uint32_t start = 0, end = 0;
    bool write = b < a.nb_padded;
    if (b >= 1 && b < a.nbuckets) {
        start = a.offsets[b];
        end = a.offsets[b + 1];
        if (end - start > a.large_threshold) {
            if (lane == 0) {
                const uint32_t k = atomicAdd(a.large_list, 1u);
                if (k < a.large_cap) a.large_list[1 + k] = b;
            }
            write = false;
            end = start;
        }
    }
    xyzz acc = xyzz_identity();
    accumulate_segment<ACC_NT>(acc, a.tables, a.vals, start + lane, end, G, acc_smem);
#pragma unroll 1
    for (int off = G >> 1; off > 0; off >>= 1) acc = xyzz_add_call(acc, shfl_xor_xyzz(acc, off));
    if (write && lane == 0) st_xyzz(a.buckets + b, acc);
}

// ---- oversized buckets: plan (slices per bucket, prefix sum), accumulate per slice, finish per bucket
struct LargeArgs {
    const affine* tables;
    const uint32_t* vals;
    const uint32_t* offsets;
    xyzz* buckets;
    uint32_t* large_list;   // [0] = count, [1 + k] = bucket
    uint32_t* slice_start;  // large_cap + 1: exclusive prefix of slices per listed bucket
    xyzz* slice_sums;       // one per slice
    uint32_t large_cap;
    uint32_t max_slices;
};

__global__ void __launch_bounds__(1024) msm_large_plan_kernel(const LargeArgs a) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry_s;
    const uint32_t nl = min(a.large_list[0], a.large_cap);
    if (threadIdx.x == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nl; base += 1024) {
        const uint32_t k = base + threadIdx.x;
        uint32_t cnt = 0;
        if (k < nl) {
            const uint32_t b = a.large_list[1 + k];
            cnt = (a.offsets[b + 1] - a.offsets[b] + LARGE_SLICE - 1) / LARGE_SLICE;
        }
        uint32_t x = cnt;  // inclusive warp scan
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, off);
            if ((threadIdx.x & 31) >= (uint32_t)off) x += y;
        }
        if ((threadIdx.x & 31) == 31) warp_tot[threadIdx.x >> 5] = x;
        __syncthreads();
        uint32_t wbase = 0;
        for (uint32_t w = 0; w < (threadIdx.x >> 5); w++) wbase += warp_tot[w];
        const uint32_t carry = carry_s;
        if (k < nl) a.slice_start[k] = carry + wbase + x - cnt;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + wbase + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) a.slice_start[nl] = carry_s;
}

// one WARP per slice of LARGE_SLICE entries: lanes stride over the slice, then a shuffle tree
__global__ void __launch_bounds__(LARGE_NT, 2) msm_large_accumulate_kernel(const LargeArgs a) {
    extern __shared__ uint4 acc_smem[];
    const uint32_t nl = min(a.large_list[0], a.large_cap);
    const uint32_t total = min(a.slice_start[nl], a.max_slices);
    const uint32_t s = blockIdx.x * (LARGE_NT / 32) + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31;
    if (s >= total) return;  // warp-uniform
    // largest k with slice_start[k] <= s
    uint32_t lo = 0, hi = nl;
    while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (a.slice_start[mid] <= s) lo = mid; else hi = mid;
    }
    const uint32_t b = a.large_list[1 + lo];
    const uint32_t first = a.offsets[b] + (s - a.slice_start[lo]) * LARGE_SLICE;
    const uint32_t end = min(first + LARGE_SLICE, a.offsets[b + 1]);
    xyzz acc = xyzz_identity();
    accumulate_segment<LARGE_NT>(acc, a.tables, a.vals, first + lane, end, 32, acc_smem);
#pragma unroll 1
    for (int off = 16; off > 0; off >>= 1) acc = xyzz_add_call(acc, shfl_down_xyzz(acc, off));
    if (lane == 0) st_xyzz(a.slice_sums + s, acc);
}

// one warp per listed bucket: sum its slice sums
__global__ void __launch_bounds__(32) msm_large_finish_kernel(const LargeArgs a) {
    const uint32_t nl = min(a.large_list[0], a.large_cap);
    const uint32_t k = blockIdx.x;
    if (k >= nl) return;
    const uint32_t s0 = a.slice_start[k], s1 = min(a.slice_start[k + 1], a.max_slices);
    xyzz acc = xyzz_identity();
#pragma unroll 1
    for (uint32_t s = s0 + threadIdx.x; s < s1; s += 32) acc = xyzz_add_call(acc, ld_xyzz(a.slice_sums + s));
#pragma unroll 1
    for (int off = 16; off > 0; off >>= 1) acc = xyzz_add_call(acc, shfl_down_xyzz(acc, off));
    if (threadIdx.x == 0) st_xyzz(a.buckets + a.large_list[1 + k], acc);
}

__global__ void msm_identity_kernel(jacobian* out) {
    const jacobian j = xyzz_to_jacobian(xyzz_identity());
    st_fe(&out->x, j.x);
    st_fe(&out->y, j.y);
    st_fe(&out->z, j.z);
}

// ------------------------------------------------------------------ small group helpers
__device__ __forceinline__ xyzz jacobian_to_xyzz(const jacobian& p) {
    xyzz r;
    if (fe_is_zero(p.z)) return xyzz_identity();
    r.x = p.x;
    r.y = p.y;
    r.zz = fe_sqr<FqP>(p.z);
    r.zzz = fe_mul<FqP>(r.zz, p.z);
    return r;
}
__device__ __forceinline__ jacobian ld_jacobian(const jacobian* p) {
    jacobian r;
    r.x = ld_fe(&p->x);
    r.y = ld_fe(&p->y);
    r.z = ld_fe(&p->z);
    return r;
}
__global__ void g1_add_kernel(const jacobian* a, const jacobian* b, jacobian* out) {
    xyzz x = jacobian_to_xyzz(ld_jacobian(a));
    const xyzz y = jacobian_to_xyzz(ld_jacobian(b));
    xyzz_add(x, y);
    const jacobian j = xyzz_to_jacobian(x);
    st_fe(&out->x, j.x);
    st_fe(&out->y, j.y);
    st_fe(&out->z, j.z);
}
__global__ void g1_to_affine_kernel(const jacobian* in, affine* out) {
    const affine r = xyzz_to_affine(jacobian_to_xyzz(ld_jacobian(in)));
    st_affine(out, r);
}

// ------------------------------------------------------------------ SRS generation (setup path)
// out[i] = tau^i * G, i < n, affine: the G1 half of KZGCommitmentScheme::new
// (/root/reference/uzkge/src/poly_commit/kzg_poly_commitment.rs:183-204, n sequential scalar multiplications
// there).  One thread per point: tau^i by square-and-multiply in Fr, then double-and-add over the bits of the
// canonical scalar, then one Fermat inversion.
__global__ void __launch_bounds__(128) g1_powers_of_tau_kernel(fe tau, affine g, uint64_t first, uint32_t count, affine* out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const fe s = fe_from_mont<FrP>(fe_pow_u64<FrP>(tau, first + t));
    xyzz acc = xyzz_identity();
#pragma unroll 1
    for (int bit = 253; bit >= 0; bit--) {
        acc = xyzz_dbl(acc);
        if ((s.l[bit >> 5] >> (bit & 31)) & 1) xyzz_madd(acc, g);
    }
    st_affine(out + t, xyzz_to_affine(acc));
}

// ------------------------------------------------------------------ fixed-base tables
// next[i] = 2^c * prev[i] for i in the slab [first, first + count).  Thread t handles the points
// first + t + j * nthreads, j < B: doubles each in XYZZ (kept in `tmp`), then normalises the B points with
// one shared inversion (Montgomery's trick; `pre` holds the running products).
struct TableArgs {
    const affine* prev;
    affine* next;
    xyzz* tmp;   // count entries
    fe* pre;     // count entries
    uint32_t count, c, per_thread, nthreads;
};
__global__ void __launch_bounds__(128) msm_table_kernel(const TableArgs a) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= a.nthreads) return;
    fe prod = fe_one<FqP>();
#pragma unroll 1
    for (uint32_t j = 0; j < a.per_thread; j++) {
        const uint32_t i = t + j * a.nthreads;
        if (i >= a.count) break;
        const affine p = ld_affine(a.prev + i);
        xyzz q;
        if (affine_is_identity(p)) {
            q = xyzz_identity();
        } else {
            q = xyzz_dbl_affine(p);
            for (uint32_t k = 1; k < a.c; k++) q = xyzz_dbl(q);
        }
        st_xyzz(a.tmp + i, q);
        if (!xyzz_is_identity(q)) prod = fe_mul<FqP>(prod, fe_mul<FqP>(q.zz, q.zzz));
        st_fe(a.pre + i, prod);
    }
    fe inv = fe_inv<FqP>(prod);
#pragma unroll 1
    for (uint32_t j = a.per_thread; j-- > 0;) {
        const uint32_t i = t + j * a.nthreads;
        if (i >= a.count) continue;
        const xyzz q = ld_xyzz(a.tmp + i);
        affine r;
        if (xyzz_is_identity(q)) {
            r.x = fe_zero();
            r.y = fe_zero();
        } else {
            const fe before = j ? ld_fe(a.pre + (i - a.nthreads)) : fe_one<FqP>();
            const fe ti = fe_mul<FqP>(inv, before);            // 1 / (zz * zzz)
            inv = fe_mul<FqP>(inv, fe_mul<FqP>(q.zz, q.zzz));
            r.x = fe_mul<FqP>(q.x, fe_mul<FqP>(ti, q.zzz));         // X / ZZ
            r.y = fe_mul<FqP>(q.y, fe_mul<FqP>(ti, q.zz));          // Y / ZZZ
        }
        st_affine(a.next + i, r);
    }
}

// ------------------------------------------------------------------ host side

// Window size for an SRS of n points, from B200 measurements (scripts/gpu_tune.py csweep).  Only sizes whose TOP
// window is wide enough are used: the top window holds t = 255 - c (W - 1) bits, its digits land in 2^(t-1) buckets,
// and when that is far fewer than the other windows' 2^(c-1) those buckets receive 2^(c-t) / W times the mean load
// (c = 18: 2 bits, c = 19: 7 bits, c = 21: 3 bits ... measured 20-30 % slower than their neighbours).
static uint32_t choose_window_bits(size_t n) {
    uint32_t lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) lg++;
    if (lg <= 9) return 9;     // t = 3 of 9: balanced
    if (lg <= 13) return 10;   // 2^12: 296 us
    if (lg <= 15) return 13;   // 2^14: 350 us
    if (lg <= 16) return 15;   // 2^16: 510 us
    if (lg <= 18) return 17;   // 2^18: 1181 us
    return 20;                 // 2^20: 2.90 ms, 2^22: 10.4 ms
}

#define UZ_CUDA_TRY(expr)                                   \
    do {                                                    \
        cudaError_t e__ = (expr);                           \
        if (e__ != cudaSuccess) return cuda_err_code(e__);  \
    } while (0)

int MsmEngine::upload(const uint64_t* affine_xy_host, size_t n, uint32_t window_bits, MsmSrs* s, cudaStream_t st) {
    if (n == 0 || n >= (1ull << 28)) return UZKGE_ERR_SIZE;
    uint32_t c = window_bits ? window_bits : choose_window_bits(n);
    if (c < 2 || c > 24) return UZKGE_ERR_SIZE;
    const uint32_t windows = (255 + c - 1) / c;
    if ((uint64_t)windows * n >= (1ull << 31)) return UZKGE_ERR_SIZE;
    *s = MsmSrs();
    s->n = n;
    s->c = c;
    s->windows = windows;
    s->nbuckets = (1u << (c - 1)) + 1;
    s->logcols = c / 2;                       // cols = 2^ceil((c-1)/2)
    s->cols = 1u << s->logcols;
    s->rows = 1u << (c - 1 - s->logcols);     // rows * cols = 2^(c-1); the top bucket 2^(c-1) is stored after the matrix
    s->nb_padded = s->nbuckets;
    const size_t m = (size_t)windows * n;
    // segments longer than max(8 * mean, 64 * lanes) are "large": at most nbuckets / 8 of them can exist
    s->large_cap = s->nbuckets / 8 + 2;
    s->max_slices = (uint32_t)(m / LARGE_SLICE + s->large_cap);

    size_t total = 0;
    auto take = [&](size_t bytes) {
        const size_t off = total;
        total += (bytes + 255) & ~(size_t)255;
        return off;
    };
    const size_t o_tables = take(sizeof(affine) * m);
    const size_t o_keys_a = take(4 * m), o_keys_b = take(4 * m), o_vals_a = take(4 * m), o_vals_b = take(4 * m);
    const size_t o_offsets = take(4 * ((size_t)s->nbuckets + 1));
    const size_t o_ord = take(4 * 4 * (size_t)s->nbuckets);
    const size_t o_counts = take(4 * ((size_t)s->nbuckets + 1));
    const size_t o_large = take(4 * ((size_t)s->large_cap + 1));
    const size_t o_slice_start = take(4 * ((size_t)s->large_cap + 1));
    const size_t o_slice_sums = take(sizeof(xyzz) * s->max_slices);
    const size_t o_buckets = take(sizeof(xyzz) * s->nb_padded);
    const size_t reduce_bytes = msm_reduce_workspace_bytes(c);
    if (reduce_bytes == 0) return UZKGE_ERR_SIZE;
    const size_t o_reduce = take(reduce_bytes);
    const size_t o_ticket = take(256);
    cub::DoubleBuffer<uint32_t> dk(nullptr, nullptr), dv(nullptr, nullptr);
    s->cub_temp_bytes = 0;
    UZ_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, s->cub_temp_bytes, dk, dv, (int)m, 0, (int)c, st));
    size_t ord_temp = 0;
    UZ_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, ord_temp, dk, dv, (int)s->nbuckets, 0, 32, st));
    if (ord_temp > s->cub_temp_bytes) s->cub_temp_bytes = ord_temp;
    size_t scan_temp = 0;
    UZ_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, scan_temp, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)s->nbuckets + 1, st));
    if (scan_temp > s->cub_temp_bytes) s->cub_temp_bytes = scan_temp;
    const size_t o_cub = take(s->cub_temp_bytes + 256);
    UZ_CUDA_TRY(cudaMalloc(&s->arena, total));
    s->bytes = total;
    char* base = (char*)s->arena;
    s->tables = (affine*)(base + o_tables);
    s->keys_a = (uint32_t*)(base + o_keys_a);
    s->keys_b = (uint32_t*)(base + o_keys_b);
    s->vals_a = (uint32_t*)(base + o_vals_a);
    s->vals_b = (uint32_t*)(base + o_vals_b);
    s->offsets = (uint32_t*)(base + o_offsets);
    s->counts = (uint32_t*)(base + o_counts);
    s->ord_keys_a = (uint32_t*)(base + o_ord);
    s->ord_keys_b = s->ord_keys_a + s->nbuckets;
    s->ord_vals_a = s->ord_keys_b + s->nbuckets;
    s->ord_vals_b = s->ord_vals_a + s->nbuckets;
    s->large_list = (uint32_t*)(base + o_large);
    s->slice_start = (uint32_t*)(base + o_slice_start);
    s->slice_sums = (xyzz*)(base + o_slice_sums);
    s->buckets = (xyzz*)(base + o_buckets);
    s->ticket = (uint32_t*)(base + o_ticket);
    s->reduce = msm_reduce_plan_create(c, s->buckets, base + o_reduce, s->ticket);
    if (!s->reduce) return UZKGE_ERR_INTERNAL;
    s->cub_temp = base + o_cub;

    cudaEvent_t e0, e1;
    UZ_CUDA_TRY(cudaEventCreate(&e0));
    UZ_CUDA_TRY(cudaEventCreate(&e1));
    UZ_CUDA_TRY(cudaMemsetAsync(s->ticket, 0, 256, st));
    UZ_CUDA_TRY(cudaMemcpyAsync(s->tables, affine_xy_host, sizeof(affine) * n, cudaMemcpyHostToDevice, st));
    UZ_CUDA_TRY(cudaEventRecord(e0, st));
    // slabs of at most 2^20 points bound the XYZZ / prefix-product scratch
    const size_t slab = n < (1u << 20) ? n : (1u << 20);
    xyzz* tmp = nullptr;
    UZ_CUDA_TRY(cudaMalloc(&tmp, slab * (sizeof(xyzz) + sizeof(fe))));
    fe* pre = (fe*)(tmp + slab);
    for (uint32_t f = 1; f < windows; f++) {
        for (size_t first = 0; first < n; first += slab) {
            TableArgs ta;
            ta.count = (uint32_t)((n - first) < slab ? (n - first) : slab);
            ta.prev = s->tables + (size_t)(f - 1) * n + first;
            ta.next = s->tables + (size_t)f * n + first;
            ta.tmp = tmp;
            ta.pre = pre;
            ta.c = c;
            uint32_t per = (uint32_t)(ta.count / ((size_t)sm_count_ * 512));
            per = per < 1 ? 1 : (per > 16 ? 16 : per);
            ta.per_thread = per;
            ta.nthreads = (ta.count + per - 1) / per;
            msm_table_kernel<<<(ta.nthreads + 127) / 128, 128, 0, st>>>(ta);
            UZ_COUNT_LAUNCH(1);
        }
    }
    UZ_CUDA_TRY(cudaGetLastError());
    UZ_CUDA_TRY(cudaEventRecord(e1, st));
    UZ_CUDA_TRY(cudaStreamSynchronize(st));
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    s->precompute_ms = ms;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(tmp);
    return UZKGE_OK;
}

void MsmEngine::release(MsmSrs* s) {
    if (s->reduce) msm_reduce_plan_destroy(s->reduce);
    if (s->arena) cudaFree(s->arena);
    *s = MsmSrs();
}

template <int G>
static cudaError_t launch_accumulate(const AccArgs& a, cudaStream_t st) {
    const uint64_t threads = (uint64_t)a.nb_padded * G;
    const uint32_t grid = (uint32_t)((threads + ACC_NT - 1) / ACC_NT);
    msm_accumulate_kernel<G><<<grid, ACC_NT, 2 * 4 * ACC_NT * sizeof(uint4), st>>>(a);
    return cudaGetLastError();
}

int MsmEngine::run(MsmSrs* s, size_t base_offset, const fe* d_scalars, size_t n, jacobian* d_out, cudaStream_t st) {
    if (base_offset > s->n || n > s->n - base_offset) return UZKGE_ERR_SIZE;
    if (n == 0) {
        msm_identity_kernel<<<1, 1, 0, st>>>(d_out);
        UZ_COUNT_LAUNCH(1);
        UZ_CUDA_TRY(cudaGetLastError());
        return UZKGE_OK;
    }
    const uint32_t m = (uint32_t)(s->windows * n);
    const int prof = g_prof.begin(Profiler::MSM, st);
    RecodeArgs ra;
    ra.scalars = d_scalars;
    ra.n = (uint32_t)n;
    ra.c = s->c;
    ra.windows = s->windows;
    ra.table_stride = (uint32_t)s->n;
    ra.base_offset = (uint32_t)base_offset;
    ra.keys = s->keys_a;
    ra.vals = s->vals_a;
    const uint32_t* vals;
    size_t temp;
    if (counting_sort_) {
        // count -> scan -> scatter: 2 x N*W global atomics instead of 3 radix passes over N*W pairs
        const size_t cbytes = 4 * ((size_t)s->nbuckets + 1);
        UZ_CUDA_TRY(cudaMemsetAsync(s->counts, 0, cbytes, st));
        msm_count_scatter_kernel<false><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ra, s->counts, nullptr, nullptr);
        UZ_CUDA_TRY(cudaGetLastError());
        g_prof.mark(prof, MSM_PH_RECODE, st);
        temp = s->cub_temp_bytes;
        UZ_CUDA_TRY(cub::DeviceScan::ExclusiveSum(s->cub_temp, temp, s->counts, s->offsets, (int)s->nbuckets + 1, st));
        UZ_CUDA_TRY(cudaMemsetAsync(s->counts, 0, cbytes, st));
        msm_count_scatter_kernel<true><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ra, s->counts, s->offsets, s->vals_b);
        UZ_CUDA_TRY(cudaGetLastError());
        vals = s->vals_b;
        g_prof.mark(prof, MSM_PH_SORT, st);
    } else {
        msm_recode_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(ra);
        UZ_CUDA_TRY(cudaGetLastError());
        g_prof.mark(prof, MSM_PH_RECODE, st);

        cub::DoubleBuffer<uint32_t> dk(s->keys_a, s->keys_b), dv(s->vals_a, s->vals_b);
        temp = s->cub_temp_bytes;
        UZ_CUDA_TRY(cub::DeviceRadixSort::SortPairs(s->cub_temp, temp, dk, dv, (int)m, 0, (int)s->c, st));
        const uint32_t* keys = dk.Current();
        vals = dv.Current();
        g_prof.mark(prof, MSM_PH_SORT, st);

        msm_offsets_kernel<<<(m + 1 + 255) / 256, 256, 0, st>>>(keys, m, s->nbuckets, s->offsets);
        UZ_CUDA_TRY(cudaGetLastError());
    }
    UZ_CUDA_TRY(cudaMemsetAsync(s->large_list, 0, 4, st));

    // lanes per bucket: ~48 entries per lane, but at least enough groups to fill every SM
    const double mean = (double)m / (double)(s->nbuckets - 1);
    uint32_t g = 1;
    while (g < 32 && mean / (g * 2) >= 40.0) g *= 2;
    while (g < 32 && (uint64_t)s->nbuckets * g * 2 <= (uint64_t)sm_count_ * 512) g *= 2;
    if (force_lanes_) g = force_lanes_;
    // segments above the threshold are split into CTA slices (msm_large_*)
    uint32_t thr = (uint32_t)(mean * 8.0);
    {
        uint32_t per_lane = (uint32_t)(m / (2ull * sm_count_ * 512));
        if (per_lane < 64) per_lane = 64;
        if (thr < per_lane * g) thr = per_lane * g;
    }

    uint32_t cap_bits = 1;
    while ((1u << cap_bits) <= thr + 1) cap_bits++;
    msm_sizes_kernel<<<(s->nbuckets + 255) / 256, 256, 0, st>>>(s->offsets, s->nbuckets, thr + 1, s->ord_keys_a, s->ord_vals_a);
    UZ_CUDA_TRY(cudaGetLastError());
    cub::DoubleBuffer<uint32_t> ok(s->ord_keys_a, s->ord_keys_b), ov(s->ord_vals_a, s->ord_vals_b);
    temp = s->cub_temp_bytes;
    UZ_CUDA_TRY(cub::DeviceRadixSort::SortPairs(s->cub_temp, temp, ok, ov, (int)s->nbuckets, 0, (int)cap_bits, st));
    g_prof.mark(prof, MSM_PH_OFFSETS, st);

    AccArgs aa;
    aa.tables = s->tables;
    aa.vals = vals;
    aa.offsets = s->offsets;
    aa.order = ov.Current();
    aa.buckets = s->buckets;
    aa.nbuckets = s->nbuckets;
    aa.nb_padded = s->nb_padded;
    aa.large_threshold = thr;
    aa.large_list = s->large_list;
    aa.large_cap = s->large_cap;
    cudaError_t e;
    switch (g) {
        case 1: e = launch_accumulate<1>(aa, st); break;
        case 2: e = launch_accumulate<2>(aa, st); break;
        case 4: e = launch_accumulate<4>(aa, st); break;
        case 8: e = launch_accumulate<8>(aa, st); break;
        case 16: e = launch_accumulate<16>(aa, st); break;
        default: e = launch_accumulate<32>(aa, st); break;
    }
    UZ_CUDA_TRY(e);
    g_prof.mark(prof, MSM_PH_ACCUMULATE, st);

    LargeArgs la;
    la.tables = s->tables;
    la.vals = vals;
    la.offsets = s->offsets;
    la.buckets = s->buckets;
    la.large_list = s->large_list;
    la.slice_start = s->slice_start;
    la.slice_sums = s->slice_sums;
    la.large_cap = s->large_cap;
    la.max_slices = s->max_slices;
    // upper bounds for this m (the kernels read the real counts on the device; surplus CTAs exit at once)
    uint32_t cap_now = (uint32_t)(m / aa.large_threshold + 1);
    if (cap_now > s->large_cap) cap_now = s->large_cap;
    const uint32_t slices_now = m / LARGE_SLICE + cap_now;
    msm_large_plan_kernel<<<1, 1024, 0, st>>>(la);
    msm_large_accumulate_kernel<<<(slices_now + LARGE_NT / 32 - 1) / (LARGE_NT / 32), LARGE_NT, 2 * 4 * LARGE_NT * sizeof(uint4), st>>>(la);
    msm_large_finish_kernel<<<cap_now, 32, 0, st>>>(la);
    UZ_CUDA_TRY(cudaGetLastError());
    g_prof.mark(prof, MSM_PH_LARGE, st);

    {
        const int rc = msm_reduce_run(s->reduce, d_out, st);
        if (rc != UZKGE_OK) return rc;
    }
    g_prof.mark(prof, MSM_PH_REDUCE, st);
    UZ_COUNT_LAUNCH(7 + 3 + 2);  // own kernels (the reduction counts its own) + CUB's sort / scan launches
    return UZKGE_OK;
}

int MsmEngine::powers_of_tau(const fe& tau, uint64_t first, uint32_t count, affine* d_out, cudaStream_t st) {
    affine g;  // the generator (1, 2)
    g.x = fe_one<FqP>();
    g.y = fe_dbl<FqP>(g.x);
    g1_powers_of_tau_kernel<<<(count + 127) / 128, 128, 0, st>>>(tau, g, first, count, d_out);
    UZ_COUNT_LAUNCH(1);
    UZ_CUDA_TRY(cudaGetLastError());
    return UZKGE_OK;
}

int MsmEngine::g1_add(const jacobian* d_a, const jacobian* d_b, jacobian* d_out, cudaStream_t st) {
    g1_add_kernel<<<1, 1, 0, st>>>(d_a, d_b, d_out);
    UZ_COUNT_LAUNCH(1);
    UZ_CUDA_TRY(cudaGetLastError());
    return UZKGE_OK;
}
int MsmEngine::g1_to_affine(const jacobian* d_in, affine* d_out, cudaStream_t st) {
    g1_to_affine_kernel<<<1, 1, 0, st>>>(d_in, d_out);
    UZ_COUNT_LAUNCH(1);
    UZ_CUDA_TRY(cudaGetLastError());
    return UZKGE_OK;
}

}  // namespace uz
