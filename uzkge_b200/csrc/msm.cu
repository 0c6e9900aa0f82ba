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
static constexpr uint32_t LARGE_SLICE_MIN = 128, LARGE_SLICE_MAX = 1024;  // entries per warp slice of an oversized bucket (chosen per run)
static constexpr int LARGE_NT = 256;

// ------------------------------------------------------------------ digits + counting sort by bucket
// A launch handles up to MSM_MAX_BATCH independent MSMs over the same SRS (blockIdx.y = MSM j); MSM j owns the global
// bucket ids j * nbuckets + d.  Pass 1 counts the entries per bucket, a scan turns counts into offsets, pass 2
// recomputes the digits and scatters  val = sign | (f * table_stride + point)  into its bucket's segment.  Zero
// digits are dropped; the order inside a bucket is arbitrary (the sum does not depend on it).
static constexpr uint32_t MSM_MAX_BATCH = 16;
struct DigitArgs {
    const fe* scalars[MSM_MAX_BATCH];
    uint32_t n[MSM_MAX_BATCH];
    uint32_t c, windows;
    uint32_t table_stride;  // SRS length (distance between tables, in points)
    uint32_t base_offset;
    uint32_t nbuckets;
};

template <bool SCATTER>
__global__ void __launch_bounds__(256) msm_count_scatter_kernel(const DigitArgs a, uint32_t* __restrict__ counts,
                                                                const uint32_t* __restrict__ offsets, uint32_t* __restrict__ sorted_vals) {
    const uint32_t j = blockIdx.y;
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= a.n[j]) return;
    fe s = fe_from_mont<FrP>(ld_fe(a.scalars[j] + i));
    const uint32_t c = a.c, mask = (1u << c) - 1, half = 1u << (c - 1);
    const uint32_t bucket0 = j * a.nbuckets;
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
            neg = 0x80000000u;
            carry = 1;
        }
        if (d == 0) continue;
        if (SCATTER) {
            const uint32_t pos = offsets[bucket0 + d] + atomicAdd(counts + bucket0 + d, 1u);
            sorted_vals[pos] = (f * a.table_stride + point) | neg;
        } else {
            atomicAdd(counts + bucket0 + d, 1u);
        }
    }
}

// key = cap - min(size, cap): an ascending sort visits the largest segments first
__global__ void __launch_bounds__(256) msm_sizes_kernel(const uint32_t* __restrict__ offsets, uint32_t total_buckets, uint32_t cap,
                                                        uint32_t* __restrict__ keys, uint32_t* __restrict__ vals) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= total_buckets) return;
    const uint32_t size = offsets[b + 1] - offsets[b];  // the zero-digit bucket of every MSM is empty by construction
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
    uint32_t nb_padded;       // batch * nbuckets global bucket ids
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
    if (b < a.nb_padded) {
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
    uint32_t slice;         // entries per slice: a lane adds slice / 32 points in a dependent chain
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
            cnt = (a.offsets[b + 1] - a.offsets[b] + a.slice - 1) / a.slice;
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

// one WARP per slice of a.slice entries: lanes stride over the slice, then a shuffle tree.  The grid is a fixed
// number of CTAs; warps walk the slice list (its length is only known on the device).
__global__ void __launch_bounds__(LARGE_NT, 2) msm_large_accumulate_kernel(const LargeArgs a) {
    extern __shared__ uint4 acc_smem[];
    const uint32_t nl = min(a.large_list[0], a.large_cap);
    const uint32_t total = min(a.slice_start[nl], a.max_slices);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t nwarps = gridDim.x * (LARGE_NT / 32);
#pragma unroll 1
    for (uint32_t s = blockIdx.x * (LARGE_NT / 32) + (threadIdx.x >> 5); s < total; s += nwarps) {
        // largest k with slice_start[k] <= s
        uint32_t lo = 0, hi = nl;
        while (hi - lo > 1) {
            const uint32_t mid = (lo + hi) >> 1;
            if (a.slice_start[mid] <= s) lo = mid; else hi = mid;
        }
        const uint32_t b = a.large_list[1 + lo];
        const uint32_t first = a.offsets[b] + (s - a.slice_start[lo]) * a.slice;
        const uint32_t end = min(first + a.slice, a.offsets[b + 1]);
        xyzz acc = xyzz_identity();
        accumulate_segment<LARGE_NT>(acc, a.tables, a.vals, first + lane, end, 32, acc_smem);
#pragma unroll 1
        for (int off = 16; off > 0; off >>= 1) acc = xyzz_add_call(acc, shfl_down_xyzz(acc, off));
        if (lane == 0) st_xyzz(a.slice_sums + s, acc);
    }
}

// one warp per listed bucket: sum its slice sums
__global__ void __launch_bounds__(128) msm_large_finish_kernel(const LargeArgs a) {
    const uint32_t nl = min(a.large_list[0], a.large_cap);
    const uint32_t lane = threadIdx.x & 31;
    const uint32_t nwarps = gridDim.x * 4;
#pragma unroll 1
    for (uint32_t k = blockIdx.x * 4 + (threadIdx.x >> 5); k < nl; k += nwarps) {
        const uint32_t s0 = a.slice_start[k], s1 = min(a.slice_start[k + 1], a.max_slices);
        xyzz acc = xyzz_identity();
#pragma unroll 1
        for (uint32_t s = s0 + lane; s < s1; s += 32) acc = xyzz_add_call(acc, ld_xyzz(a.slice_sums + s));
#pragma unroll 1
        for (int off = 16; off > 0; off >>= 1) acc = xyzz_add_call(acc, shfl_down_xyzz(acc, off));
        if (lane == 0) st_xyzz(a.buckets + a.large_list[1 + k], acc);
    }
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
// out[j] = sum_{r < count} parts[r * stride + j], j < k: the N - 1 projective additions that merge per-GPU partial sums (k results of
// a batch, gathered rank-major), without leaving the device
__global__ void g1_sum_kernel(const jacobian* parts, uint32_t count, uint32_t stride, uint32_t k, jacobian* out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= k) return;
    xyzz acc = jacobian_to_xyzz(ld_jacobian(parts + j));
    for (uint32_t i = 1; i < count; i++) {
        const xyzz y = jacobian_to_xyzz(ld_jacobian(parts + (size_t)i * stride + j));
        xyzz_add(acc, y);
    }
    const jacobian r = xyzz_to_jacobian(acc);
    st_fe(&out[j].x, r.x);
    st_fe(&out[j].y, r.y);
    st_fe(&out[j].z, r.z);
}
__global__ void g1_to_affine_kernel(const jacobian* in, affine* out) {
    const affine r = xyzz_to_affine(jacobian_to_xyzz(ld_jacobian(in)));
    st_affine(out, r);
}

// out[i] = scalars[i] * G (affine): the Lagrange-basis SRS of a synthetic setup, L_i(tau) * G with the scalars L_i(tau) coming from
// an inverse transform of the powers of tau (api.cu: uzkge_cuda_srs_generate_lagrange).
__global__ void __launch_bounds__(128) g1_fixed_base_mul_kernel(const fe* __restrict__ scalars, affine g, uint32_t count, affine* out) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const fe s = fe_from_mont<FrP>(ld_fe(scalars + t));
    xyzz acc = xyzz_identity();
#pragma unroll 1
    for (int bit = 253; bit >= 0; bit--) {
        acc = xyzz_dbl(acc);
        if ((s.l[bit >> 5] >> (bit & 31)) & 1) xyzz_madd(acc, g);
    }
    st_affine(out + t, xyzz_to_affine(acc));
}

// *out (+)= sum_{j < k} scalars[j] * bases[idx[j]], k <= 32: one lane per term (double-and-add), warp-shuffle sum.  The blind
// factors of a commitment, apply_blind_factors (kzg_poly_commitment.rs:299-313): C += sum_i b_i * (SRS[i] - SRS[n + i]).
struct SmallMsmArgs {
    const affine* bases;
    uint32_t idx[32];
    fe scalars[32];   // Montgomery
    uint32_t k;
    uint32_t accumulate;
    jacobian* out;
};
__global__ void __launch_bounds__(32) g1_small_msm_kernel(const __grid_constant__ SmallMsmArgs a) {
    const uint32_t lane = threadIdx.x;
    xyzz acc = xyzz_identity();
    if (lane < a.k) {
        const affine p = ld_affine(a.bases + a.idx[lane]);
        const fe s = fe_from_mont<FrP>(a.scalars[lane]);
        if (!affine_is_identity(p)) {
#pragma unroll 1
            for (int bit = 253; bit >= 0; bit--) {
                acc = xyzz_dbl(acc);
                if ((s.l[bit >> 5] >> (bit & 31)) & 1) xyzz_madd(acc, p);
            }
        }
    }
    for (int d = 16; d; d >>= 1) {
        const xyzz o = shfl_down_xyzz(acc, d);
        xyzz_add(acc, o);
    }
    if (lane == 0) {
        if (a.accumulate) {
            const xyzz prev = jacobian_to_xyzz(ld_jacobian(a.out));
            xyzz_add(acc, prev);
        }
        const jacobian r = xyzz_to_jacobian(acc);
        st_fe(&a.out->x, r.x);
        st_fe(&a.out->y, r.y);
        st_fe(&a.out->z, r.z);
    }
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
    if (lg <= 12) return 10;   // 2^12: 296 us
    if (lg <= 15) return 13;   // 2^14: 350 us; 2^13 (zmatchmaking's batches of 5-8 commitments): proof 4.85 ms at c = 10, 4.63 at c = 13
    if (lg <= 16) return 15;   // 2^16: 510 us
    if (lg <= 18) return 17;   // 2^18: 1181 us
    return 20;                 // 2^20: 2.90 ms, 2^22: 10.4 ms
}

#define UZ_CUDA_TRY(expr)                                   \
    do {                                                    \
        cudaError_t e__ = (expr);                           \
        if (e__ != cudaSuccess) return cuda_err_code(e__);  \
    } while (0)

// Batch slots: how many independent MSMs one pass can carry (their bucket sets, reduction workspaces and sorted-entry
// segments live side by side).  Small SRS (PlonK-sized commitments, 16 per proof) get up to 16; large ones 1.
static uint32_t choose_batch_slots(size_t n, uint32_t nbuckets, uint32_t windows, size_t reduce_bytes) {
    const size_t per_slot = (size_t)nbuckets * (sizeof(xyzz) + 6 * 4) + reduce_bytes + (size_t)windows * n * 4;
    size_t slots = ((size_t)256 << 20) / (per_slot ? per_slot : 1);
    if (slots < 1) slots = 1;
    if (slots > MSM_MAX_BATCH) slots = MSM_MAX_BATCH;
    return (uint32_t)slots;
}

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
    s->nbuckets = (1u << (c - 1)) + 1;  // 0 (zero digit, always empty) .. 2^(c-1)
    const size_t reduce_bytes = msm_reduce_workspace_bytes(c, (uint32_t)sm_count_);
    if (reduce_bytes == 0) return UZKGE_ERR_SIZE;
    s->reduce_bytes = reduce_bytes;
    s->slots = choose_batch_slots(n, s->nbuckets, windows, reduce_bytes);
    const size_t m = (size_t)windows * n;          // entries of one full-size MSM
    const size_t m_all = m * s->slots;
    const size_t nb_all = (size_t)s->nbuckets * s->slots;
    if (m_all >= (1ull << 31)) return UZKGE_ERR_SIZE;
    // segments longer than max(8 * mean, 64 * lanes) are "large": at most nb_all / 8 of them can exist
    s->large_cap = (uint32_t)(nb_all / 8 + 2);
    s->max_slices = (uint32_t)(m_all / LARGE_SLICE_MIN + s->large_cap);

    size_t total = 0;
    auto take = [&](size_t bytes) {
        const size_t off = total;
        total += (bytes + 255) & ~(size_t)255;
        return off;
    };
    const size_t o_tables = take(sizeof(affine) * m);
    cub::DoubleBuffer<uint32_t> dk(nullptr, nullptr), dv(nullptr, nullptr);
    s->cub_temp_bytes = 0;
    UZ_CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, s->cub_temp_bytes, dk, dv, (int)nb_all, 0, 32, st));
    size_t scan_temp = 0;
    UZ_CUDA_TRY(cub::DeviceScan::ExclusiveSum(nullptr, scan_temp, (uint32_t*)nullptr, (uint32_t*)nullptr, (int)nb_all + 1, st));
    if (scan_temp > s->cub_temp_bytes) s->cub_temp_bytes = scan_temp;
    size_t o_vals[2], o_offsets[2], o_counts[2], o_ord[2], o_large[2], o_slice_start[2], o_slice_sums[2], o_buckets[2], o_reduce[2], o_ticket[2],
        o_cub[2];
    for (int l = 0; l < 2; l++) {
        o_vals[l] = take(4 * m_all);
        o_offsets[l] = take(4 * (nb_all + 1));
        o_counts[l] = take(4 * (nb_all + 1));
        o_ord[l] = take(4 * 4 * nb_all);
        o_large[l] = take(4 * ((size_t)s->large_cap + 1));
        o_slice_start[l] = take(4 * ((size_t)s->large_cap + 1));
        o_slice_sums[l] = take(sizeof(xyzz) * s->max_slices);
        o_buckets[l] = take(sizeof(xyzz) * nb_all);
        o_reduce[l] = take(reduce_bytes * s->slots);
        o_ticket[l] = take(256 * s->slots);
        o_cub[l] = take(s->cub_temp_bytes + 256);
    }
    UZ_CUDA_TRY(cudaMalloc(&s->arena, total));
    s->bytes = total;
    char* base = (char*)s->arena;
    s->tables = (affine*)(base + o_tables);
    for (int l = 0; l < 2; l++) {
        MsmWork& w = s->work[l];
        w.vals = (uint32_t*)(base + o_vals[l]);
        w.offsets = (uint32_t*)(base + o_offsets[l]);
        w.counts = (uint32_t*)(base + o_counts[l]);
        w.ord_keys_a = (uint32_t*)(base + o_ord[l]);
        w.ord_keys_b = w.ord_keys_a + nb_all;
        w.ord_vals_a = w.ord_keys_b + nb_all;
        w.ord_vals_b = w.ord_vals_a + nb_all;
        w.large_list = (uint32_t*)(base + o_large[l]);
        w.slice_start = (uint32_t*)(base + o_slice_start[l]);
        w.slice_sums = (xyzz*)(base + o_slice_sums[l]);
        w.buckets = (xyzz*)(base + o_buckets[l]);
        w.cub_temp = base + o_cub[l];
        for (uint32_t j = 0; j < s->slots; j++) {
            w.reduce[j] = msm_reduce_plan_create(c, (uint32_t)sm_count_, w.buckets + (size_t)j * s->nbuckets, base + o_reduce[l] + reduce_bytes * j,
                                                 (uint32_t*)(base + o_ticket[l] + 256 * j));
            if (!w.reduce[j]) return UZKGE_ERR_INTERNAL;
        }
        UZ_CUDA_TRY(cudaEventCreateWithFlags(&w.sorted, cudaEventDisableTiming));
        UZ_CUDA_TRY(cudaEventCreateWithFlags(&w.accumulated, cudaEventDisableTiming));
        UZ_CUDA_TRY(cudaEventCreateWithFlags(&w.reduced, cudaEventDisableTiming));
        UZ_CUDA_TRY(cudaMemsetAsync(base + o_ticket[l], 0, 256 * s->slots, st));
    }

    cudaEvent_t e0, e1;
    UZ_CUDA_TRY(cudaEventCreate(&e0));
    UZ_CUDA_TRY(cudaEventCreate(&e1));
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
    for (MsmWork& w : s->work) {
        for (uint32_t j = 0; j < MSM_MAX_BATCH; j++)
            if (w.reduce[j]) msm_reduce_plan_destroy(w.reduce[j]);
        if (w.aff_ws) cudaFree(w.aff_ws);
        if (w.sorted) cudaEventDestroy(w.sorted);
        if (w.accumulated) cudaEventDestroy(w.accumulated);
        if (w.reduced) cudaEventDestroy(w.reduced);
    }
    if (s->arena) cudaFree(s->arena);
    *s = MsmSrs();
}

MsmEngine::~MsmEngine() {
    for (cudaStream_t a : aux_) cudaStreamDestroy(a);
    for (cudaStream_t a : {s_sort_, s_acc_, s_red_})
        if (a) cudaStreamDestroy(a);
    if (start_) cudaEventDestroy(start_);
    if (aff_stream_) cudaStreamDestroy(aff_stream_);
    if (aff_fork_) cudaEventDestroy(aff_fork_);
    if (aff_join_) cudaEventDestroy(aff_join_);
    if (fork_) cudaEventDestroy(fork_);
    for (cudaEvent_t e : join_) cudaEventDestroy(e);
}

template <int G>
static cudaError_t launch_accumulate(const AccArgs& a, cudaStream_t st) {
    const uint64_t threads = (uint64_t)a.nb_padded * G;
    const uint32_t grid = (uint32_t)((threads + ACC_NT - 1) / ACC_NT);
    msm_accumulate_kernel<G><<<grid, ACC_NT, 2 * 4 * ACC_NT * sizeof(uint4), st>>>(a);
    return cudaGetLastError();
}

int MsmEngine::run(MsmSrs* s, size_t base_offset, const fe* d_scalars, size_t n, jacobian* d_out, cudaStream_t st) {
    const fe* sc[1] = {d_scalars};
    const size_t nn[1] = {n};
    return run_batch(s, base_offset, sc, nn, 1, d_out, st);
}

struct MsmEngine::GroupPlan {
    uint32_t k = 0, nb_all = 0, lanes = 1, thr = 0;
    uint64_t m = 0;        // sorted entries of the group
    double mean = 0;       // mean bucket load
    const uint32_t* order = nullptr;
    uint32_t n_max = 0;    // points of the largest MSM of the group
    bool empty = false;
};

// digits -> counting sort by bucket -> bucket ids in decreasing size.  Everything the accumulate kernel needs is left in `w`.
int MsmEngine::stage_sort(MsmSrs* s, MsmWork& w, size_t base_offset, const fe* const* d_scalars, const size_t* n, uint32_t k,
                          GroupPlan* plan, cudaStream_t st, int prof) {
    uint64_t n_all = 0;
    uint32_t n_max = 0;
    DigitArgs da;
    for (uint32_t j = 0; j < k; j++) {
        da.scalars[j] = d_scalars[j];
        da.n[j] = (uint32_t)n[j];
        n_all += n[j];
        if (n[j] > n_max) n_max = (uint32_t)n[j];
    }
    plan->k = k;
    plan->n_max = n_max;
    plan->empty = n_all == 0;
    if (plan->empty) return UZKGE_OK;
    da.c = s->c;
    da.windows = s->windows;
    da.table_stride = (uint32_t)s->n;
    da.base_offset = (uint32_t)base_offset;
    da.nbuckets = s->nbuckets;
    const uint32_t nb_all = s->nbuckets * k;
    const uint64_t m = (uint64_t)s->windows * n_all;

    // count -> scan -> scatter
    const size_t cbytes = 4 * ((size_t)nb_all + 1);
    const dim3 dgrid((n_max + 255) / 256, k);
    UZ_CUDA_TRY(cudaMemsetAsync(w.counts, 0, cbytes, st));
    msm_count_scatter_kernel<false><<<dgrid, 256, 0, st>>>(da, w.counts, nullptr, nullptr);
    UZ_CUDA_TRY(cudaGetLastError());
    g_prof.mark(prof, MSM_PH_RECODE, st);
    size_t temp = s->cub_temp_bytes;
    UZ_CUDA_TRY(cub::DeviceScan::ExclusiveSum(w.cub_temp, temp, w.counts, w.offsets, (int)nb_all + 1, st));
    UZ_CUDA_TRY(cudaMemsetAsync(w.counts, 0, cbytes, st));
    msm_count_scatter_kernel<true><<<dgrid, 256, 0, st>>>(da, w.counts, w.offsets, w.vals);
    UZ_CUDA_TRY(cudaGetLastError());
    UZ_CUDA_TRY(cudaMemsetAsync(w.large_list, 0, 4, st));
    g_prof.mark(prof, MSM_PH_SORT, st);

    // lanes per bucket: at most ~80 entries per lane; and for small problems (the prover's batches of 2^12..2^18-point MSMs), where
    // one wave of long per-lane chains is latency-bound, more lanes until about 2.75 x the resident threads are in flight,
    // keeping >= 5 entries per lane.  Measured with scripts/gpu_msm_lanes.py: 5 x 2^14: 1224 -> 623 us, 8 x 2^14: 1326 -> 846 us,
    // 2^18: 1184 -> 1034 us; 2^20 and up keep their lane count.
    const double mean = (double)m / (double)(nb_all - k);
    uint32_t g = 1;
    while (g < 32 && mean / (g * 2) >= 40.0) g *= 2;
    while (g < 32 && (uint64_t)nb_all * g * 2 <= (uint64_t)sm_count_ * 1408 && mean / (g * 2) >= 5.0) g *= 2;
    if (force_lanes_) g = force_lanes_;
    // segments above the threshold are split into warp slices (msm_large_*)
    uint32_t thr = (uint32_t)(mean * 8.0);
    {
        uint32_t per_lane = (uint32_t)(m / (2ull * sm_count_ * 512));
        if (per_lane < 64) per_lane = 64;
        if (thr < per_lane * g) thr = per_lane * g;
    }
    // visit the buckets in decreasing size
    uint32_t cap_bits = 1;
    while ((1u << cap_bits) <= thr + 1) cap_bits++;
    msm_sizes_kernel<<<(nb_all + 255) / 256, 256, 0, st>>>(w.offsets, nb_all, thr + 1, w.ord_keys_a, w.ord_vals_a);
    UZ_CUDA_TRY(cudaGetLastError());
    cub::DoubleBuffer<uint32_t> ok(w.ord_keys_a, w.ord_keys_b), ov(w.ord_vals_a, w.ord_vals_b);
    temp = s->cub_temp_bytes;
    UZ_CUDA_TRY(cub::DeviceRadixSort::SortPairs(w.cub_temp, temp, ok, ov, (int)nb_all, 0, (int)cap_bits, st));
    g_prof.mark(prof, MSM_PH_OFFSETS, st);
    plan->nb_all = nb_all;
    plan->m = m;
    plan->mean = mean;
    plan->lanes = g;
    plan->thr = thr;
    plan->order = ov.Current();
    UZ_COUNT_LAUNCH(3 + 2);  // own kernels + CUB's scan / sort launches
    return UZKGE_OK;
}

int MsmEngine::stage_accumulate(MsmSrs* s, MsmWork& w, const GroupPlan& plan, cudaStream_t st, int prof) {
    if (plan.empty) return UZKGE_OK;
    // large single MSMs: pairwise affine tree with shared inversions (6 products per addition instead of 10), then the XYZZ tail
    // (mode 2 forces it for any single MSM -- the test-suite's edge cases run at small sizes)
    const bool aff_ok = plan.k == 1 && (affine_mode_ == 2 || (affine_mode_ == 1 && plan.nb_all >= (1u << 16)));
    const uint32_t aff_rounds = aff_ok ? msm_affine_rounds(plan.mean) : 0;
    bool affine_done = false;
    if (aff_rounds >= (affine_mode_ == 2 ? 1u : 2u)) {
        const size_t need = msm_affine_workspace_bytes((uint64_t)s->windows * s->n, s->nbuckets);
        if (w.aff_ws_bytes < need) {
            if (w.aff_ws) cudaFree(w.aff_ws);
            w.aff_ws = nullptr;
            w.aff_ws_bytes = 0;
            if (cudaMalloc(&w.aff_ws, need) == cudaSuccess)
                w.aff_ws_bytes = need;
            else
                cudaGetLastError();   // no room for the affine arrays: the XYZZ kernel below does the work
        }
        if (w.aff_ws) {
            if (!aff_stream_) {
                UZ_CUDA_TRY(cudaStreamCreateWithFlags(&aff_stream_, cudaStreamNonBlocking));
                UZ_CUDA_TRY(cudaEventCreateWithFlags(&aff_fork_, cudaEventDisableTiming));
                UZ_CUDA_TRY(cudaEventCreateWithFlags(&aff_join_, cudaEventDisableTiming));
            }
            const int rc = msm_affine_accumulate(s, w, w.aff_ws, plan.m, plan.nb_all, plan.order, plan.thr, aff_rounds, st, aff_stream_,
                                                 aff_fork_, aff_join_);
            if (rc != UZKGE_OK) return rc;
            affine_done = true;
        }
    }
    AccArgs aa;
    aa.tables = s->tables;
    aa.vals = w.vals;
    aa.offsets = w.offsets;
    aa.order = plan.order;
    aa.buckets = w.buckets;
    aa.nb_padded = plan.nb_all;
    aa.large_threshold = plan.thr;
    aa.large_list = w.large_list;
    aa.large_cap = s->large_cap;
    cudaError_t e = cudaSuccess;
    if (!affine_done) {
        switch (plan.lanes) {
            case 1: e = launch_accumulate<1>(aa, st); break;
            case 2: e = launch_accumulate<2>(aa, st); break;
            case 4: e = launch_accumulate<4>(aa, st); break;
            case 8: e = launch_accumulate<8>(aa, st); break;
            case 16: e = launch_accumulate<16>(aa, st); break;
            default: e = launch_accumulate<32>(aa, st); break;
        }
    }
    UZ_CUDA_TRY(e);
    g_prof.mark(prof, MSM_PH_ACCUMULATE, st);

    LargeArgs la;
    la.tables = s->tables;
    la.vals = w.vals;
    la.offsets = w.offsets;
    la.buckets = w.buckets;
    la.large_list = w.large_list;
    la.slice_start = w.slice_start;
    la.slice_sums = w.slice_sums;
    la.large_cap = s->large_cap;
    la.max_slices = s->max_slices;
    // slice length: a bucket holds at most one entry per point, so n_max / 1024 keeps the slices of the largest possible bucket
    // (summed by ONE warp in msm_large_finish_kernel) around a thousand, while small problems -- a prover round's batch of 2^14-point
    // commitments over bit / small-integer witnesses -- get short dependent chains (4 points per lane instead of 32: measured
    // 286 -> us on zshuffle-52's round 1)
    la.slice = LARGE_SLICE_MIN;
    while (la.slice < LARGE_SLICE_MAX && la.slice < plan.n_max / 1024) la.slice <<= 1;
    // the list lengths are only known on the device: fixed grids whose warps walk the lists
    msm_large_plan_kernel<<<1, 1024, 0, st>>>(la);
    msm_large_accumulate_kernel<<<sm_count_ * 4, LARGE_NT, 2 * 4 * LARGE_NT * sizeof(uint4), st>>>(la);
    msm_large_finish_kernel<<<sm_count_ * 2, 128, 0, st>>>(la);
    UZ_CUDA_TRY(cudaGetLastError());
    g_prof.mark(prof, MSM_PH_LARGE, st);
    UZ_COUNT_LAUNCH(4);
    return UZKGE_OK;
}

// the k bucket reductions of a batch: one launch per level, the slot is the grid's second dimension (msm_reduce.cu)
int MsmEngine::stage_reduce(MsmSrs* s, MsmWork& w, uint32_t k, jacobian* d_out, cudaStream_t st) {
    return msm_reduce_run(w.reduce[0], k, w.buckets, s->nbuckets, s->reduce_bytes, d_out, st);
}

static int msm_check_group(const MsmSrs* s, size_t base_offset, const size_t* n, size_t k) {
    if (base_offset > s->n) return UZKGE_ERR_SIZE;
    for (size_t j = 0; j < k; j++)
        if (n[j] > s->n - base_offset) return UZKGE_ERR_SIZE;
    return UZKGE_OK;
}

// k <= s->slots independent MSMs over srs[base_offset ..] in one pass on `st`; d_out receives k Jacobian points
int MsmEngine::run_batch(MsmSrs* s, size_t base_offset, const fe* const* d_scalars, const size_t* n, uint32_t k, jacobian* d_out,
                         cudaStream_t st) {
    if (k == 0) return UZKGE_OK;
    if (k > s->slots || msm_check_group(s, base_offset, n, k) != UZKGE_OK) return UZKGE_ERR_SIZE;
    MsmWork& w = s->work[0];
    GroupPlan plan;
    const int prof = g_prof.begin(Profiler::MSM, st);
    int rc = stage_sort(s, w, base_offset, d_scalars, n, k, &plan, st, prof);
    if (rc != UZKGE_OK) return rc;
    if (plan.empty) {
        for (uint32_t j = 0; j < k; j++) msm_identity_kernel<<<1, 1, 0, st>>>(d_out + j);
        UZ_COUNT_LAUNCH(k);
        UZ_CUDA_TRY(cudaGetLastError());
        return UZKGE_OK;
    }
    rc = stage_accumulate(s, w, plan, st, prof);
    if (rc != UZKGE_OK) return rc;
    rc = stage_reduce(s, w, k, d_out, st);
    if (rc != UZKGE_OK) return rc;
    g_prof.mark(prof, MSM_PH_REDUCE, st);
    return UZKGE_OK;
}

// Any number of MSMs, in groups of s->slots.  Group g uses workspace g % 2 and three internal streams:
//   s_sort : sort(g)        after  the caller's stream at entry,  accumulate(g - 2)   (it read this workspace's entries)
//   s_acc  : accumulate(g)  after  sort(g),  reduce(g - 2)                          (it read this workspace's buckets)
//   s_red  : reduce(g)      after  accumulate(g)
// so the L2-atomic-bound sort and the latency-bound reduction run under the multiplier-bound accumulate kernel of a neighbour.
int MsmEngine::run_pipelined(MsmSrs* s, size_t base_offset, const fe* const* d_scalars, const size_t* n, size_t k, jacobian* d_out,
                             cudaStream_t st, const cudaEvent_t* ready) {
    if (k == 0) return UZKGE_OK;
    if (msm_check_group(s, base_offset, n, k) != UZKGE_OK) return UZKGE_ERR_SIZE;
    if (k <= s->slots) {
        if (ready) UZ_CUDA_TRY(cudaStreamWaitEvent(st, ready[k - 1], 0));
        return run_batch(s, base_offset, d_scalars, n, (uint32_t)k, d_out, st);
    }
    if (!s_sort_) {
        // the short sort / reduction kernels must not queue behind the thousands of pending CTAs of an accumulate launch
        int prio_lo = 0, prio_hi = 0;
        UZ_CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        UZ_CUDA_TRY(cudaStreamCreateWithPriority(&s_sort_, cudaStreamNonBlocking, prio_hi));
        UZ_CUDA_TRY(cudaStreamCreateWithPriority(&s_acc_, cudaStreamNonBlocking, prio_lo));
        UZ_CUDA_TRY(cudaStreamCreateWithPriority(&s_red_, cudaStreamNonBlocking, prio_hi));
        UZ_CUDA_TRY(cudaEventCreateWithFlags(&start_, cudaEventDisableTiming));
    }
    UZ_CUDA_TRY(cudaEventRecord(start_, st));
    UZ_CUDA_TRY(cudaStreamWaitEvent(s_sort_, start_, 0));
    UZ_CUDA_TRY(cudaStreamWaitEvent(s_red_, start_, 0));   // d_out may be in use by earlier work of the caller's stream
    size_t g = 0;
    for (size_t j0 = 0; j0 < k; j0 += s->slots, g++) {
        const uint32_t kk = (uint32_t)((k - j0) < s->slots ? (k - j0) : s->slots);
        MsmWork& w = s->work[g & 1];
        GroupPlan plan;
        if (g >= 2) UZ_CUDA_TRY(cudaStreamWaitEvent(s_sort_, w.accumulated, 0));
        if (ready) UZ_CUDA_TRY(cudaStreamWaitEvent(s_sort_, ready[j0 + kk - 1], 0));   // events of one stream fire in order
        int rc = stage_sort(s, w, base_offset, d_scalars + j0, n + j0, kk, &plan, s_sort_, -1);
        if (rc != UZKGE_OK) return rc;
        UZ_CUDA_TRY(cudaEventRecord(w.sorted, s_sort_));
        UZ_CUDA_TRY(cudaStreamWaitEvent(s_acc_, w.sorted, 0));
        if (g >= 2) UZ_CUDA_TRY(cudaStreamWaitEvent(s_acc_, w.reduced, 0));
        rc = stage_accumulate(s, w, plan, s_acc_, -1);
        if (rc != UZKGE_OK) return rc;
        UZ_CUDA_TRY(cudaEventRecord(w.accumulated, s_acc_));
        UZ_CUDA_TRY(cudaStreamWaitEvent(s_red_, w.accumulated, 0));
        if (plan.empty) {
            for (uint32_t j = 0; j < kk; j++) msm_identity_kernel<<<1, 1, 0, s_red_>>>(d_out + j0 + j);
            UZ_COUNT_LAUNCH(kk);
            UZ_CUDA_TRY(cudaGetLastError());
        } else {
            rc = stage_reduce(s, w, kk, d_out + j0, s_red_);
            if (rc != UZKGE_OK) return rc;
        }
        UZ_CUDA_TRY(cudaEventRecord(w.reduced, s_red_));
    }
    // the caller's stream continues once every group is reduced (s_red_ is in order: its last event covers them all)
    UZ_CUDA_TRY(cudaStreamWaitEvent(st, s->work[(g - 1) & 1].reduced, 0));
    if (g >= 2) UZ_CUDA_TRY(cudaStreamWaitEvent(st, s->work[g & 1].reduced, 0));
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

int MsmEngine::fixed_base_mul(const fe* d_scalars, uint32_t count, affine* d_out, cudaStream_t st) {
    affine g;
    g.x = fe_one<FqP>();
    g.y = fe_dbl<FqP>(g.x);
    g1_fixed_base_mul_kernel<<<(count + 127) / 128, 128, 0, st>>>(d_scalars, g, count, d_out);
    UZ_COUNT_LAUNCH(1);
    UZ_CUDA_TRY(cudaGetLastError());
    return UZKGE_OK;
}

int MsmEngine::small_msm(const MsmSrs* s, const size_t* idx, const uint64_t* scalars, uint32_t k, bool accumulate, jacobian* d_out,
                         cudaStream_t st) {
    if (k > 32) return UZKGE_ERR_SIZE;
    SmallMsmArgs a;
    memset(&a, 0, sizeof(a));
    for (uint32_t j = 0; j < k; j++) {
        if (idx[j] >= s->n) return UZKGE_ERR_SIZE;
        a.idx[j] = (uint32_t)idx[j];
        memcpy(&a.scalars[j], scalars + 4 * j, sizeof(fe));
    }
    a.bases = s->tables;   // table 0 = the bases themselves
    a.k = k;
    a.accumulate = accumulate ? 1 : 0;
    a.out = d_out;
    g1_small_msm_kernel<<<1, 32, 0, st>>>(a);
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
int MsmEngine::g1_sum(const jacobian* d_parts, uint32_t count, uint32_t stride, uint32_t k, jacobian* d_out, cudaStream_t st) {
    if (count == 0 || k == 0) return UZKGE_ERR_SIZE;
    g1_sum_kernel<<<(k + 31) / 32, 32, 0, st>>>(d_parts, count, stride, k, d_out);
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
