// msm_affine.cu -- bucket accumulation by batched AFFINE additions (a pairwise tree per bucket) for large MSMs.
//
// The XYZZ accumulate kernel (msm.cu) runs at the multiplier roof with 10 products per mixed addition, so only fewer products
// per addition make it faster.  An affine addition costs 1 inversion + 2M + 1S; with Montgomery's trick the inversion of a whole
// batch costs 3 products per element plus ONE shared inversion: 6 products per addition.  Per tree round r (three launches):
//   every thread owns the pairs (2j, 2j+1) of one bucket (largest first, so the lanes of a warp carry equal work);
//   pass A  walks its pairs, forms the denominators d = x2 - x1 (substituting for the exceptional cases), stores the running
//           prefix products in HBM scratch and the thread total;
//   invert  the thread totals are inverted together: chunks of 32 per thread (prefix products, one Fermat inversion, back
//           substitution) -- 1/32 inversion per thread of the round;
//   pass B  walks the pairs backwards, peels 1/d off the running inverse (2M), finishes the addition (2M + 1S) and writes the
//           sum to the round's point array; an odd leftover point is copied.
// (A first version kept the three steps in one kernel with one inversion per CTA behind a barrier: every CTA then idles for the
// ~120 us of a Fermat chain per round -- 2^20: 5.3 ms against 2.06 ms for the XYZZ kernel, fmaheavy 22-34 %.)
// Bucket b's points of round r live at off_r(b) = ceil(off_{r-1}(b) / 2) + b, count ceil(count_{r-1} / 2): computable from the
// sorted offsets alone, no scans between rounds.  After R rounds the <= ~3 points left per bucket are summed in XYZZ into the
// bucket array the reduction consumes.  Exceptional pairs are handled exactly: an identity operand (x = y = 0: padded SRS
// entries, cancelled pairs) copies the other point, P + P doubles (d = 2y, numerator 3x^2), P + (-P) yields the identity.
// Buckets above the oversized threshold are left to the msm_large_* kernels, as in the XYZZ path.
#include <cuda_runtime.h>

#include "devmem.cuh"
#include "ec_compact.cuh"
#include "internal.h"

namespace uz {

static constexpr int AFF_NT = 128;

struct AffArgs {
    const affine* tables;
    const uint32_t* vals;
    const uint32_t* offsets;
    const uint32_t* order;
    const affine* src;   // points of round - 1 (unused in round 1: they are table references)
    affine* dst;         // points of this round
    fe* pre;             // prefix products, same geometry as dst
    fe* tot;             // per-thread (= per-bucket-slot) totals of pass A
    fe* tot_inv;         // their inverses
    xyzz* buckets;       // tail only
    uint32_t* large_list;
    uint32_t large_cap;
    uint32_t nb;         // bucket ids
    uint32_t round;      // 1-based; tail: number of rounds done
    uint32_t large_threshold;
    uint32_t phase, stride;   // this launch owns the slots phase, phase + stride, ... (the two halves run on two streams)
    uint32_t nslots;          // number of slots it owns
};

struct Seg {
    uint32_t cnt0, cnt_prev, off_prev, cnt_cur, off_cur;
};
__device__ __forceinline__ Seg seg_of(const uint32_t* __restrict__ offsets, uint32_t b, uint32_t round) {
    Seg s;
    uint32_t off = offsets[b], cnt = offsets[b + 1] - off;
    s.cnt0 = cnt;
    s.cnt_prev = cnt;
    s.off_prev = off;
    for (uint32_t r = 0; r < round; r++) {
        s.cnt_prev = cnt;
        s.off_prev = off;
        off = ((off + 1) >> 1) + b;
        cnt = (cnt + 1) >> 1;
    }
    s.cnt_cur = cnt;
    s.off_cur = off;
    return s;
}

template <bool FIRST>
__device__ __forceinline__ affine fetch_point(const AffArgs& a, uint32_t pos) {
    if (FIRST) {
        const uint32_t v = a.vals[pos];
        affine p = ld_affine(a.tables + (v & 0x7fffffffu));
        if (v >> 31) p = affine_neg(p);
        return p;
    }
    return ld_affine(a.src + pos);
}

template <bool FIRST>
__device__ __forceinline__ fe fetch_x(const AffArgs& a, uint32_t pos) {
    if (FIRST) return ld_fe(&a.tables[a.vals[pos] & 0x7fffffffu].x);
    return ld_fe(&a.src[pos].x);
}

// 0: generic (d = x2 - x1), 1: an operand is the identity (copy the other), 2: doubling (d = 2 y1), 3: cancellation
__device__ __forceinline__ int pair_denominator(const affine& p, const affine& q, fe& d) {
    if (affine_is_identity(p) || affine_is_identity(q)) {
        d = fe_one<FqP>();
        return 1;
    }
    d = fe_sub<FqP>(q.x, p.x);
    if (fe_is_zero(d)) {
        if (fe_is_zero(fe_sub<FqP>(q.y, p.y))) {
            d = fe_dbl<FqP>(p.y);
            return 2;
        }
        d = fe_one<FqP>();
        return 3;
    }
    return 0;
}

template <bool FIRST>
__global__ void __launch_bounds__(AFF_NT) msm_affine_pass_a_kernel(const AffArgs a) {
    const uint32_t tid = blockIdx.x * AFF_NT + threadIdx.x;
    if (tid >= a.nslots) return;
    const uint32_t slot = a.phase + tid * a.stride;
    const uint32_t b = a.order[slot];
    const Seg s = seg_of(a.offsets, b, a.round);
    fe prod = fe_one<FqP>();
    if (s.cnt0 <= a.large_threshold) {
        const uint32_t npairs = s.cnt_prev >> 1;
#pragma unroll 1
        for (uint32_t j = 0; j < npairs; j++) {
            // the x coordinates decide everything but the exceptional cases: half the gather traffic of the full points
            const fe px = fetch_x<FIRST>(a, s.off_prev + 2 * j);
            const fe qx = fetch_x<FIRST>(a, s.off_prev + 2 * j + 1);
            fe d = fe_sub<FqP>(qx, px);
            if (fe_is_zero(px) || fe_is_zero(qx) || fe_is_zero(d)) {
                const affine p = fetch_point<FIRST>(a, s.off_prev + 2 * j);
                const affine q = fetch_point<FIRST>(a, s.off_prev + 2 * j + 1);
                pair_denominator(p, q, d);
            }
            st_fe(a.pre + s.off_cur + j, prod);
            prod = fe_mul<FqP>(prod, d);
        }
    }
    st_fe(a.tot + tid, prod);
}

// inv[i] = 1 / tot[i] (all non-zero), chunks of AFF_CHUNK values per thread with one inversion each; `inv` doubles as scratch
static constexpr uint32_t AFF_CHUNK = 32;
__global__ void __launch_bounds__(64) msm_affine_invert_kernel(const fe* __restrict__ tot, fe* __restrict__ inv, uint32_t n) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lo = t * AFF_CHUNK;
    if (lo >= n) return;
    const uint32_t hi = min(lo + AFF_CHUNK, n);
    fe acc = fe_one<FqP>();
#pragma unroll 1
    for (uint32_t i = lo; i < hi; i++) {
        st_fe(inv + i, acc);                    // exclusive prefix product
        acc = fq_mul_call(acc, ld_fe(tot + i));
    }
    acc = fe_inv<FqP>(acc);
#pragma unroll 1
    for (uint32_t i = hi; i > lo; i--) {
        const fe pre = ld_fe(inv + i - 1);
        st_fe(inv + i - 1, fq_mul_call(acc, pre));
        acc = fq_mul_call(acc, ld_fe(tot + i - 1));
    }
}

template <bool FIRST>
__global__ void __launch_bounds__(AFF_NT) msm_affine_pass_b_kernel(const AffArgs a) {
    const uint32_t tid = blockIdx.x * AFF_NT + threadIdx.x;
    if (tid >= a.nslots) return;
    const uint32_t slot = a.phase + tid * a.stride;
    const uint32_t b = a.order[slot];
    const Seg s = seg_of(a.offsets, b, a.round);
    if (s.cnt0 > a.large_threshold) return;
    fe inv = ld_fe(a.tot_inv + tid);
    const uint32_t npairs = s.cnt_prev >> 1;
    if (s.cnt_prev & 1) st_affine(a.dst + s.off_cur + npairs, fetch_point<FIRST>(a, s.off_prev + s.cnt_prev - 1));
#pragma unroll 1
    for (int j = (int)npairs - 1; j >= 0; j--) {
        const affine p = fetch_point<FIRST>(a, s.off_prev + 2 * j);
        const affine q = fetch_point<FIRST>(a, s.off_prev + 2 * j + 1);
        fe d;
        const int kind = pair_denominator(p, q, d);
        const fe dinv = fe_mul<FqP>(inv, ld_fe(a.pre + s.off_cur + j));
        inv = fe_mul<FqP>(inv, d);
        affine r;
        if (kind == 1) {
            r = affine_is_identity(p) ? q : p;
        } else if (kind == 3) {
            r.x = fe_zero();
            r.y = fe_zero();
        } else {
            fe num;
            if (kind == 2) {
                const fe xx = fe_sqr<FqP>(p.x);
                num = fe_add<FqP>(fe_dbl<FqP>(xx), xx);
            } else {
                num = fe_sub<FqP>(q.y, p.y);
            }
            const fe lam = fe_mul<FqP>(num, dinv);
            r.x = fe_sub<FqP>(fe_sub<FqP>(fe_sqr<FqP>(lam), p.x), q.x);
            r.y = fe_sub<FqP>(fe_mul<FqP>(lam, fe_sub<FqP>(p.x, r.x)), p.y);
        }
        st_affine(a.dst + s.off_cur + j, r);
    }
}

// the few points left per bucket after the last round, summed in XYZZ into the bucket array; oversized buckets are listed
// for the msm_large_* kernels exactly as msm_accumulate_kernel does
__global__ void __launch_bounds__(128) msm_affine_tail_kernel(const AffArgs a) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= a.nb) return;
    const uint32_t b = a.order[slot];
    const Seg s = seg_of(a.offsets, b, a.round);
    if (s.cnt0 > a.large_threshold) {
        const uint32_t k = atomicAdd(a.large_list, 1u);
        if (k < a.large_cap) a.large_list[1 + k] = b;
        return;
    }
    xyzz acc = xyzz_identity();
#pragma unroll 1
    for (uint32_t j = 0; j < s.cnt_cur; j++) xyzz_madd(acc, ld_affine(a.src + s.off_cur + j));
    st_xyzz(a.buckets + b, acc);
}

// ------------------------------------------------------------------ host side
size_t msm_affine_workspace_bytes(uint64_t m, uint32_t nb) {
    const size_t a = (size_t)(m / 2 + 2ull * nb + 64), b = (size_t)(m / 4 + 2ull * nb + 64);
    return (a + b) * sizeof(affine) + (a + 2ull * nb + 512) * sizeof(fe) + 1024;
}

// rounds worth doing for a mean bucket load.  A round costs a fixed ~0.24 ms (the latency of one Fermat chain in the inversion
// kernel) on top of 6 products per addition; measured at 2^22 (load 104), rounds with fewer than ~6 M pairs lose against the XYZZ
// kernel, so the tree stops while ~10 points per bucket are left and the XYZZ tail sums those.
uint32_t msm_affine_rounds(double mean_load) {
    uint32_t r = 0;
    while (r < 6 && mean_load / (double)(1u << (r + 1)) >= 10.0) r++;
    return r;
}

int msm_affine_accumulate(const MsmSrs* s, const MsmWork& w, void* aff_ws, uint64_t m, uint32_t nb, const uint32_t* order, uint32_t thr,
                          uint32_t rounds, cudaStream_t st, cudaStream_t st2, cudaEvent_t fork, cudaEvent_t join) {
    const size_t na = (size_t)(m / 2 + 2ull * nb + 64), nbb = (size_t)(m / 4 + 2ull * nb + 64);
    affine* buf_a = (affine*)aff_ws;
    affine* buf_b = buf_a + na;
    fe* pre = (fe*)(buf_b + nbb);
    AffArgs a;
    a.tables = s->tables;
    a.vals = w.vals;
    a.offsets = w.offsets;
    a.order = order;
    a.pre = pre;
    a.buckets = w.buckets;
    a.large_list = w.large_list;
    a.large_cap = s->large_cap;
    a.nb = nb;
    a.large_threshold = thr;
    // One chain  A -> invert -> B  per round on one stream.  (Dealing the slots to two halves on two streams, so that the
    // latency-bound inversion kernel of one half runs under the passes of the other, needs separate point / prefix arrays per
    // half -- a half running a round ahead overwrites the other's layouts -- and gained nothing without a deliberate skew:
    // both halves reach their inversions at the same time.)
    (void)st2;
    (void)fork;
    (void)join;
    a.phase = 0;
    a.stride = 1;
    a.nslots = nb;
    a.tot = pre + na;
    a.tot_inv = pre + na + (size_t)nb + 128;
    const affine* src = nullptr;
    const uint32_t grid = (nb + AFF_NT - 1) / AFF_NT;
    const uint32_t inv_threads = (nb + AFF_CHUNK - 1) / AFF_CHUNK;
    for (uint32_t r = 1; r <= rounds; r++) {
        a.round = r;
        a.src = src;
        a.dst = (r & 1) ? buf_a : buf_b;
        if (r == 1)
            msm_affine_pass_a_kernel<true><<<grid, AFF_NT, 0, st>>>(a);
        else
            msm_affine_pass_a_kernel<false><<<grid, AFF_NT, 0, st>>>(a);
        msm_affine_invert_kernel<<<(inv_threads + 63) / 64, 64, 0, st>>>(a.tot, a.tot_inv, nb);
        if (r == 1)
            msm_affine_pass_b_kernel<true><<<grid, AFF_NT, 0, st>>>(a);
        else
            msm_affine_pass_b_kernel<false><<<grid, AFF_NT, 0, st>>>(a);
        src = a.dst;
    }
    a.round = rounds;
    a.src = (rounds & 1) ? buf_a : buf_b;
    a.dst = nullptr;
    a.phase = 0;
    a.stride = 1;
    a.nslots = nb;
    msm_affine_tail_kernel<<<(nb + 127) / 128, 128, 0, st>>>(a);
    UZ_COUNT_LAUNCH(3 * rounds + 1);
    return cudaGetLastError() == cudaSuccess ? UZKGE_OK : UZKGE_ERR_CUDA;
}

}  // namespace uz
