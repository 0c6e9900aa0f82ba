// msm_reduce.cu -- the bucket reduction  sum_b b * B_b  of the MSM (see msm.cu for the pipeline).
//
// Cost model (measured on B200): one group operation issued by a warp occupies that SM sub-partition's integer
// multiplier for ~4 us (14 Montgomery products x 137 IMAD) whether 1 or 32 lanes are active, and a single warp
// chaining dependent operations sees 6-8 us each.  The reduction therefore avoids long dependent chains and,
// where few additions are available at a time, splits ONE addition over a TEAM of 4 lanes (the four independent
// products of each level of the XYZZ formulas run in different lanes: ~1.5 us per addition, 8 additions per
// warp at a time).
//
// Algorithm: no scalar multiplications, no running sums.  An array X[0..m) (m a power of two, weight = index) is
// viewed as a rows x cols matrix, j = hi * cols + lo:
//       W(X) = sum_j j X[j] = cols * W(R) + W(C),   R[hi] = sum_lo X[hi][lo],  C[lo] = sum_hi X[hi][lo]
// and the same identity is applied to R and C until every array has <= 32 entries (2 or 3 levels).  A leaf array
// is finished by bit decomposition: W(X) = sum_l 2^l * (sum of the X[j] with bit l of j set); every (leaf, bit)
// pair is one warp: masked tree sum, then its own doublings -- all pairs in parallel -- and the last CTA adds
// the partial results.  The first level on the bucket array is throughput-bound (2 additions per bucket): its
// sums are first taken over short strips by one thread each (msm_strips_kernel, fully inlined arithmetic).
#include <cuda_runtime.h>

#include <vector>

#include "devmem.cuh"
#include "ec_team.cuh"
#include "internal.h"

namespace uz {

static constexpr int SUM_NT = 128;  // 4 warps per CTA, one per SM sub-partition

// The k reductions of a batch run in ONE launch per level: blockIdx.y is the slot, and every pointer of slot 0's plan is moved into
// slot j's copy of the same layout -- by j * bucket_stride when it points into the bucket array, else by j * ws_stride (the
// reduction workspaces of the slots are laid out identically, msm.cu: MsmEngine::upload).
struct SlotMap {
    const xyzz* bucket_lo;
    const xyzz* bucket_hi;
    size_t bucket_stride, ws_stride;   // in xyzz units
    uint32_t ticket_stride;            // in uint32 units
};
template <class T>
__device__ __forceinline__ T* slot_ptr(T* p, const SlotMap& m, uint32_t j) {
    const xyzz* q = (const xyzz*)p;
    return (T*)(q + (size_t)j * ((q >= m.bucket_lo && q < m.bucket_hi) ? m.bucket_stride : m.ws_stride));
}

// ------------------------------------------------------------------ level 0: strip sums over the buckets
// thread t < rows * nq : RP[hi][q] = sum of the LR consecutive entries of row hi starting at q * LR
// thread t >= rows * nq: CP[s][lo] = sum of the LC entries of column lo in rows s * LC ..
struct StripArgs {
    const xyzz* src;
    xyzz* rp;
    xyzz* cp;
    uint32_t rows, cols, lr, lc;
};
__global__ void __launch_bounds__(256, 2) msm_strips_kernel(StripArgs a, const SlotMap sm) {
    a.src = slot_ptr(a.src, sm, blockIdx.y);
    a.rp = slot_ptr(a.rp, sm, blockIdx.y);
    a.cp = slot_ptr(a.cp, sm, blockIdx.y);
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nq = a.cols / a.lr, ns = a.rows / a.lc;
    const uint32_t n_rp = a.rows * nq;
    const xyzz* p;
    uint32_t count, stride;
    xyzz* dst;
    if (t < n_rp) {
        p = a.src + (size_t)t * a.lr;  // (hi * nq + q) * lr = hi * cols + q * lr
        count = a.lr;
        stride = 1;
        dst = a.rp + t;
    } else if (t - n_rp < ns * a.cols) {
        const uint32_t u = t - n_rp, s = u / a.cols, lo = u % a.cols;
        p = a.src + (size_t)s * a.lc * a.cols + lo;
        count = a.lc;
        stride = a.cols;
        dst = a.cp + u;
    } else {
        return;
    }
    xyzz acc = ld_xyzz(p);
#pragma unroll 1
    for (uint32_t e = 1; e < count; e++) {
        const xyzz v = ld_xyzz(p + (size_t)e * stride);
        xyzz_add(acc, v);
    }
    st_xyzz(dst, acc);
}

// ------------------------------------------------------------------ sums: one warp per output
struct SumFamily {
    const xyzz* src;
    xyzz* dst;
    uint32_t n_out, count, ostride, estride;  // dst[o] = sum_{e < count} src[o * ostride + e * estride]
};
struct SumsArgs {
    SumFamily f[8];
    uint32_t nfam, total;
};
__global__ void __launch_bounds__(SUM_NT) msm_sums_kernel(const SumsArgs a, const SlotMap sm) {
    __shared__ xyzz bufs[SUM_NT / 32][32];
    uint32_t o = blockIdx.x * (SUM_NT / 32) + (threadIdx.x >> 5);
    const uint32_t lane = threadIdx.x & 31;
    if (o >= a.total) return;  // warp-uniform
    uint32_t k = 0;
    while (o >= a.f[k].n_out) {
        o -= a.f[k].n_out;
        k++;
    }
    const SumFamily& f = a.f[k];
    const xyzz* base = slot_ptr(f.src, sm, blockIdx.y) + (size_t)o * f.ostride;
    xyzz acc = xyzz_identity();
    if (lane < f.count) acc = ld_xyzz(base + (size_t)lane * f.estride);
#pragma unroll 1
    for (uint32_t e = lane + 32; e < f.count; e += 32) acc = xyzz_add_call(acc, ld_xyzz(base + (size_t)e * f.estride));
    xyzz* buf = bufs[threadIdx.x >> 5];
    buf[lane] = acc;
    __syncwarp();
    uint32_t n = 1;
    while (n < f.count && n < 32) n <<= 1;
    warp_team_tree(buf, n);
    if (lane == 0) st_xyzz(slot_ptr(f.dst, sm, blockIdx.y) + o, buf[0]);
}

// ------------------------------------------------------------------ leaves: one warp per (array, bit) item
struct FinalItem {
    const xyzz* src;
    uint32_t len;    // <= 32
    int32_t bit;     // select entries whose index has this bit set; -1: all entries
    uint32_t shift;  // multiply the sum by 2^shift
};
static constexpr uint32_t MAX_ITEMS = 64;
struct FinalArgs {
    FinalItem it[MAX_ITEMS];
    uint32_t nitems;
    xyzz* partial;  // nitems
    uint32_t* ticket;
    jacobian* out;
};
__global__ void __launch_bounds__(SUM_NT) msm_leaves_kernel(const FinalArgs a, const SlotMap sm) {
    xyzz* const partial = slot_ptr(a.partial, sm, blockIdx.y);
    uint32_t* const ticket = a.ticket + (size_t)blockIdx.y * sm.ticket_stride;
    jacobian* const out = a.out + blockIdx.y;
    __shared__ xyzz bufs[SUM_NT / 32][32];
    __shared__ uint32_t is_last;
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t w = blockIdx.x * (SUM_NT / 32) + warp;
    xyzz* buf = bufs[warp];
    if (w < a.nitems) {
        const FinalItem& it = a.it[w];
        xyzz v = xyzz_identity();
        if (lane < it.len && (it.bit < 0 || ((lane >> it.bit) & 1))) v = ld_xyzz(slot_ptr(it.src, sm, blockIdx.y) + lane);
        buf[lane] = v;
        __syncwarp();
        uint32_t n = 1;
        while (n < it.len) n <<= 1;
        warp_team_tree(buf, n);
        v = buf[0];
#pragma unroll 1
        for (uint32_t i = 0; i < it.shift; i++) v = team4_dbl(v);
        if (lane == 0) st_xyzz(partial + w, v);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        is_last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last || warp != 0) return;
    __threadfence();
    xyzz v = xyzz_identity();
    if (lane < a.nitems) v = ld_xyzz(partial + lane);
    if (lane + 32 < a.nitems) v = xyzz_add_call(v, ld_xyzz(partial + lane + 32));
    buf[lane] = v;
    __syncwarp();
    warp_team_tree(buf, 32);
    if (lane == 0) {
        const jacobian j = xyzz_to_jacobian<FqCall>(buf[0]);
        st_fe(&out->x, j.x);
        st_fe(&out->y, j.y);
        st_fe(&out->z, j.z);
        *ticket = 0;
    }
}

// ------------------------------------------------------------------ host side: the plan is fixed per window size
struct ReducePlan {
    bool has_strips = false;
    StripArgs strips;
    uint32_t strip_threads = 0;
    std::vector<SumsArgs> passes;
    FinalArgs fin;
};

static uint32_t ilog2(uint32_t x) {
    uint32_t l = 0;
    while ((1u << (l + 1)) <= x) l++;
    return l;
}

// layout of the workspace (in xyzz units) for window size c; plan == nullptr only measures
static size_t plan_layout(uint32_t c, uint32_t sm_count, xyzz* buckets, xyzz* ws, uint32_t* ticket, ReducePlan* plan) {
    struct Arr {
        xyzz* ptr;
        uint32_t len, shift;
    };
    size_t used = 0;
    auto take = [&](size_t n) {
        xyzz* p = ws ? ws + used : nullptr;
        used += n;
        return p;
    };
    std::vector<Arr> queue{{buckets, 1u << (c - 1), 0}};
    bool first = true;
    for (;;) {
        bool any = false;
        for (const Arr& a : queue) any = any || a.len > 32;
        if (!any) break;
        std::vector<Arr> next;
        SumsArgs sa;
        sa.nfam = 0;
        sa.total = 0;
        std::vector<SumsArgs> flushed;
        auto push_family = [&](const SumFamily& f) {
            if (sa.nfam == 8) {
                flushed.push_back(sa);
                sa.nfam = 0;
                sa.total = 0;
            }
            sa.f[sa.nfam++] = f;
            sa.total += f.n_out;
        };
        for (const Arr& a : queue) {
            if (a.len <= 32) {
                next.push_back(a);
                continue;
            }
            const uint32_t loglen = ilog2(a.len), logcols = (loglen + 1) / 2;
            const uint32_t cols = 1u << logcols, rows = a.len >> logcols;
            xyzz* R = take(rows);
            xyzz* C = take(cols);
            if (first && a.len >= (1u << 15)) {
                // level 0 over the buckets: strips so that every marginal has <= 64 partial sums and the strip
                // threads fit ONE wave of 2 x 256 threads per SM (the kernel is multiplier-bound: a partial second
                // wave only idles SMs -- measured 318 us at 1.3 waves for c = 20)
                uint32_t len = 4;
                while ((uint64_t)2 * a.len / len > (uint64_t)sm_count * 512) len <<= 1;
                while (cols / len > 64 || rows / len > 64) len <<= 1;
                const uint32_t lr = len < cols ? len : cols, lc = len < rows ? len : rows;
                const uint32_t nq = cols / lr, ns = rows / lc;
                xyzz* rp = take((size_t)rows * nq);
                xyzz* cp = take((size_t)ns * cols);
                if (plan) {
                    plan->has_strips = true;
                    plan->strips = StripArgs{a.ptr, rp, cp, rows, cols, lr, lc};
                    plan->strip_threads = rows * nq + ns * cols;
                }
                push_family(SumFamily{rp, R, rows, nq, nq, 1});
                push_family(SumFamily{cp, C, cols, ns, 1, cols});
            } else {
                push_family(SumFamily{a.ptr, R, rows, cols, cols, 1});
                push_family(SumFamily{a.ptr, C, cols, rows, 1, cols});
            }
            next.push_back(Arr{R, rows, a.shift + logcols});
            next.push_back(Arr{C, cols, a.shift});
        }
        if (plan) {
            for (const SumsArgs& f : flushed) plan->passes.push_back(f);
            if (sa.nfam) plan->passes.push_back(sa);
        }
        queue = next;
        first = false;
    }
    // leaves -> (array, bit) items; the top bucket 2^(c-1) sits right after the matrix
    uint32_t nitems = 0;
    auto add_item = [&](const FinalItem& it) {
        if (plan && nitems < MAX_ITEMS) plan->fin.it[nitems] = it;
        nitems++;
    };
    for (const Arr& a : queue)
        for (uint32_t l = 0; (1u << l) < a.len; l++) add_item(FinalItem{a.ptr, a.len, (int32_t)l, a.shift + l});
    add_item(FinalItem{buckets ? buckets + ((size_t)1 << (c - 1)) : nullptr, 1, -1, c - 1});
    xyzz* partial = take(MAX_ITEMS);
    if (plan) {
        plan->fin.nitems = nitems;
        plan->fin.partial = partial;
        plan->fin.ticket = ticket;
        plan->fin.out = nullptr;
    }
    if (nitems > MAX_ITEMS) return 0;  // cannot happen for c <= 24 (checked by the caller)
    return used;
}

size_t msm_reduce_workspace_bytes(uint32_t c, uint32_t sm_count) {
    return plan_layout(c, sm_count, nullptr, nullptr, nullptr, nullptr) * sizeof(xyzz);
}

ReducePlan* msm_reduce_plan_create(uint32_t c, uint32_t sm_count, xyzz* buckets, void* workspace, uint32_t* ticket) {
    ReducePlan* p = new ReducePlan();
    if (plan_layout(c, sm_count, buckets, (xyzz*)workspace, ticket, p) == 0) {
        delete p;
        return nullptr;
    }
    return p;
}
void msm_reduce_plan_destroy(ReducePlan* p) { delete p; }

// k reductions (slots 0 .. k - 1 of one workspace, plan of slot 0) in one launch per level; d_out receives k consecutive points
int msm_reduce_run(const ReducePlan* p, uint32_t k, const xyzz* buckets, size_t bucket_stride, size_t ws_stride_bytes, jacobian* d_out,
                   cudaStream_t st) {
    if (k == 0) return UZKGE_OK;
    SlotMap sm;
    sm.bucket_lo = buckets;
    sm.bucket_hi = buckets + bucket_stride;
    sm.bucket_stride = bucket_stride;
    sm.ws_stride = ws_stride_bytes / sizeof(xyzz);
    sm.ticket_stride = 256 / 4;
    uint32_t launches = 0;
    if (p->has_strips) {
        msm_strips_kernel<<<dim3((p->strip_threads + 255) / 256, k), 256, 0, st>>>(p->strips, sm);
        launches++;
    }
    for (const SumsArgs& sa : p->passes) {
        msm_sums_kernel<<<dim3((sa.total + SUM_NT / 32 - 1) / (SUM_NT / 32), k), SUM_NT, 0, st>>>(sa, sm);
        launches++;
    }
    FinalArgs fa = p->fin;
    fa.out = d_out;
    msm_leaves_kernel<<<dim3((fa.nitems + SUM_NT / 32 - 1) / (SUM_NT / 32), k), SUM_NT, 0, st>>>(fa, sm);
    launches++;
    UZ_COUNT_LAUNCH(launches);
    return cudaGetLastError() == cudaSuccess ? UZKGE_OK : UZKGE_ERR_CUDA;
}

}  // namespace uz
