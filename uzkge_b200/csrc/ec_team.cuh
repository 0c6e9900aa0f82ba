// ec_team.cuh -- group operations split over a TEAM of 4 lanes, and a warp-private tree sum built on them (device only).
//
// One XYZZ addition issued by a warp keeps that SM sub-partition's integer multiplier busy for ~4 us (14 products)
// whether 1 or 32 lanes are active.  Where only a few additions are available at a time (reduction trees), lane r
// of a team computes product r of each of the 4 dependency levels of the formula and the results are exchanged
// with shuffles: ~1.5 us per addition, 8 additions per warp at a time.
#pragma once
#include "devmem.cuh"
#include "ec_compact.cuh"

namespace uz {

__device__ __forceinline__ fe sel4(uint32_t r, const fe& a0, const fe& a1, const fe& a2, const fe& a3) {
    fe o;
    const bool b0 = r & 1, b1 = r & 2;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const uint32_t lo = b0 ? a1.l[i] : a0.l[i];
        const uint32_t hi = b0 ? a3.l[i] : a2.l[i];
        o.l[i] = b1 ? hi : lo;
    }
    return o;
}
__device__ __forceinline__ fe shfl_fe(const fe& a, uint32_t src) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_sync(0xffffffffu, a.l[i], src);
    return r;
}

// P + Q.  Must be called by all 32 lanes; the 4 lanes of a team (lane / 4) pass identical arguments and all
// receive the result.  Lane r of the team computes product r of each of the 4 levels of add-2008-s.
static __device__ __noinline__ xyzz team4_add(xyzz P, xyzz Q) {
    const uint32_t lane = threadIdx.x & 31, r = lane & 3, base = lane & ~3u;
    fe m = fe_mul<FqP>(sel4(r, P.x, Q.x, P.y, Q.y), sel4(r, Q.zz, P.zz, Q.zzz, P.zzz));
    const fe u1 = shfl_fe(m, base), u2 = shfl_fe(m, base + 1), s1 = shfl_fe(m, base + 2), s2 = shfl_fe(m, base + 3);
    const fe pd = FQ_SUB(u2, u1), rd = FQ_SUB(s2, s1);
    m = fe_mul<FqP>(sel4(r, pd, rd, P.zz, P.zzz), sel4(r, pd, rd, Q.zz, Q.zzz));
    const fe pp = shfl_fe(m, base), rr = shfl_fe(m, base + 1), zz12 = shfl_fe(m, base + 2), zzz12 = shfl_fe(m, base + 3);
    m = fe_mul<FqP>(sel4(r, pd, u1, zz12, zz12), pp);
    const fe ppp = shfl_fe(m, base), qq = shfl_fe(m, base + 1), zz3 = shfl_fe(m, base + 2);
    xyzz o;
    o.x = FQ_SUB(FQ_SUB(rr, ppp), FQ_DBL(qq));
    m = fe_mul<FqP>(sel4(r, rd, s1, zzz12, zzz12), sel4(r, FQ_SUB(qq, o.x), ppp, ppp, ppp));
    const fe t1 = shfl_fe(m, base), t2 = shfl_fe(m, base + 1);
    o.zzz = shfl_fe(m, base + 2);
    o.y = FQ_SUB(t1, t2);
    o.zz = zz3;
    // special cases are uniform inside a team; nothing below shuffles
    if (xyzz_is_identity(Q)) return P;
    if (xyzz_is_identity(P)) return Q;
    if (fe_is_zero(pd)) {
        if (fe_is_zero(rd)) return xyzz_dbl_call(P);
        return xyzz_identity();
    }
    return o;
}

// 2 * P with the same calling convention (dbl-2008-s-1: U = 2Y, V = U^2, W = UV, S = XV, M = 3X^2)
static __device__ __noinline__ xyzz team4_dbl(xyzz P) {
    const uint32_t lane = threadIdx.x & 31, r = lane & 3, base = lane & ~3u;
    const fe u = FQ_DBL(P.y);
    fe m = fe_mul<FqP>(sel4(r, u, P.x, u, u), sel4(r, u, P.x, P.zzz, P.y));
    const fe v = shfl_fe(m, base), xx = shfl_fe(m, base + 1), uz = shfl_fe(m, base + 2), uy = shfl_fe(m, base + 3);
    const fe M = FQ_ADD(FQ_DBL(xx), xx);
    m = fe_mul<FqP>(sel4(r, P.x, M, v, v), sel4(r, v, M, P.zz, uz));
    const fe s = shfl_fe(m, base), mm = shfl_fe(m, base + 1);
    xyzz o;
    o.zz = shfl_fe(m, base + 2);
    o.zzz = shfl_fe(m, base + 3);
    o.x = FQ_SUB(mm, FQ_DBL(s));
    const fe d = FQ_SUB(s, o.x);
    m = fe_mul<FqP>(sel4(r, M, v, M, v), sel4(r, d, uy, d, uy));
    o.y = FQ_SUB(shfl_fe(m, base), shfl_fe(m, base + 1));
    if (xyzz_is_identity(P)) return P;
    return o;
}

// buf[0] <- buf[0] + ... + buf[n-1], n a power of two <= 32, buf in shared memory and private to the warp.
__device__ __forceinline__ void warp_team_tree(xyzz* buf, uint32_t n) {
    const uint32_t lane = threadIdx.x & 31, team = lane >> 2;
    while (n > 1) {
        const uint32_t half = n >> 1;
        for (uint32_t first = 0; first < half; first += 8) {
            const uint32_t p = first + team;
            const bool active = p < half;
            const xyzz a = active ? buf[2 * p] : xyzz_identity();
            const xyzz b = active ? buf[2 * p + 1] : xyzz_identity();
            const xyzz s = team4_add(a, b);
            __syncwarp();
            if (active && (lane & 3) == 0) buf[p] = s;
            __syncwarp();
        }
        n = half;
    }
}

}  // namespace uz
