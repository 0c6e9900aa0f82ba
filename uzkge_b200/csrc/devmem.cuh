// devmem.cuh -- 128-bit global / shared memory access helpers for 256-bit field elements (device only).
#pragma once
#include <cuda_runtime.h>

#include "ec.cuh"

namespace uz {

__device__ __forceinline__ fe fe_from_u4(const uint4& a, const uint4& b) {
    fe x;
    x.l[0] = a.x; x.l[1] = a.y; x.l[2] = a.z; x.l[3] = a.w;
    x.l[4] = b.x; x.l[5] = b.y; x.l[6] = b.z; x.l[7] = b.w;
    return x;
}
// read-only path (tables that never change while the kernel runs)
__device__ __forceinline__ fe ldg_fe(const fe* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    return fe_from_u4(__ldg(q), __ldg(q + 1));
}
__device__ __forceinline__ fe ld_fe(const fe* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    return fe_from_u4(q[0], q[1]);
}
__device__ __forceinline__ void st_fe(fe* p, const fe& x) {
    uint4* q = reinterpret_cast<uint4*>(p);
    q[0] = make_uint4(x.l[0], x.l[1], x.l[2], x.l[3]);
    q[1] = make_uint4(x.l[4], x.l[5], x.l[6], x.l[7]);
}

__device__ __forceinline__ xyzz ld_xyzz(const xyzz* p) {
    xyzz r;
    r.x = ld_fe(&p->x);
    r.y = ld_fe(&p->y);
    r.zz = ld_fe(&p->zz);
    r.zzz = ld_fe(&p->zzz);
    return r;
}
__device__ __forceinline__ void st_xyzz(xyzz* p, const xyzz& v) {
    st_fe(&p->x, v.x);
    st_fe(&p->y, v.y);
    st_fe(&p->zz, v.zz);
    st_fe(&p->zzz, v.zzz);
}
__device__ __forceinline__ affine ld_affine(const affine* p) {
    affine r;
    r.x = ld_fe(&p->x);
    r.y = ld_fe(&p->y);
    return r;
}
__device__ __forceinline__ void st_affine(affine* p, const affine& v) {
    st_fe(&p->x, v.x);
    st_fe(&p->y, v.y);
}

// ---- Ampere-style asynchronous 16-byte copies global -> shared (SASS: LDGSTS), bypassing L1
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

__device__ __forceinline__ fe shfl_xor_fe(const fe& a, int mask) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_xor_sync(0xffffffffu, a.l[i], mask);
    return r;
}
__device__ __forceinline__ xyzz shfl_xor_xyzz(const xyzz& a, int mask) {
    xyzz r;
    r.x = shfl_xor_fe(a.x, mask);
    r.y = shfl_xor_fe(a.y, mask);
    r.zz = shfl_xor_fe(a.zz, mask);
    r.zzz = shfl_xor_fe(a.zzz, mask);
    return r;
}
__device__ __forceinline__ xyzz shfl_xyzz(const xyzz& a, int src) {
    xyzz r;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        r.x.l[i] = __shfl_sync(0xffffffffu, a.x.l[i], src);
        r.y.l[i] = __shfl_sync(0xffffffffu, a.y.l[i], src);
        r.zz.l[i] = __shfl_sync(0xffffffffu, a.zz.l[i], src);
        r.zzz.l[i] = __shfl_sync(0xffffffffu, a.zzz.l[i], src);
    }
    return r;
}
__device__ __forceinline__ fe shfl_down_fe(const fe& a, int delta) {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = __shfl_down_sync(0xffffffffu, a.l[i], delta);
    return r;
}
__device__ __forceinline__ xyzz shfl_down_xyzz(const xyzz& a, int delta) {
    xyzz r;
    r.x = shfl_down_fe(a.x, delta);
    r.y = shfl_down_fe(a.y, delta);
    r.zz = shfl_down_fe(a.zz, delta);
    r.zzz = shfl_down_fe(a.zzz, delta);
    return r;
}

}  // namespace uz
