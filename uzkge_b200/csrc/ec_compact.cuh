// ec_compact.cuh -- out-of-line ("call") variants of the group law for latency-bound kernels (device only).
//
// A fully inlined XYZZ addition is ~2500 SASS instructions (40 KB): a kernel that strings a few dozen DEPENDENT
// additions together (bucket reduction trees, scans) runs out of the instruction caches and becomes fetch-bound
// (~12 us per addition measured).  Here the group operations exist once per translation unit and are called
// (arguments and results travel in registers, no stack frame).
#pragma once
#include "ec.cuh"

namespace uz {

static __device__ __noinline__ fe fq_mul_call(fe a, fe b) { return fe_mul<FqP>(a, b); }
struct FqCall {
    __device__ __forceinline__ static fe mul(const fe& a, const fe& b) { return fq_mul_call(a, b); }
    __device__ __forceinline__ static fe sqr(const fe& a) { return fq_mul_call(a, a); }
};

// One product at a time.  Measured on B200 (c = 17 reduction): 323 us with these; 369 us with variants that issue
// four independent products per call (register marshalling and padding products cost more than the interleaving
// gains); 520 us fully inlined.
static __device__ __noinline__ xyzz xyzz_add_call(xyzz a, xyzz b) {
    xyzz_add<FqCall>(a, b);
    return a;
}
static __device__ __noinline__ xyzz xyzz_dbl_call(xyzz a) { return xyzz_dbl<FqCall>(a); }

}  // namespace uz
