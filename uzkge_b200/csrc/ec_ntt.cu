// ec_ntt.cu -- inverse transform "in the exponent": the Lagrange-basis SRS from the monomial one, WITHOUT the trapdoor.
//
//   L_i(tau) * G = (1 / n) * sum_j w^(-i j) * (tau^j * G),   i < n = 2^k
//
// i.e. the G1 analogue of FpPolynomial::ifft_with_domain applied to public_parameter_group_1 -- what produced the reference's
// bundled `lagrange-srs-*.bin` (uzkge/src/gen_params/mod.rs:42-65 loads them; the generator is not part of the repository) and
// what lets circuits larger than the bundled sizes use prover_with_lagrange.  Setup path, run once per SRS.
//
// Radix-2 decimation in frequency over XYZZ points, natural order in, bit-reversed out, one launch per stage; a butterfly is
// (u, v) -> (u + v, w^-t * (u - v)) with the twiddle applied by a 254-step double-and-add (the only "multiplication" a group
// offers): (n / 2) log2 n scalar multiplications of ~4 k Fq products each -- 7 ms of multiplier time at n = 2^14, ~3 s at 2^22.
// The last pass undoes the bit reversal, multiplies by 1 / n and normalises to affine.
#include <cuda_runtime.h>

#include "devmem.cuh"
#include "ec_compact.cuh"
#include "internal.h"

namespace uz {

// s * p, s a canonical (non-Montgomery) 254-bit scalar; out-of-line group operations keep the loop inside the instruction cache
static __device__ __noinline__ xyzz xyzz_mul_scalar(xyzz p, fe s) {
    xyzz acc = xyzz_identity();
    int top = 7;
    while (top >= 0 && s.l[top] == 0) top--;
    if (top < 0 || xyzz_is_identity(p)) return acc;
    int bit = 32 * top + 31 - __clz(s.l[top]);
#pragma unroll 1
    for (; bit >= 0; bit--) {
        acc = xyzz_dbl_call(acc);
        if ((s.l[bit >> 5] >> (bit & 31)) & 1) acc = xyzz_add_call(acc, p);
    }
    return acc;
}

__device__ __forceinline__ xyzz xyzz_neg(xyzz p) {
    if (!xyzz_is_identity(p)) p.y = fe_neg<FqP>(p.y);
    return p;
}

__global__ void __launch_bounds__(128) ec_load_kernel(const affine* __restrict__ in, uint32_t n, xyzz* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) st_xyzz(out + i, xyzz_from_affine(ld_affine(in + i)));
}

// one stage: blocks of `len` points, butterfly j < len / 2 uses tw[j * (n / len)] = w_n^(-j n / len) (Montgomery Fr)
__global__ void __launch_bounds__(128) ec_intt_stage_kernel(xyzz* __restrict__ x, uint32_t n, uint32_t len, const fe* __restrict__ tw) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n / 2) return;
    const uint32_t half = len >> 1;
    const uint32_t j = t & (half - 1);
    const uint32_t base = (t / half) * len + j;
    xyzz u = ld_xyzz(x + base);
    const xyzz v = ld_xyzz(x + base + half);
    xyzz d = xyzz_add_call(u, xyzz_neg(v));
    u = xyzz_add_call(u, v);
    if (j) d = xyzz_mul_scalar(d, fe_from_mont<FrP>(ld_fe(tw + (size_t)j * (n / len))));
    st_xyzz(x + base, u);
    st_xyzz(x + base + half, d);
}

// out[i] = n^-1 * x[bitrev(i)], affine
__global__ void __launch_bounds__(128) ec_intt_finish_kernel(const xyzz* __restrict__ x, uint32_t n, uint32_t log_n, fe n_inv, affine* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t r = log_n ? (__brev(i) >> (32 - log_n)) : 0;
    const xyzz p = xyzz_mul_scalar(ld_xyzz(x + r), n_inv);
    st_affine(out + i, xyzz_to_affine(p));
}

// d_points: n affine points (in), d_out: n affine points; d_work: n XYZZ points; d_tw: n / 2 Montgomery Fr twiddles w^-t
int ec_intt_run(const affine* d_points, uint32_t n, uint32_t log_n, const fe* d_tw, const fe& n_inv_canonical, xyzz* d_work, affine* d_out,
                cudaStream_t st) {
    const unsigned grid_n = (n + 127) / 128, grid_h = (n / 2 + 127) / 128;
    ec_load_kernel<<<grid_n, 128, 0, st>>>(d_points, n, d_work);
    for (uint32_t len = n; len >= 2; len >>= 1) ec_intt_stage_kernel<<<grid_h ? grid_h : 1, 128, 0, st>>>(d_work, n, len, d_tw);
    ec_intt_finish_kernel<<<grid_n, 128, 0, st>>>(d_work, n, log_n, n_inv_canonical, d_out);
    UZ_COUNT_LAUNCH(2 + log_n);
    return cudaGetLastError() == cudaSuccess ? UZKGE_OK : UZKGE_ERR_CUDA;
}

}  // namespace uz
