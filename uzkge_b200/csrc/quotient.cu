// quotient.cu -- the pointwise map of the TurboPlonK quotient polynomial on the coset k[1] * <w_m>, m = 6n.
//
// Replaces the rayon loop of t_poly (/root/reference/uzkge/src/plonk/helpers.rs:284-669) for the feature set without
// `shuffle` (terms 1-11; helpers.rs:643-652): inputs are the coset evaluations the prover already has on the device
// after the coset FFTs (helpers.rs:256-266) and the preprocessed coset evaluations of PlonkProverParams
// (plonk/indexer.rs:77-139); the output feeds the coset iFFT (helpers.rs:673-677).
//
// One thread per point: 32 loads of 32 B, ~95 Fr products -- 11 B per product against a machine balance of ~95 B per
// product (6.5 TB/s / 68 G mul/s): multiplier-bound, not HBM-bound.
#include <cuda_runtime.h>

#include "devmem.cuh"
#include "internal.h"

namespace uz {

struct QuotientDev {
    const fe* w[5];
    const fe* q[9];
    const fe* pi;
    const fe* z;
    const fe* s[5];
    const fe* coset_quotient;
    const fe* l1;
    const fe* qb;
    const fe* q_prk[4];
    fe k[5];
    fe alpha_pow[10];   // alpha^0 .. alpha^9
    fe beta, gamma, g, g_inv, g2p1;
    fe z_h_inv[16];
    uint64_t m;
    uint32_t factor;
    // the points this launch covers: p = start + step * idx, idx < count (the whole domain: 0, 1, m; one coset g_j <w_n> of it:
    // j, factor, n -- the omega-shifted point p + factor stays inside the coset and p % factor = j picks the coset's Z_H^-1)
    uint64_t start, step, count;
    fe* out;
    // `shuffle` feature set (terms 12-18, helpers.rs:416-640); unused otherwise
    const fe* w_sel[3];
    const fe* q_ecc;
    const fe* pk[12];    // q_shuffle_public_key_coset_evals: x_00..x_11, y_00..y_11, dxy_00..dxy_11
    const fe* gen[12];   // q_shuffle_generator_coset_evals
    fe edwards_a;
    fe alpha_pow_hi[7];  // alpha^10 .. alpha^16
};

#define FR_MUL(a, b) fe_mul<FrP>(a, b)
#define FR_ADD(a, b) fe_add<FrP>(a, b)
#define FR_SUB(a, b) fe_sub<FrP>(a, b)

__device__ __forceinline__ fe fr_pow5(const fe& x) {
    const fe x2 = FR_MUL(x, x);
    return FR_MUL(x, FR_MUL(x2, x2));
}

int g_quotient_min_blocks = 4;  // tuning knob (uzkge_cuda_configure "quotient_min_blocks"): 1 = 138 registers, 4 = 128 registers

// sum_c sel_c * v_c[p] over the four selector combinations 00, 01, 10, 11
__device__ __forceinline__ fe sel_combine(const fe* const* v, uint64_t p, const fe* sel) {
    fe r = FR_MUL(sel[0], ld_fe(v[0] + p));
    r = FR_ADD(r, FR_MUL(sel[1], ld_fe(v[1] + p)));
    r = FR_ADD(r, FR_MUL(sel[2], ld_fe(v[2] + p)));
    return FR_ADD(r, FR_MUL(sel[3], ld_fe(v[3] + p)));
}

// terms 12-18 (the `shuffle` feature): the elliptic-curve remark gates and the witness-selector constraints
__device__ __noinline__ fe quotient_shuffle_terms(const QuotientDev& a, uint64_t p, uint64_t pn, const fe* w) {
    const fe one = fe_one<FrP>();
    const fe ws0 = ld_fe(a.w_sel[0] + p), ws1 = ld_fe(a.w_sel[1] + p), ws2 = ld_fe(a.w_sel[2] + p), q_ecc = ld_fe(a.q_ecc + p);
    const fe n0 = FR_SUB(one, ws0), n1 = FR_SUB(one, ws1);
    fe sel[4];
    sel[0] = FR_SUB(FR_ADD(FR_MUL(n0, n1), q_ecc), one);
    sel[1] = FR_MUL(ws0, n1);
    sel[2] = FR_MUL(n0, ws1);
    sel[3] = FR_MUL(ws0, ws1);
    const fe ssum = FR_ADD(FR_ADD(sel[0], sel[1]), FR_ADD(sel[2], sel[3]));
    const fe w0n = ld_fe(a.w[0] + pn), w1n = ld_fe(a.w[1] + pn), w2n = ld_fe(a.w[2] + pn);
    const fe pkx = sel_combine(a.pk, p, sel), pky = sel_combine(a.pk + 4, p, sel), pkd = sel_combine(a.pk + 8, p, sel);
    const fe gx = sel_combine(a.gen, p, sel), gy = sel_combine(a.gen + 4, p, sel), gd = sel_combine(a.gen + 8, p, sel);
    const fe w01 = FR_MUL(w[0], w[1]), w23 = FR_MUL(w[2], w[3]);
    const fe ws2s = FR_MUL(ws2, ssum);
    // 12: ws2 w0' S - ws2 w0 PKY - w1 PKX + w0 w1 w0' PKD
    fe t12 = FR_SUB(FR_MUL(ws2s, w0n), FR_MUL(FR_MUL(ws2, w[0]), pky));
    t12 = FR_ADD(FR_SUB(t12, FR_MUL(w[1], pkx)), FR_MUL(FR_MUL(w01, w0n), pkd));
    // 13: ws2 w1' S + a w0 PKX - ws2 w1 PKY - w0 w1 w1' PKD
    fe t13 = FR_ADD(FR_MUL(ws2s, w1n), FR_MUL(FR_MUL(w[0], a.edwards_a), pkx));
    t13 = FR_SUB(FR_SUB(t13, FR_MUL(FR_MUL(ws2, w[1]), pky)), FR_MUL(FR_MUL(w01, w1n), pkd));
    // 14: ws2 w2' S - ws2 w2 GY - w3 GX + w2 w3 w2' GD
    fe t14 = FR_SUB(FR_MUL(ws2s, w2n), FR_MUL(FR_MUL(ws2, w[2]), gy));
    t14 = FR_ADD(FR_SUB(t14, FR_MUL(w[3], gx)), FR_MUL(FR_MUL(w23, w2n), gd));
    // 15: ws2 w4 S + a w2 GX - ws2 w3 GY - w2 w3 w4 GD
    fe t15 = FR_ADD(FR_MUL(ws2s, w[4]), FR_MUL(FR_MUL(w[2], a.edwards_a), gx));
    t15 = FR_SUB(FR_SUB(t15, FR_MUL(FR_MUL(ws2, w[3]), gy)), FR_MUL(FR_MUL(w23, w[4]), gd));
    // 16, 17: q_ecc ws (1 - ws) + (1 - q_ecc) ws;   18: q_ecc (1 + ws2) (1 - ws2)
    const fe nq = FR_SUB(one, q_ecc);
    const fe t16 = FR_ADD(FR_MUL(FR_MUL(q_ecc, ws0), n0), FR_MUL(nq, ws0));
    const fe t17 = FR_ADD(FR_MUL(FR_MUL(q_ecc, ws1), n1), FR_MUL(nq, ws1));
    const fe t18 = FR_MUL(FR_MUL(q_ecc, FR_ADD(one, ws2)), FR_SUB(one, ws2));
    fe r = FR_MUL(a.alpha_pow_hi[0], t12);
    r = FR_ADD(r, FR_MUL(a.alpha_pow_hi[1], t13));
    r = FR_ADD(r, FR_MUL(a.alpha_pow_hi[2], t14));
    r = FR_ADD(r, FR_MUL(a.alpha_pow_hi[3], t15));
    r = FR_ADD(r, FR_MUL(a.alpha_pow_hi[4], t16));
    r = FR_ADD(r, FR_MUL(a.alpha_pow_hi[5], t17));
    return FR_ADD(r, FR_MUL(a.alpha_pow_hi[6], t18));
}

template <int MIN_BLOCKS, bool SHUFFLE>
__global__ void __launch_bounds__(128, MIN_BLOCKS) plonk_quotient_kernel(const QuotientDev a) {
    const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.count) return;
    const uint64_t p = a.start + a.step * idx;
    const uint64_t pn = (p + a.factor) % a.m;  // the omega-shifted point (helpers.rs:308, 349-351)
    const fe one = fe_one<FrP>();
    fe w[5];
#pragma unroll
    for (int j = 0; j < 5; j++) w[j] = ld_fe(a.w[j] + p);
    const fe z = ld_fe(a.z + p), zn = ld_fe(a.z + pn);

    // term1: the gate function (constraint_system/turbo/mod.rs:193-222)
    const fe w01 = FR_MUL(w[0], w[1]), w23 = FR_MUL(w[2], w[3]);
    fe t = FR_MUL(ld_fe(a.q[0] + p), w[0]);
    t = FR_ADD(t, FR_MUL(ld_fe(a.q[1] + p), w[1]));
    t = FR_ADD(t, FR_MUL(ld_fe(a.q[2] + p), w[2]));
    t = FR_ADD(t, FR_MUL(ld_fe(a.q[3] + p), w[3]));
    t = FR_ADD(t, FR_MUL(ld_fe(a.q[4] + p), w01));
    t = FR_ADD(t, FR_MUL(ld_fe(a.q[5] + p), w23));
    t = FR_ADD(t, FR_ADD(ld_fe(a.q[6] + p), ld_fe(a.pi + p)));
    t = FR_ADD(t, FR_MUL(ld_fe(a.q[7] + p), FR_MUL(FR_MUL(w01, w23), w[4])));
    t = FR_SUB(t, FR_MUL(ld_fe(a.q[8] + p), w[4]));

    // term2, term3: the permutation argument (helpers.rs:298-317)
    const fe bcq = FR_MUL(a.beta, ld_fe(a.coset_quotient + p));
    fe t2 = FR_MUL(a.alpha_pow[1], z), t3 = FR_MUL(a.alpha_pow[1], zn);
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const fe wg = FR_ADD(w[j], a.gamma);
        t2 = FR_MUL(t2, FR_ADD(wg, FR_MUL(bcq, a.k[j])));
        t3 = FR_MUL(t3, FR_ADD(wg, FR_MUL(a.beta, ld_fe(a.s[j] + p))));
    }
    t = FR_ADD(t, t2);
    // term4 - term3 (helpers.rs:319-322, 646)
    const fe t4 = FR_MUL(FR_MUL(a.alpha_pow[2], ld_fe(a.l1 + p)), FR_SUB(z, one));
    t = FR_ADD(t, FR_SUB(t4, t3));
    // term5..7: boolean constraints (helpers.rs:324-346)
    // (every term carries the factor qb: rows without the boolean selector -- almost all -- skip the 9 products)
    const fe qb = ld_fe(a.qb + p);
    if (!fe_is_zero(qb)) {
#pragma unroll
        for (int j = 1; j <= 3; j++)
            t = FR_ADD(t, FR_MUL(FR_MUL(FR_MUL(a.alpha_pow[2 + j], qb), w[j]), FR_SUB(w[j], one)));
    }

    // term8..11: the Anemoi round constraints (helpers.rs:348-433)
    if (SHUFFLE) t = FR_ADD(t, quotient_shuffle_terms(a, p, pn, w));
    // all four carry the factor q_prk3, which is zero outside the Anemoi rows: those points skip the 34 products and 7 loads
    const fe prk3 = ld_fe(a.q_prk[2] + p);
    if (fe_is_zero(prk3)) {
        st_fe(a.out + p, FR_MUL(t, a.z_h_inv[p % a.factor]));
        return;
    }
    const fe w0n = ld_fe(a.w[0] + pn), w1n = ld_fe(a.w[1] + pn), w2n = ld_fe(a.w[2] + pn);
    const fe prk1 = ld_fe(a.q_prk[0] + p), prk2 = ld_fe(a.q_prk[1] + p), prk4 = ld_fe(a.q_prk[3] + p);
    const fe w30 = FR_ADD(w[0], w[3]), w21 = FR_ADD(w[1], w[2]);
    const fe w320 = FR_ADD(w[0], w30), w221 = FR_ADD(w[1], w21);
    {
        const fe tmp = FR_ADD(FR_ADD(w30, FR_MUL(a.g, w21)), prk3);
        const fe p5 = fr_pow5(FR_SUB(tmp, w2n));
        const fe e8 = FR_SUB(FR_ADD(p5, FR_MUL(a.g, FR_MUL(tmp, tmp))), FR_ADD(FR_ADD(w320, FR_MUL(a.g, w221)), prk1));
        const fe e10 = FR_SUB(FR_ADD(FR_ADD(p5, FR_MUL(a.g, FR_MUL(w2n, w2n))), a.g_inv), w0n);
        t = FR_SUB(t, FR_MUL(FR_MUL(a.alpha_pow[6], prk3), e8));
        t = FR_SUB(t, FR_MUL(FR_MUL(a.alpha_pow[8], prk3), e10));
    }
    {
        const fe tmp = FR_ADD(FR_ADD(FR_MUL(a.g, w30), FR_MUL(a.g2p1, w21)), prk4);
        const fe p5 = fr_pow5(FR_SUB(tmp, w[4]));
        const fe e9 = FR_SUB(FR_ADD(p5, FR_MUL(a.g, FR_MUL(tmp, tmp))),
                             FR_ADD(FR_ADD(FR_MUL(a.g, w320), FR_MUL(a.g2p1, w221)), prk2));
        const fe e11 = FR_SUB(FR_ADD(FR_ADD(p5, FR_MUL(a.g, FR_MUL(w[4], w[4]))), a.g_inv), w1n);
        t = FR_SUB(t, FR_MUL(FR_MUL(a.alpha_pow[7], prk3), e9));
        t = FR_SUB(t, FR_MUL(FR_MUL(a.alpha_pow[9], prk3), e11));
    }
    st_fe(a.out + p, FR_MUL(t, a.z_h_inv[p % a.factor]));
}

int plonk_quotient_run(const uzkge_quotient_args* args, const uzkge_quotient_shuffle_args* sh, void* d_out, cudaStream_t st) {
    return plonk_quotient_range_run(args, sh, 0, 1, args ? args->m : 0, d_out, st);
}

int plonk_quotient_range_run(const uzkge_quotient_args* args, const uzkge_quotient_shuffle_args* sh, uint64_t start, uint64_t step,
                             uint64_t count, void* d_out, cudaStream_t st) {
    if (!args || !d_out) return UZKGE_ERR_ARG;
    if (args->factor == 0 || args->factor > 16 || args->m == 0 || args->m % args->factor) return UZKGE_ERR_SIZE;
    if (count == 0) return UZKGE_OK;
    if (step == 0 || start >= args->m || (count - 1) > (args->m - 1 - start) / step) return UZKGE_ERR_SIZE;
    QuotientDev d;
    for (int j = 0; j < 5; j++) {
        d.w[j] = (const fe*)args->w[j];
        d.s[j] = (const fe*)args->s[j];
        memcpy(&d.k[j], args->k[j], sizeof(fe));
        if (!d.w[j] || !d.s[j]) return UZKGE_ERR_ARG;
    }
    for (int j = 0; j < 9; j++) {
        d.q[j] = (const fe*)args->q[j];
        if (!d.q[j]) return UZKGE_ERR_ARG;
    }
    for (int j = 0; j < 4; j++) {
        d.q_prk[j] = (const fe*)args->q_prk[j];
        if (!d.q_prk[j]) return UZKGE_ERR_ARG;
    }
    d.pi = (const fe*)args->pi;
    d.z = (const fe*)args->z;
    d.coset_quotient = (const fe*)args->coset_quotient;
    d.l1 = (const fe*)args->l1;
    d.qb = (const fe*)args->qb;
    if (!d.pi || !d.z || !d.coset_quotient || !d.l1 || !d.qb) return UZKGE_ERR_ARG;
    fe alpha;
    memcpy(&alpha, args->alpha, sizeof(fe));
    memcpy(&d.beta, args->beta, sizeof(fe));
    memcpy(&d.gamma, args->gamma, sizeof(fe));
    memcpy(&d.g, args->anemoi_generator, sizeof(fe));
    memcpy(&d.g_inv, args->anemoi_generator_inv, sizeof(fe));
    d.g2p1 = fe_add<FrP>(fe_sqr<FrP>(d.g), fe_one<FrP>());
    d.alpha_pow[0] = fe_one<FrP>();
    for (int i = 1; i < 10; i++) d.alpha_pow[i] = fe_mul<FrP>(d.alpha_pow[i - 1], alpha);
    if (sh) {
        fe ap = d.alpha_pow[9];
        for (int i = 0; i < 7; i++) {
            ap = fe_mul<FrP>(ap, alpha);
            d.alpha_pow_hi[i] = ap;
        }
        memcpy(&d.edwards_a, sh->edwards_a, sizeof(fe));
        for (int j = 0; j < 3; j++) d.w_sel[j] = (const fe*)sh->w_sel[j];
        d.q_ecc = (const fe*)sh->q_ecc;
        for (int j = 0; j < 12; j++) {
            d.pk[j] = (const fe*)sh->pk[j];
            d.gen[j] = (const fe*)sh->gen[j];
            if (!d.pk[j] || !d.gen[j]) return UZKGE_ERR_ARG;
        }
        if (!d.w_sel[0] || !d.w_sel[1] || !d.w_sel[2] || !d.q_ecc) return UZKGE_ERR_ARG;
    }
    for (uint32_t i = 0; i < 16; i++) d.z_h_inv[i] = fe_zero();
    for (uint32_t i = 0; i < args->factor; i++) memcpy(&d.z_h_inv[i], args->z_h_inv[i], sizeof(fe));
    d.m = args->m;
    d.factor = (uint32_t)args->factor;
    d.out = (fe*)d_out;
    d.start = start;
    d.step = step;
    d.count = count;
    const unsigned grid = (unsigned)((count + 127) / 128);
    if (sh)
        plonk_quotient_kernel<1, true><<<grid, 128, 0, st>>>(d);
    else if (g_quotient_min_blocks >= 4)
        plonk_quotient_kernel<4, false><<<grid, 128, 0, st>>>(d);
    else
        plonk_quotient_kernel<1, false><<<grid, 128, 0, st>>>(d);
    UZ_COUNT_LAUNCH(1);
    return cudaGetLastError() == cudaSuccess ? UZKGE_OK : UZKGE_ERR_CUDA;
}

}  // namespace uz
