// Host-compiled view of the product's device headers (ff.cuh / ec.cuh plain-C twins).  TEST ONLY:
// lets the CPU suite pin the exact limb-level algorithms the kernels run, without a GPU.
#include <cstddef>
#include <cstring>
#include "../../uzkge_b200/csrc/ec.cuh"

using namespace uz;

extern "C" {
void uzhost_fq_mul(const fe* a, const fe* b, fe* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fe_mul<FqP>(a[i], b[i]); }
void uzhost_fr_mul(const fe* a, const fe* b, fe* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fe_mul<FrP>(a[i], b[i]); }
void uzhost_fq_sqr(const fe* a, fe* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fe_sqr<FqP>(a[i]); }
void uzhost_fr_sqr(const fe* a, fe* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fe_sqr<FrP>(a[i]); }
void uzhost_fq_add(const fe* a, const fe* b, fe* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fe_add<FqP>(a[i], b[i]); }
void uzhost_fq_sub(const fe* a, const fe* b, fe* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fe_sub<FqP>(a[i], b[i]); }
void uzhost_fr_add(const fe* a, const fe* b, fe* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fe_add<FrP>(a[i], b[i]); }
void uzhost_fr_sub(const fe* a, const fe* b, fe* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fe_sub<FrP>(a[i], b[i]); }
void uzhost_fr_inv(const fe* a, fe* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fe_inv<FrP>(a[i]); }
void uzhost_consts(fe* out /* fq one, fq r2, fr one, fr r2 */) {
    out[0] = fe_one<FqP>();
    out[1] = fe_to_mont<FqP>(fe_one<FqP>());
    out[2] = fe_one<FrP>();
    out[3] = fe_to_mont<FrP>(fe_one<FrP>());
}
// sum of n affine points with signs, via madd; output Jacobian
void uzhost_madd_chain(const affine* pts, const int* neg, size_t n, jacobian* out) {
    xyzz acc = xyzz_identity();
    for (size_t i = 0; i < n; i++) {
        affine p = neg[i] ? affine_neg(pts[i]) : pts[i];
        xyzz_madd(acc, p);
    }
    *out = xyzz_to_jacobian(acc);
}
// tree of full adds over singleton XYZZ points + doubling + small scalar mul
void uzhost_add_tree(const affine* pts, size_t n, uint32_t k, jacobian* out_sum, jacobian* out_ksum, affine* out_aff) {
    xyzz acc = xyzz_identity();
    for (size_t i = 0; i < n; i += 2) {
        xyzz a = xyzz_from_affine(pts[i]);
        if (i + 1 < n) { xyzz b = xyzz_from_affine(pts[i + 1]); xyzz_add(a, b); }
        xyzz_add(acc, a);
    }
    *out_sum = xyzz_to_jacobian(acc);
    *out_ksum = xyzz_to_jacobian(xyzz_mul_u32(acc, k));
    *out_aff = xyzz_to_affine(acc);
}
}

// ------------------------------------------------------------------ NTT plan emulator
// Executes the pass decomposition of ntt_plan.h with the same index helpers the kernels use (ntt_in_index,
// ntt_out_index, ntt_tw_exponent, ntt_bitrev) and the same order of fused scalings as ntt.cu, on the CPU.
#include <vector>
#include "../../uzkge_b200/csrc/ntt_plan.h"

static void emu_pass(const NttPass& p, const std::vector<fe>& in, std::vector<fe>& out, uint64_t n, uint64_t len_in,
                     const fe& omega, const fe* coset, bool inverse, bool zero_pad, bool pre_coset, bool post_coset,
                     bool has_scale, const fe& n_inv, uint32_t logm) {
    const uint32_t R = 1u << p.logR, C = 1u << p.logC;
    const uint32_t log_rtab = logm < 12 ? logm : 12;
    const fe w_rtab = fe_pow_u64<FrP>(omega, n >> log_rtab);
    std::vector<fe> tile((size_t)R * C);
    for (uint32_t b = 0; b < p.batch; b++)
        for (uint32_t o = 0; o < p.outer; o++)
            for (uint32_t t = 0; t < p.inner_tiles; t++) {
                for (uint32_t r = 0; r < R; r++)
                    for (uint32_t c = 0; c < C; c++) {
                        const uint64_t g = ntt_in_index(p, b, o, t, r, c);
                        fe x;
                        if (zero_pad && g >= len_in) x = fe_zero();
                        else {
                            x = in[g];
                            if (pre_coset) x = fe_mul<FrP>(x, fe_pow_u64<FrP>(*coset, g));
                        }
                        tile[(size_t)r * C + c] = x;
                    }
                for (int s = (int)p.logR - 1; s >= 0; s--) {
                    const uint32_t h = 1u << s, tw_shift = p.logR - 1 - s;
                    for (uint32_t pr = 0; pr < R / 2; pr++)
                        for (uint32_t c = 0; c < C; c++) {
                            const uint32_t j = pr & (h - 1), r = ((pr >> s) << (s + 1)) | j;
                            fe u = tile[(size_t)r * C + c], v = tile[(size_t)(r + h) * C + c];
                            fe d = fe_sub<FrP>(u, v);
                            if (s > 0) d = fe_mul<FrP>(d, fe_pow_u64<FrP>(w_rtab, (uint64_t)(j << tw_shift) * p.stage_stride));
                            tile[(size_t)r * C + c] = fe_add<FrP>(u, v);
                            tile[(size_t)(r + h) * C + c] = d;
                        }
                }
                for (uint32_t k = 0; k < R; k++)
                    for (uint32_t c = 0; c < C; c++) {
                        fe x = tile[(size_t)ntt_bitrev(k, p.logR) * C + c];
                        if (p.tw_mul) {
                            const uint64_t e = ntt_tw_exponent(p, t, k, c);
                            if (e) x = fe_mul<FrP>(x, fe_pow_u64<FrP>(omega, e));
                        }
                        uint64_t g = ntt_out_index(p, b, o, t, k, c);
                        if (p.last) {
                            if (inverse && g) g = n - g;
                            if (post_coset) x = fe_mul<FrP>(fe_mul<FrP>(x, fe_pow_u64<FrP>(*coset, g)), n_inv);
                            else if (has_scale) x = fe_mul<FrP>(x, n_inv);
                        }
                        out[g] = x;
                    }
            }
}

extern "C" {
// returns 0 on success, 1 if the size is unsupported.  data: n elements (first len_in are the input).
int uzhost_ntt(fe* data, uint64_t len_in, uint64_t n, int inverse, const fe* coset, uint32_t log_tile, uint32_t max_log_r,
               uint32_t two_pass_max, uint32_t* npass_out) {
    NttPlan pl;
    if (!ntt_make_plan(n, log_tile, max_log_r, two_pass_max, &pl)) return 1;
    if (npass_out) *npass_out = pl.npass;
    bool ok;
    const fe omega = ntt_root_of_unity(n, &ok);
    if (!ok) return 1;
    fe nf = fe_zero();
    nf.l[0] = (uint32_t)n;
    nf.l[1] = (uint32_t)(n >> 32);
    const fe n_inv = fe_inv<FrP>(fe_to_mont<FrP>(nf));
    std::vector<fe> src(data, data + n), scratch(n), out(n);
    bool consumed = false;
    if (pl.mixed) {
        const fe w3 = fe_pow_u64<FrP>(omega, pl.m), w3sq = fe_sqr<FrP>(w3);
        for (uint64_t i = 0; i < pl.m; i++) {
            fe x[3];
            for (int k = 0; k < 3; k++) {
                const uint64_t g = (uint64_t)k * pl.m + i;
                if (g < len_in) {
                    x[k] = src[g];
                    if (coset && !inverse) x[k] = fe_mul<FrP>(x[k], fe_pow_u64<FrP>(*coset, g));
                } else x[k] = fe_zero();
            }
            fe y0 = fe_add<FrP>(fe_add<FrP>(x[0], x[1]), x[2]);
            fe y1 = fe_add<FrP>(fe_add<FrP>(x[0], fe_mul<FrP>(x[1], w3)), fe_mul<FrP>(x[2], w3sq));
            fe y2 = fe_add<FrP>(fe_add<FrP>(x[0], fe_mul<FrP>(x[1], w3sq)), fe_mul<FrP>(x[2], w3));
            y1 = fe_mul<FrP>(y1, fe_pow_u64<FrP>(omega, i));
            y2 = fe_mul<FrP>(y2, fe_pow_u64<FrP>(omega, 2 * i));
            scratch[i] = y0;
            scratch[pl.m + i] = y1;
            scratch[2 * pl.m + i] = y2;
        }
        src = scratch;
        consumed = true;
    }
    for (uint32_t i = 0; i < pl.npass; i++) {
        const NttPass& p = pl.pass[i];
        std::vector<fe>& dst = p.last ? out : scratch;
        std::vector<fe> in_copy = src;
        emu_pass(p, in_copy, dst, n, len_in, omega, coset, inverse != 0, !consumed && len_in < n, !consumed && coset && !inverse,
                 p.last && coset && inverse, p.last && inverse && !coset, n_inv, pl.logm);
        src = dst;
        consumed = true;
    }
    memcpy(data, out.data(), n * sizeof(fe));
    return 0;
}
void uzhost_root_of_unity(uint64_t n, fe* out, int* ok) {
    bool b;
    *out = ntt_root_of_unity(n, &b);
    *ok = b;
}
}
