// Host-compiled view of the product's device headers (ff.cuh / ec.cuh plain-C twins).  TEST ONLY:
// lets the CPU suite pin the exact limb-level algorithms the kernels run, without a GPU.
#include <cstddef>
#include <cstring>
#include "../../uzkge_b200/csrc/ec.cuh"

using namespace uz;

extern "C" {
void uzhost_fq_mul(const fe* a, const fe* b, fe* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fe_mul<FqP>(a[i], b[i]); }
void uzhost_fr_mul(const fe* a, const fe* b, fe* o, size_t n) { for (size_t i = 0; i < n; i++) o[i] = fe_mul<FrP>(a[i], b[i]); }
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
