// ec.cuh -- BN254 G1 (y^2 = x^3 + 3, a = 0) group law for the MSM kernels.
//
// Replaces ark_ec::short_weierstrass::{Affine, Projective}<bn254::g1::Config> as used by
// G1Projective::msm (/root/reference/uzkge/src/poly_commit/kzg_poly_commitment.rs:287-290).
//
// Buckets are kept in extended Jacobian "XYZZ" coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; identity
// ZZ = 0): the mixed addition with an affine base costs 8M + 2S, the cheapest complete-enough formula for a
// bucket method, and needs no inversion.  The C ABI returns Jacobian (X, Y, Z) like arkworks' Projective;
// XYZZ -> Jacobian is (X*ZZ, Y*ZZZ, ZZ).  Affine identity on the ABI is x = y = 0.
#pragma once
#include "ff.cuh"

namespace uz {

struct affine {
    fe x, y;
};
struct xyzz {
    fe x, y, zz, zzz;
};
struct jacobian {
    fe x, y, z;
};

// Multiplication policy: the throughput-bound kernels inline every product; the latency-bound reduction kernels
// call ONE shared copy (ec_compact.cuh) so that their code stays inside the instruction caches.
struct FqInline {
    UZ_HD static fe mul(const fe& a, const fe& b) { return fe_mul<FqP>(a, b); }
    UZ_HD static fe sqr(const fe& a) { return fe_sqr<FqP>(a); }
};
#define FQ_MUL(a, b) M::mul(a, b)
#define FQ_SQR(a) M::sqr(a)
#define FQ_ADD(a, b) fe_add<FqP>(a, b)
#define FQ_SUB(a, b) fe_sub<FqP>(a, b)
#define FQ_DBL(a) fe_dbl<FqP>(a)

UZ_HD bool affine_is_identity(const affine& p) { return fe_is_zero(p.x) && fe_is_zero(p.y); }
UZ_HD bool xyzz_is_identity(const xyzz& p) { return fe_is_zero(p.zz); }

UZ_HD xyzz xyzz_identity() {
    xyzz r;
    r.x = fe_zero();
    r.y = fe_zero();
    r.zz = fe_zero();
    r.zzz = fe_zero();
    return r;
}
UZ_HD xyzz xyzz_from_affine(const affine& p) {
    xyzz r;
    if (affine_is_identity(p)) return xyzz_identity();
    r.x = p.x;
    r.y = p.y;
    r.zz = fe_one<FqP>();
    r.zzz = fe_one<FqP>();
    return r;
}

// 2 * (affine p), p != identity  (mdbl-2008-s-1)
template <class M = FqInline>
UZ_HD xyzz xyzz_dbl_affine(const affine& p) {
    xyzz r;
    fe u = FQ_DBL(p.y);
    fe v = FQ_SQR(u);
    fe w = FQ_MUL(u, v);
    fe s = FQ_MUL(p.x, v);
    fe m = FQ_SQR(p.x);
    m = FQ_ADD(FQ_DBL(m), m);
    r.x = FQ_SUB(FQ_SQR(m), FQ_DBL(s));
    r.y = FQ_SUB(FQ_MUL(m, FQ_SUB(s, r.x)), FQ_MUL(w, p.y));
    r.zz = v;
    r.zzz = w;
    return r;
}

// 2 * p  (dbl-2008-s-1); a point of order 2 does not exist on G1 (odd prime order), y == 0 only at identity
template <class M = FqInline>
UZ_HD xyzz xyzz_dbl(const xyzz& p) {
    if (xyzz_is_identity(p)) return p;
    xyzz r;
    fe u = FQ_DBL(p.y);
    fe v = FQ_SQR(u);
    fe w = FQ_MUL(u, v);
    fe s = FQ_MUL(p.x, v);
    fe m = FQ_SQR(p.x);
    m = FQ_ADD(FQ_DBL(m), m);
    r.x = FQ_SUB(FQ_SQR(m), FQ_DBL(s));
    r.y = FQ_SUB(FQ_MUL(m, FQ_SUB(s, r.x)), FQ_MUL(w, p.y));
    r.zz = FQ_MUL(v, p.zz);
    r.zzz = FQ_MUL(w, p.zzz);
    return r;
}

// acc += (x2, y2) with y2 already sign-adjusted by the caller  (madd-2008-s: 8M + 2S)
template <class M = FqInline>
UZ_HD void xyzz_madd(xyzz& acc, const affine& q) {
    if (affine_is_identity(q)) return;
    if (xyzz_is_identity(acc)) {
        acc = xyzz_from_affine(q);
        return;
    }
    fe u2 = FQ_MUL(q.x, acc.zz);
    fe s2 = FQ_MUL(q.y, acc.zzz);
    fe p = FQ_SUB(u2, acc.x);
    fe r = FQ_SUB(s2, acc.y);
    if (fe_is_zero(p)) {
        if (fe_is_zero(r))
            acc = xyzz_dbl_affine<M>(q);
        else
            acc = xyzz_identity();
        return;
    }
    fe pp = FQ_SQR(p);
    fe ppp = FQ_MUL(p, pp);
    fe qq = FQ_MUL(acc.x, pp);
    fe x3 = FQ_SUB(FQ_SUB(FQ_SQR(r), ppp), FQ_DBL(qq));
    fe y3 = FQ_SUB(FQ_MUL(r, FQ_SUB(qq, x3)), FQ_MUL(acc.y, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = FQ_MUL(acc.zz, pp);
    acc.zzz = FQ_MUL(acc.zzz, ppp);
}

// acc += q  (add-2008-s: 12M + 2S)
template <class M = FqInline>
UZ_HD void xyzz_add(xyzz& acc, const xyzz& q) {
    if (xyzz_is_identity(q)) return;
    if (xyzz_is_identity(acc)) {
        acc = q;
        return;
    }
    fe u1 = FQ_MUL(acc.x, q.zz);
    fe u2 = FQ_MUL(q.x, acc.zz);
    fe s1 = FQ_MUL(acc.y, q.zzz);
    fe s2 = FQ_MUL(q.y, acc.zzz);
    fe p = FQ_SUB(u2, u1);
    fe r = FQ_SUB(s2, s1);
    if (fe_is_zero(p)) {
        if (fe_is_zero(r))
            acc = xyzz_dbl<M>(acc);
        else
            acc = xyzz_identity();
        return;
    }
    fe pp = FQ_SQR(p);
    fe ppp = FQ_MUL(p, pp);
    fe qq = FQ_MUL(u1, pp);
    fe x3 = FQ_SUB(FQ_SUB(FQ_SQR(r), ppp), FQ_DBL(qq));
    fe y3 = FQ_SUB(FQ_MUL(r, FQ_SUB(qq, x3)), FQ_MUL(s1, ppp));
    acc.x = x3;
    acc.y = y3;
    acc.zz = FQ_MUL(FQ_MUL(acc.zz, q.zz), pp);
    acc.zzz = FQ_MUL(FQ_MUL(acc.zzz, q.zzz), ppp);
}

template <class M = FqInline>
UZ_HD jacobian xyzz_to_jacobian(const xyzz& p) {
    jacobian r;
    if (xyzz_is_identity(p)) {  // arkworks' Projective::zero() is (1, 1, 0)
        r.x = fe_one<FqP>();
        r.y = fe_one<FqP>();
        r.z = fe_zero();
        return r;
    }
    r.x = FQ_MUL(p.x, p.zz);
    r.y = FQ_MUL(p.y, p.zzz);
    r.z = p.zz;
    return r;
}

// XYZZ -> affine (one inversion); identity -> (0, 0)
template <class M = FqInline>
UZ_HD affine xyzz_to_affine(const xyzz& p) {
    affine r;
    if (xyzz_is_identity(p)) {
        r.x = fe_zero();
        r.y = fe_zero();
        return r;
    }
    // 1/ZZZ gives both: 1/ZZ = ZZ^2 / ZZZ^2 ... use two-step: i3 = 1/ZZZ, i2 = i3^2 * ZZ^... simpler: invert ZZ*ZZZ
    fe t = FQ_MUL(p.zz, p.zzz);
    fe ti = fe_inv<FqP>(t);
    fe izz = FQ_MUL(ti, p.zzz);
    fe izzz = FQ_MUL(ti, p.zz);
    r.x = FQ_MUL(p.x, izz);
    r.y = FQ_MUL(p.y, izzz);
    return r;
}

UZ_HD affine affine_neg(const affine& p) {
    affine r;
    r.x = p.x;
    r.y = fe_is_zero(p.y) ? p.y : fe_neg<FqP>(p.y);
    return r;
}

// k * p for a small unsigned k (bucket-segment offsets); double-and-add, MSB first
template <class M = FqInline>
UZ_HD xyzz xyzz_mul_u32(const xyzz& p, uint32_t k) {
    xyzz acc = xyzz_identity();
    for (int i = 31; i >= 0; i--) {
        acc = xyzz_dbl<M>(acc);
        if ((k >> i) & 1) xyzz_add<M>(acc, p);
    }
    return acc;
}

}  // namespace uz
