/*
 * oracle.c -- CPU restatement of the arithmetic behind uzkge's MSM / NTT hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Imported by tests/, __graft_entry__.smoke() and the
 * CPU-baseline legs of bench.py; never linked into, or called from, the CUDA product.
 *
 * The reference (/root/reference) is pure Rust and delegates this arithmetic to
 * un-vendored crates (ark-ff/ec/poly/bn254-zypher "0.4", /root/reference/Cargo.toml:28-38;
 * no Cargo.lock, no Rust toolchain in this image), so it cannot be compiled here.  This
 * file restates the PUBLISHED arkworks 0.4 algorithms that sit behind the reference's
 * call sites; parity is pinned through the reference's fixtures (tests/golden).
 *
 *   fp_mul / fp_add / fp_sub     ark_ff::Fp<MontBackend, 4>: 4 x u64 limbs, Montgomery R = 2^256
 *   oracle_msm_g1                ark_ec VariableBaseMSM::msm (msm_bigint_wnaf): signed-digit
 *                                Pippenger, window c = 3 if n < 32 else floor(log2 n)*69/100 + 2,
 *                                one task per window (rayon -> OpenMP), running-sum bucket
 *                                reduction, doubling ladder.  Call site:
 *                                /root/reference/uzkge/src/poly_commit/kzg_poly_commitment.rs:287-290
 *   oracle_ntt_fr                ark_poly Radix2EvaluationDomain::{fft,ifft}: roots table rebuilt per
 *                                call, forward = DIF butterflies + bit reversal, inverse = bit reversal
 *                                + DIT butterflies + 1/n; MixedRadixEvaluationDomain for 3 * 2^k:
 *                                digit-reversal permutation, radix-3 stage, radix-2 stages.  Coset
 *                                variants restate FpPolynomial::mul_var_assign (serial power loop).
 *                                Call sites: /root/reference/uzkge/src/poly_commit/field_polynomial.rs:470-477,583-607
 *
 * All field elements cross this API as 4 x u64 little-endian limbs in Montgomery form,
 * exactly like the CUDA C ABI (include/uzkge_cuda.h).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef unsigned __int128 u128;
typedef struct { uint64_t l[4]; } fe;

typedef struct {
    fe p;          /* modulus */
    uint64_t inv;  /* -p^-1 mod 2^64 */
    fe r;          /* R mod p   (Montgomery one) */
    fe r2;         /* R^2 mod p */
} field_t;

static field_t FQ, FR;
static int g_init = 0;

/* ------------------------------------------------------------------ field core */
static inline int fe_geq(const fe *a, const fe *b) {
    for (int i = 3; i >= 0; i--) {
        if (a->l[i] > b->l[i]) return 1;
        if (a->l[i] < b->l[i]) return 0;
    }
    return 1;
}
static inline int fe_is_zero(const fe *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int fe_eq(const fe *a, const fe *b) {
    return ((a->l[0] ^ b->l[0]) | (a->l[1] ^ b->l[1]) | (a->l[2] ^ b->l[2]) | (a->l[3] ^ b->l[3])) == 0;
}
static inline uint64_t raw_sub(fe *o, const fe *a, const fe *b) {
    u128 br = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a->l[i] - b->l[i] - (uint64_t)br;
        o->l[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
    return (uint64_t)br;
}
static inline uint64_t raw_add(fe *o, const fe *a, const fe *b) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
        c += (u128)a->l[i] + b->l[i];
        o->l[i] = (uint64_t)c;
        c >>= 64;
    }
    return (uint64_t)c;
}
static inline void fp_add(const field_t *F, fe *o, const fe *a, const fe *b) {
    fe t;
    raw_add(&t, a, b); /* p < 2^254: no carry out */
    if (fe_geq(&t, &F->p)) raw_sub(&t, &t, &F->p);
    *o = t;
}
static inline void fp_sub(const field_t *F, fe *o, const fe *a, const fe *b) {
    fe t;
    if (raw_sub(&t, a, b)) raw_add(&t, &t, &F->p);
    *o = t;
}
static inline void fp_neg(const field_t *F, fe *o, const fe *a) {
    if (fe_is_zero(a)) { *o = *a; return; }
    raw_sub(o, &F->p, a);
}
static inline void fp_dbl(const field_t *F, fe *o, const fe *a) { fp_add(F, o, a, a); }

/* CIOS Montgomery multiplication (ark-ff MontBackend::mul_assign, no-carry-free general form) */
static inline void fp_mul(const field_t *F, fe *o, const fe *a, const fe *b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)a->l[j] * b->l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * F->inv;
        c = (u128)m * F->p.l[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)m * F->p.l[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    fe r = {{t[0], t[1], t[2], t[3]}};
    if (t[4] || fe_geq(&r, &F->p)) raw_sub(&r, &r, &F->p);
    *o = r;
}
static inline void fp_sqr(const field_t *F, fe *o, const fe *a) { fp_mul(F, o, a, a); }

static void fp_pow(const field_t *F, fe *o, const fe *a, const fe *e) {
    fe acc = F->r, base = *a;
    for (int i = 0; i < 256; i++) {
        if ((e->l[i >> 6] >> (i & 63)) & 1) fp_mul(F, &acc, &acc, &base);
        fp_sqr(F, &base, &base);
    }
    *o = acc;
}
static void fp_inv(const field_t *F, fe *o, const fe *a) { /* Fermat */
    fe e = F->p, two = {{2, 0, 0, 0}};
    raw_sub(&e, &e, &two);
    fp_pow(F, o, a, &e);
}
static inline void fp_from_mont(const field_t *F, fe *o, const fe *a) {
    fe one = {{1, 0, 0, 0}};
    fp_mul(F, o, a, &one);
}
static inline void fp_to_mont(const field_t *F, fe *o, const fe *a) { fp_mul(F, o, a, &F->r2); }
static void fp_from_u64(const field_t *F, fe *o, uint64_t v) {
    fe t = {{v, 0, 0, 0}};
    fp_to_mont(F, o, &t);
}

static void field_setup(field_t *F, const uint64_t p[4]) {
    memcpy(F->p.l, p, 32);
    uint64_t inv = 1;
    for (int i = 0; i < 6; i++) inv *= 2 - p[0] * inv; /* Newton: p^-1 mod 2^64 */
    F->inv = (uint64_t)0 - inv;
    /* R mod p by 256 modular doublings of 1; R^2 by 256 more */
    fe x = {{1, 0, 0, 0}};
    for (int i = 0; i < 512; i++) {
        fe t;
        uint64_t c = raw_add(&t, &x, &x);
        if (c || fe_geq(&t, &F->p)) raw_sub(&t, &t, &F->p);
        x = t;
        if (i == 255) F->r = x;
    }
    F->r2 = x;
}

void oracle_init(void) {
    if (g_init) return;
    /* BN254 base and scalar moduli */
    static const uint64_t q[4] = {0x3c208c16d87cfd47ULL, 0x97816a916871ca8dULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    static const uint64_t r[4] = {0x43e1f593f0000001ULL, 0x2833e84879b97091ULL, 0xb85045b68181585dULL, 0x30644e72e131a029ULL};
    field_setup(&FQ, q);
    field_setup(&FR, r);
    g_init = 1;
}

/* vector helpers used by the tests to pin the CUDA field library */
void oracle_fr_mul(const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) {
    oracle_init();
    for (size_t i = 0; i < n; i++) fp_mul(&FR, (fe *)(o + 4 * i), (const fe *)(a + 4 * i), (const fe *)(b + 4 * i));
}
void oracle_fq_mul(const uint64_t *a, const uint64_t *b, uint64_t *o, size_t n) {
    oracle_init();
    for (size_t i = 0; i < n; i++) fp_mul(&FQ, (fe *)(o + 4 * i), (const fe *)(a + 4 * i), (const fe *)(b + 4 * i));
}
void oracle_fr_to_mont(const uint64_t *a, uint64_t *o, size_t n) {
    oracle_init();
    for (size_t i = 0; i < n; i++) fp_to_mont(&FR, (fe *)(o + 4 * i), (const fe *)(a + 4 * i));
}
void oracle_fr_from_mont(const uint64_t *a, uint64_t *o, size_t n) {
    oracle_init();
    for (size_t i = 0; i < n; i++) fp_from_mont(&FR, (fe *)(o + 4 * i), (const fe *)(a + 4 * i));
}
void oracle_fq_to_mont(const uint64_t *a, uint64_t *o, size_t n) {
    oracle_init();
    for (size_t i = 0; i < n; i++) fp_to_mont(&FQ, (fe *)(o + 4 * i), (const fe *)(a + 4 * i));
}
void oracle_fq_from_mont(const uint64_t *a, uint64_t *o, size_t n) {
    oracle_init();
    for (size_t i = 0; i < n; i++) fp_from_mont(&FQ, (fe *)(o + 4 * i), (const fe *)(a + 4 * i));
}

/* ------------------------------------------------------------------ G1 (Jacobian, a = 0) */
typedef struct { fe x, y, z; } jac;
typedef struct { fe x, y; } aff; /* identity: x = y = 0 (C ABI convention) */

static inline int aff_is_id(const aff *p) { return fe_is_zero(&p->x) && fe_is_zero(&p->y); }
static inline void jac_set_id(jac *p) { p->x = FQ.r; p->y = FQ.r; memset(&p->z, 0, sizeof(fe)); }
static inline int jac_is_id(const jac *p) { return fe_is_zero(&p->z); }

static void jac_double(jac *o, const jac *p) { /* dbl-2009-l */
    if (jac_is_id(p)) { *o = *p; return; }
    fe a, b, c, d, e, f, t;
    fp_sqr(&FQ, &a, &p->x);
    fp_sqr(&FQ, &b, &p->y);
    fp_sqr(&FQ, &c, &b);
    fp_add(&FQ, &t, &p->x, &b);
    fp_sqr(&FQ, &t, &t);
    fp_sub(&FQ, &t, &t, &a);
    fp_sub(&FQ, &t, &t, &c);
    fp_dbl(&FQ, &d, &t);
    fp_dbl(&FQ, &e, &a);
    fp_add(&FQ, &e, &e, &a);
    fp_sqr(&FQ, &f, &e);
    jac r;
    fp_mul(&FQ, &r.z, &p->y, &p->z);
    fp_dbl(&FQ, &r.z, &r.z);
    fp_dbl(&FQ, &t, &d);
    fp_sub(&FQ, &r.x, &f, &t);
    fp_sub(&FQ, &t, &d, &r.x);
    fp_mul(&FQ, &t, &e, &t);
    fp_dbl(&FQ, &c, &c);
    fp_dbl(&FQ, &c, &c);
    fp_dbl(&FQ, &c, &c);
    fp_sub(&FQ, &r.y, &t, &c);
    *o = r;
}

static void jac_add_mixed(jac *o, const jac *p, const aff *q) { /* madd-2007-bl */
    if (aff_is_id(q)) { *o = *p; return; }
    if (jac_is_id(p)) { o->x = q->x; o->y = q->y; o->z = FQ.r; return; }
    fe z1z1, u2, s2, h, hh, i, j, r, v, t;
    fp_sqr(&FQ, &z1z1, &p->z);
    fp_mul(&FQ, &u2, &q->x, &z1z1);
    fp_mul(&FQ, &s2, &q->y, &p->z);
    fp_mul(&FQ, &s2, &s2, &z1z1);
    if (fe_eq(&u2, &p->x)) {
        if (fe_eq(&s2, &p->y)) { jac_double(o, p); return; }
        jac_set_id(o);
        return;
    }
    fp_sub(&FQ, &h, &u2, &p->x);
    fp_sqr(&FQ, &hh, &h);
    fp_dbl(&FQ, &i, &hh);
    fp_dbl(&FQ, &i, &i);
    fp_mul(&FQ, &j, &h, &i);
    fp_sub(&FQ, &r, &s2, &p->y);
    fp_dbl(&FQ, &r, &r);
    fp_mul(&FQ, &v, &p->x, &i);
    jac out;
    fp_sqr(&FQ, &out.x, &r);
    fp_sub(&FQ, &out.x, &out.x, &j);
    fp_sub(&FQ, &out.x, &out.x, &v);
    fp_sub(&FQ, &out.x, &out.x, &v);
    fp_sub(&FQ, &t, &v, &out.x);
    fp_mul(&FQ, &t, &r, &t);
    fp_mul(&FQ, &j, &p->y, &j);
    fp_dbl(&FQ, &j, &j);
    fp_sub(&FQ, &out.y, &t, &j);
    fp_add(&FQ, &t, &p->z, &h);
    fp_sqr(&FQ, &t, &t);
    fp_sub(&FQ, &t, &t, &z1z1);
    fp_sub(&FQ, &out.z, &t, &hh);
    *o = out;
}

static void jac_add(jac *o, const jac *p, const jac *q) { /* add-2007-bl */
    if (jac_is_id(p)) { *o = *q; return; }
    if (jac_is_id(q)) { *o = *p; return; }
    fe z1z1, z2z2, u1, u2, s1, s2, h, i, j, r, v, t;
    fp_sqr(&FQ, &z1z1, &p->z);
    fp_sqr(&FQ, &z2z2, &q->z);
    fp_mul(&FQ, &u1, &p->x, &z2z2);
    fp_mul(&FQ, &u2, &q->x, &z1z1);
    fp_mul(&FQ, &s1, &p->y, &q->z);
    fp_mul(&FQ, &s1, &s1, &z2z2);
    fp_mul(&FQ, &s2, &q->y, &p->z);
    fp_mul(&FQ, &s2, &s2, &z1z1);
    if (fe_eq(&u1, &u2)) {
        if (fe_eq(&s1, &s2)) { jac_double(o, p); return; }
        jac_set_id(o);
        return;
    }
    fp_sub(&FQ, &h, &u2, &u1);
    fp_dbl(&FQ, &i, &h);
    fp_sqr(&FQ, &i, &i);
    fp_mul(&FQ, &j, &h, &i);
    fp_sub(&FQ, &r, &s2, &s1);
    fp_dbl(&FQ, &r, &r);
    fp_mul(&FQ, &v, &u1, &i);
    jac out;
    fp_sqr(&FQ, &out.x, &r);
    fp_sub(&FQ, &out.x, &out.x, &j);
    fp_sub(&FQ, &out.x, &out.x, &v);
    fp_sub(&FQ, &out.x, &out.x, &v);
    fp_sub(&FQ, &t, &v, &out.x);
    fp_mul(&FQ, &t, &r, &t);
    fp_mul(&FQ, &s1, &s1, &j);
    fp_dbl(&FQ, &s1, &s1);
    fp_sub(&FQ, &out.y, &t, &s1);
    fp_add(&FQ, &t, &p->z, &q->z);
    fp_sqr(&FQ, &t, &t);
    fp_sub(&FQ, &t, &t, &z1z1);
    fp_sub(&FQ, &t, &t, &z2z2);
    fp_mul(&FQ, &out.z, &t, &h);
    *o = out;
}

static void jac_to_affine(aff *o, const jac *p) {
    if (jac_is_id(p)) { memset(o, 0, sizeof(aff)); return; }
    fe zi, zi2, zi3;
    fp_inv(&FQ, &zi, &p->z);
    fp_sqr(&FQ, &zi2, &zi);
    fp_mul(&FQ, &zi3, &zi2, &zi);
    fp_mul(&FQ, &o->x, &p->x, &zi2);
    fp_mul(&FQ, &o->y, &p->y, &zi3);
}

/* signed radix-2^c digits of a canonical 256-bit scalar (ark_ec::scalar_mul::variable_base::make_digits) */
static void make_digits(const fe *s, int c, int ndig, int64_t *out) {
    uint64_t carry = 0;
    const uint64_t radix = 1ULL << c, mask = radix - 1;
    for (int i = 0; i < ndig; i++) {
        int bit = i * c, w = bit >> 6, sh = bit & 63;
        uint64_t v = 0;
        if (w < 4) {
            v = s->l[w] >> sh;
            if (sh + c > 64 && w + 1 < 4) v |= s->l[w + 1] << (64 - sh);
        }
        uint64_t coef = carry + (v & mask);
        carry = (coef + radix / 2) >> c;
        int64_t d = (int64_t)coef - (int64_t)(carry << c);
        if (i == ndig - 1) d += (int64_t)(carry << c); /* top digit absorbs the carry */
        out[i] = d;
    }
}

static int msm_window(size_t n) {
    if (n < 32) return 3;
    int lg = 0;
    while (((size_t)1 << (lg + 1)) <= n) lg++;
    return lg * 69 / 100 + 2;
}

/* out = sum_i scalars[i] * bases[i];  bases: n x (x[4], y[4]) Montgomery affine, identity = zeros;
 * scalars: n x 4 Montgomery Fr;  out: X[4] Y[4] Z[4] Montgomery Jacobian, Z = 0 for the identity. */
int oracle_msm_g1(const uint64_t *bases, const uint64_t *scalars, size_t n, uint64_t *out_jac) {
    oracle_init();
    jac total;
    jac_set_id(&total);
    if (n == 0) { memcpy(out_jac, &total, 96); return 0; }
    const aff *B = (const aff *)bases;
    const int c = msm_window(n);
    const int nwin = (254 + c - 1) / c;
    int64_t *digits = (int64_t *)malloc(sizeof(int64_t) * n * (size_t)nwin);
    if (!digits) return 1;
#pragma omp parallel for schedule(static)
    for (size_t i = 0; i < n; i++) {
        fe s;
        fp_from_mont(&FR, &s, (const fe *)(scalars + 4 * i)); /* into_bigint */
        make_digits(&s, c, nwin, digits + i * (size_t)nwin);
    }
    jac *wsum = (jac *)malloc(sizeof(jac) * (size_t)nwin);
    int err = 0;
#pragma omp parallel for schedule(dynamic, 1)
    for (int w = 0; w < nwin; w++) {
        size_t nb = (size_t)1 << c; /* arkworks allocates 1 << c buckets per window */
        jac *bk = (jac *)malloc(sizeof(jac) * nb);
        if (!bk) { err = 1; jac_set_id(&wsum[w]); continue; }
        for (size_t k = 0; k < nb; k++) jac_set_id(&bk[k]);
        for (size_t i = 0; i < n; i++) {
            int64_t d = digits[i * (size_t)nwin + w];
            if (d > 0) {
                jac_add_mixed(&bk[d - 1], &bk[d - 1], &B[i]);
            } else if (d < 0) {
                aff m = B[i];
                if (!aff_is_id(&m)) fp_neg(&FQ, &m.y, &m.y);
                jac_add_mixed(&bk[-d - 1], &bk[-d - 1], &m);
            }
        }
        jac run, res;
        jac_set_id(&run);
        jac_set_id(&res);
        for (size_t k = nb; k-- > 0;) {
            jac_add(&run, &run, &bk[k]);
            jac_add(&res, &res, &run);
        }
        wsum[w] = res;
        free(bk);
    }
    for (int w = nwin - 1; w >= 1; w--) {
        jac_add(&total, &total, &wsum[w]);
        for (int k = 0; k < c; k++) jac_double(&total, &total);
    }
    jac_add(&total, &total, &wsum[0]);
    memcpy(out_jac, &total, 96);
    free(wsum);
    free(digits);
    return err;
}

/* scalar multiplication (double-and-add), used for naive MSM checks and the trapdoor known answer */
void oracle_g1_mul(const uint64_t *base_aff, const uint64_t *scalar_mont, uint64_t *out_jac) {
    oracle_init();
    fe s;
    fp_from_mont(&FR, &s, (const fe *)scalar_mont);
    jac acc;
    jac_set_id(&acc);
    const aff *b = (const aff *)base_aff;
    for (int i = 255; i >= 0; i--) {
        jac_double(&acc, &acc);
        if ((s.l[i >> 6] >> (i & 63)) & 1) jac_add_mixed(&acc, &acc, b);
    }
    memcpy(out_jac, &acc, 96);
}
void oracle_g1_add_jac(const uint64_t *a, const uint64_t *b, uint64_t *out_jac) {
    oracle_init();
    jac r;
    jac_add(&r, (const jac *)a, (const jac *)b);
    memcpy(out_jac, &r, 96);
}
void oracle_g1_to_affine(const uint64_t *in_jac, uint64_t *out_aff) {
    oracle_init();
    jac_to_affine((aff *)out_aff, (const jac *)in_jac);
}
int oracle_g1_on_curve(const uint64_t *p_aff) {
    oracle_init();
    const aff *p = (const aff *)p_aff;
    if (aff_is_id(p)) return 1;
    fe y2, x3, b;
    fp_sqr(&FQ, &y2, &p->y);
    fp_sqr(&FQ, &x3, &p->x);
    fp_mul(&FQ, &x3, &x3, &p->x);
    fp_from_u64(&FQ, &b, 3);
    fp_add(&FQ, &x3, &x3, &b);
    return fe_eq(&y2, &x3);
}

/* ---- synthetic bases.  splitmix64 stream -> x candidates, try-and-increment, y = rhs^((p+1)/4) */
static inline uint64_t splitmix64(uint64_t *s) {
    uint64_t z = (*s += 0x9e3779b97f4a7c15ULL);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
    return z ^ (z >> 31);
}
void oracle_g1_random_points(uint64_t seed, size_t n, uint64_t *out_aff) {
    oracle_init();
    fe e = FQ.p, one = {{1, 0, 0, 0}}, three;
    raw_add(&e, &e, &one); /* (p + 1) / 4 */
    for (int i = 0; i < 4; i++) e.l[i] = (e.l[i] >> 2) | (i < 3 ? e.l[i + 1] << 62 : 0);
    fp_from_u64(&FQ, &three, 3);
#pragma omp parallel for schedule(dynamic, 64)
    for (size_t i = 0; i < n; i++) {
        uint64_t st = seed ^ (0xD1B54A32D192ED03ULL * (i + 1));
        fe x;
        for (int k = 0; k < 4; k++) x.l[k] = splitmix64(&st);
        x.l[3] &= 0x0fffffffffffffffULL; /* < 2^252 < p : already a valid Montgomery residue */
        uint64_t sign = splitmix64(&st) & 1;
        for (;;) {
            fe rhs, y, y2;
            fp_sqr(&FQ, &rhs, &x);
            fp_mul(&FQ, &rhs, &rhs, &x);
            fp_add(&FQ, &rhs, &rhs, &three);
            fp_pow(&FQ, &y, &rhs, &e);
            fp_sqr(&FQ, &y2, &y);
            if (fe_eq(&y2, &rhs)) {
                if (sign) fp_neg(&FQ, &y, &y);
                aff *o = (aff *)(out_aff + 8 * i);
                o->x = x;
                o->y = y;
                break;
            }
            fp_add(&FQ, &x, &x, &FQ.r); /* x += 1 */
        }
    }
}

/* arithmetic-progression bases P_i = P_0 + i*Q (affine chain with one inversion per step batch);
 * gives the O(n) trapdoor answer  sum s_i P_i = (sum s_i) P_0 + (sum i s_i) Q  for huge MSMs. */
void oracle_g1_progression(const uint64_t *p0_aff, const uint64_t *q_aff, size_t n, uint64_t *out_aff) {
    oracle_init();
    if (n == 0) return;
    aff *o = (aff *)out_aff;
    const aff *Q = (const aff *)q_aff;
    o[0] = *(const aff *)p0_aff;
    /* chunked: Jacobian running point, batch-invert Z per chunk */
    enum { CH = 1024 };
    jac cur;
    cur.x = o[0].x; cur.y = o[0].y; cur.z = FQ.r;
    jac *buf = (jac *)malloc(sizeof(jac) * CH);
    fe *pre = (fe *)malloc(sizeof(fe) * CH);
    size_t i = 1;
    while (i < n) {
        size_t m = n - i < CH ? n - i : CH;
        for (size_t k = 0; k < m; k++) {
            jac_add_mixed(&cur, &cur, Q);
            buf[k] = cur;
        }
        fe acc = FQ.r;
        for (size_t k = 0; k < m; k++) { pre[k] = acc; fp_mul(&FQ, &acc, &acc, &buf[k].z); }
        fe inv;
        fp_inv(&FQ, &inv, &acc);
        for (size_t k = m; k-- > 0;) {
            fe zi, zi2, zi3;
            fp_mul(&FQ, &zi, &inv, &pre[k]);
            fp_mul(&FQ, &inv, &inv, &buf[k].z);
            fp_sqr(&FQ, &zi2, &zi);
            fp_mul(&FQ, &zi3, &zi2, &zi);
            fp_mul(&FQ, &o[i + k].x, &buf[k].x, &zi2);
            fp_mul(&FQ, &o[i + k].y, &buf[k].y, &zi3);
        }
        i += m;
    }
    free(buf);
    free(pre);
}

/* ------------------------------------------------------------------ NTT */
static void fr_root_of_unity(fe *o, size_t n, int a3, int b2) {
    /* arkworks FftField::get_root_of_unity with a small subgroup: generator 5, 2-adicity 28, 3^2 | r-1.
     * LARGE = 5^((r-1)/(2^28 * 9));  w_n = LARGE^(3^(2-a) * 2^(28-b))  for n = 3^a 2^b  (SURVEY 8c-S4). */
    (void)n;
    fe g, e = FR.p, one = {{1, 0, 0, 0}};
    fp_from_u64(&FR, &g, 5);
    raw_sub(&e, &e, &one);
    /* e = (r - 1) / 2^28 / 9 */
    for (int k = 0; k < 28; k++)
        for (int i = 0; i < 4; i++) e.l[i] = (e.l[i] >> 1) | (i < 3 ? e.l[i + 1] << 63 : 0);
    u128 rem = 0;
    for (int i = 3; i >= 0; i--) {
        u128 cur = (rem << 64) | e.l[i];
        e.l[i] = (uint64_t)(cur / 9);
        rem = cur % 9;
    }
    fe large;
    fp_pow(&FR, &large, &g, &e);
    for (int k = 0; k < 2 - a3; k++) { fe t; fp_sqr(&FR, &t, &large); fp_mul(&FR, &large, &t, &large); }
    for (int k = 0; k < 28 - b2; k++) fp_sqr(&FR, &large, &large);
    *o = large;
}

static int factor_domain(size_t n, int *a3, int *b2) {
    int a = 0, b = 0;
    if (n == 0) return 1;
    while (n % 3 == 0) { n /= 3; a++; }
    while (n % 2 == 0) { n /= 2; b++; }
    if (n != 1 || a > 1 || b > 28) return 1; /* uzkge only uses 2^k and 3 * 2^k (field_polynomial.rs:561-567) */
    *a3 = a; *b2 = b;
    return 0;
}

void oracle_fr_root_of_unity(size_t n, uint64_t *out_mont) {
    oracle_init();
    int a, b;
    if (factor_domain(n, &a, &b)) { memset(out_mont, 0, 32); return; }
    fr_root_of_unity((fe *)out_mont, n, a, b);
}

static size_t bitrev(size_t x, int bits) {
    size_t r = 0;
    for (int i = 0; i < bits; i++) { r = (r << 1) | (x & 1); x >>= 1; }
    return r;
}

/* roots[i] = w^i, i < cnt  (ark_poly roots_of_unity / compute_powers: rebuilt on every call) */
static void build_roots(fe *roots, size_t cnt, const fe *w) {
    if (cnt == 0) return;
    enum { BLK = 4096 };
    size_t nblk = (cnt + BLK - 1) / BLK;
    fe *starts = (fe *)malloc(sizeof(fe) * nblk);
    fe wb = FR.r, t = *w;
    size_t e = BLK; /* wb = w^BLK */
    while (e) { if (e & 1) fp_mul(&FR, &wb, &wb, &t); fp_sqr(&FR, &t, &t); e >>= 1; }
    starts[0] = FR.r;
    for (size_t b = 1; b < nblk; b++) fp_mul(&FR, &starts[b], &starts[b - 1], &wb);
#pragma omp parallel for schedule(static)
    for (size_t b = 0; b < nblk; b++) {
        fe cur = starts[b];
        size_t hi = (b + 1) * BLK < cnt ? (b + 1) * BLK : cnt;
        for (size_t i = b * BLK; i < hi; i++) { roots[i] = cur; fp_mul(&FR, &cur, &cur, w); }
    }
    free(starts);
}

static void derange(fe *x, size_t n, int logn) {
    for (size_t i = 0; i < n; i++) {
        size_t j = bitrev(i, logn);
        if (i < j) { fe t = x[i]; x[i] = x[j]; x[j] = t; }
    }
}

/* in-order -> bit-reversed: DIF (Gentleman-Sande) layers, ark_poly radix2 io_helper */
static void radix2_dif(fe *x, size_t n, const fe *roots /* n/2 powers of w_n */) {
    for (size_t gap = n / 2; gap >= 1; gap >>= 1) {
        size_t step = n / (2 * gap);
#pragma omp parallel for schedule(static) if (n >= 4096)
        for (size_t k = 0; k < n / 2; k++) {
            size_t blk = k / gap, j = k % gap;
            fe *lo = &x[blk * 2 * gap + j], *hi = lo + gap;
            fe d;
            fp_sub(&FR, &d, lo, hi);
            fp_add(&FR, lo, lo, hi);
            fp_mul(&FR, hi, &d, &roots[j * step]);
        }
    }
}
/* bit-reversed -> in-order: DIT (Cooley-Tukey) layers, ark_poly radix2 oi_helper */
static void radix2_dit(fe *x, size_t n, const fe *roots) {
    for (size_t gap = 1; gap < n; gap <<= 1) {
        size_t step = n / (2 * gap);
#pragma omp parallel for schedule(static) if (n >= 4096)
        for (size_t k = 0; k < n / 2; k++) {
            size_t blk = k / gap, j = k % gap;
            fe *lo = &x[blk * 2 * gap + j], *hi = lo + gap;
            fe t;
            fp_mul(&FR, &t, hi, &roots[j * step]);
            fp_sub(&FR, hi, lo, &t);
            fp_add(&FR, lo, lo, &t);
        }
    }
}

/* size 3 * 2^k, Cooley-Tukey with the radix-3 split outermost:
 *   X[k1 + 3 k2] = sum_{n2 < M} w_N^(n2 k1) [ sum_{n1 < 3} x[M n1 + n2] w_3^(n1 k1) ] w_M^(n2 k2),  M = 2^k */
static void mixed_fft(fe *x, size_t n, const fe *w, int logm) {
    size_t m = n / 3;
    fe *roots = (fe *)malloc(sizeof(fe) * n);
    build_roots(roots, n, w);
    const fe *w3 = &roots[m], *w3sq = &roots[2 * m];
    fe *tmp = (fe *)malloc(sizeof(fe) * n);
#pragma omp parallel for schedule(static) if (n >= 4096)
    for (size_t j = 0; j < m; j++) {
        const fe *a = &x[j], *b = &x[m + j], *c = &x[2 * m + j];
        fe s, t1, t2, u1, u2, y1, y2;
        fp_add(&FR, &s, a, b);
        fp_add(&FR, &s, &s, c);
        fp_mul(&FR, &t1, b, w3);
        fp_mul(&FR, &t2, c, w3sq);
        fp_add(&FR, &y1, a, &t1);
        fp_add(&FR, &y1, &y1, &t2);
        fp_mul(&FR, &u1, b, w3sq);
        fp_mul(&FR, &u2, c, w3); /* w3^4 = w3 */
        fp_add(&FR, &y2, a, &u1);
        fp_add(&FR, &y2, &y2, &u2);
        tmp[j] = s;
        fp_mul(&FR, &tmp[m + j], &y1, &roots[j]);
        fp_mul(&FR, &tmp[2 * m + j], &y2, &roots[(2 * j) % n]);
    }
    /* three radix-2 transforms of size m with root w^3 */
    fe *r2 = (fe *)malloc(sizeof(fe) * (m / 2 + 1));
    for (size_t i = 0; i < m / 2; i++) r2[i] = roots[3 * i];
    for (int k1 = 0; k1 < 3; k1++) {
        fe *row = tmp + (size_t)k1 * m;
        if (m > 1) {
            radix2_dif(row, m, r2);
            derange(row, m, logm);
        }
        for (size_t k2 = 0; k2 < m; k2++) x[k1 + 3 * k2] = row[k2];
    }
    free(r2);
    free(tmp);
    free(roots);
}

/* Same contract as uzkge_cuda_ntt_fr (include/uzkge_cuda.h): data holds `n` elements, the first `len_in`
 * are the input (rest treated as zero), output is n elements in natural order.
 *   forward: optional pre-scale c_j *= g^j  (coset_fft_with_domain), then DFT
 *   inverse: inverse DFT incl. 1/n, then optional post-scale c_j *= g^j  where the caller passes g = k^-1 */
int oracle_ntt_fr(uint64_t *data, size_t len_in, size_t n, int inverse, const uint64_t *coset) {
    oracle_init();
    int a3, b2;
    if (factor_domain(n, &a3, &b2) || len_in > n) return 1;
    fe *x = (fe *)data;
    for (size_t i = len_in; i < n; i++) memset(&x[i], 0, sizeof(fe));
    fe w;
    fr_root_of_unity(&w, n, a3, b2);
    if (!inverse && coset) { /* serial power loop, FpPolynomial::mul_var_assign */
        fe p = FR.r;
        const fe *g = (const fe *)coset;
        for (size_t i = 0; i < len_in; i++) { fp_mul(&FR, &x[i], &x[i], &p); fp_mul(&FR, &p, &p, g); }
    }
    if (inverse) fp_inv(&FR, &w, &w);
    if (n > 1) {
        if (a3 == 0) {
            fe *roots = (fe *)malloc(sizeof(fe) * (n / 2));
            build_roots(roots, n / 2, &w);
            if (!inverse) {
                radix2_dif(x, n, roots);
                derange(x, n, b2);
            } else {
                derange(x, n, b2);
                radix2_dit(x, n, roots);
            }
            free(roots);
        } else {
            mixed_fft(x, n, &w, b2);
        }
    }
    if (inverse) {
        fe ninv;
        fp_from_u64(&FR, &ninv, (uint64_t)n);
        fp_inv(&FR, &ninv, &ninv);
#pragma omp parallel for schedule(static) if (n >= 4096)
        for (size_t i = 0; i < n; i++) fp_mul(&FR, &x[i], &x[i], &ninv);
        if (coset) {
            fe p = FR.r;
            const fe *g = (const fe *)coset;
            for (size_t i = 0; i < n; i++) { fp_mul(&FR, &x[i], &x[i], &p); fp_mul(&FR, &p, &p, g); }
        }
    }
    return 0;
}

/* Horner evaluation at a canonical-Montgomery point: spot check for big transforms (check_fft contract) */
void oracle_fr_eval(const uint64_t *coefs, size_t n, const uint64_t *x_mont, uint64_t *out_mont) {
    oracle_init();
    fe acc;
    memset(&acc, 0, sizeof(fe));
    const fe *c = (const fe *)coefs;
    for (size_t i = n; i-- > 0;) {
        fp_mul(&FR, &acc, &acc, (const fe *)x_mont);
        fp_add(&FR, &acc, &acc, &c[i]);
    }
    memcpy(out_mont, &acc, 32);
}
void oracle_fr_pow_u64(const uint64_t *a_mont, uint64_t e, uint64_t *out_mont) {
    oracle_init();
    fe ee = {{e, 0, 0, 0}};
    fp_pow(&FR, (fe *)out_mont, (const fe *)a_mont, &ee);
}
void oracle_fr_inv(const uint64_t *a_mont, uint64_t *out_mont) {
    oracle_init();
    fp_inv(&FR, (fe *)out_mont, (const fe *)a_mont);
}
/* sums for the progression trapdoor: s0 = sum s_i, s1 = sum i * s_i (Montgomery in / out) */
void oracle_fr_weighted_sums(const uint64_t *s, size_t n, uint64_t *s0_out, uint64_t *s1_out) {
    oracle_init();
    fe s0, s1;
    memset(&s0, 0, sizeof(fe));
    memset(&s1, 0, sizeof(fe));
    for (size_t i = 0; i < n; i++) {
        fe idx, t;
        fp_from_u64(&FR, &idx, (uint64_t)i);
        fp_mul(&FR, &t, &idx, (const fe *)(s + 4 * i));
        fp_add(&FR, &s1, &s1, &t);
        fp_add(&FR, &s0, &s0, (const fe *)(s + 4 * i));
    }
    memcpy(s0_out, &s0, 32);
    memcpy(s1_out, &s1, 32);
}
/* ------------------------------------------------------------------ TurboPlonK prover pieces (CPU prover baseline)
 * The bulk steps of prover_with_lagrange (/root/reference/uzkge/src/plonk/prover.rs:88-394) that are neither an MSM nor a
 * transform, for oracle/cpu_prover.py (the proofs/s CPU baseline, default feature set).  Parallel where the reference is
 * (t_poly's cfg_into_iter over the evaluation points, helpers.rs:284), serial where it is serial. */
static inline void fp_pow5(const field_t *F, fe *o, const fe *a) {
    fe a2, a4;
    fp_sqr(F, &a2, a);
    fp_sqr(F, &a4, &a2);
    fp_mul(F, o, &a4, a);
}
/* t_poly's loop body, helpers.rs:284-669, terms 1-11 (build without the `shuffle` feature), times 1 / Z_H:
 * cols = w[5] q[9] pi z s[5] coset_quotient l1 qb prk[4] (28 columns of m elements), sc = k[5] alpha beta gamma g g_inv */
void oracle_plonk_quotient(const uint64_t *const *cols, const uint64_t *sc, const uint64_t *z_h_inv, size_t factor, size_t m,
                           uint64_t *out) {
    oracle_init();
    const field_t *F = &FR;
    const fe *k = (const fe *)sc, *alpha = k + 5, *beta = k + 6, *gamma = k + 7, *g = k + 8, *g_inv = k + 9;
    fe a[10], g2p1;
    a[0] = F->r;
    for (int i = 1; i < 10; i++) fp_mul(F, &a[i], &a[i - 1], alpha);
    fp_mul(F, &g2p1, g, g);
    fp_add(F, &g2p1, &g2p1, &F->r);
#define COL(c, i) ((const fe *)cols[c] + (i))
#pragma omp parallel for schedule(static)
    for (size_t p = 0; p < m; p++) {
        const size_t pn = (p + factor) % m;
        fe w[5], t, u, num, term2, term3;
        for (int j = 0; j < 5; j++) w[j] = *COL(j, p);
        /* term 1: the gate function (turbo/mod.rs:193-222) */
        memset(&num, 0, sizeof(fe));
        for (int j = 0; j < 4; j++) { fp_mul(F, &t, COL(5 + j, p), &w[j]); fp_add(F, &num, &num, &t); }
        fe w01, w23;
        fp_mul(F, &w01, &w[0], &w[1]);
        fp_mul(F, &w23, &w[2], &w[3]);
        fp_mul(F, &t, COL(9, p), &w01); fp_add(F, &num, &num, &t);
        fp_mul(F, &t, COL(10, p), &w23); fp_add(F, &num, &num, &t);
        fp_add(F, &num, &num, COL(11, p));
        fp_add(F, &num, &num, COL(14, p));                      /* public input */
        fp_mul(F, &t, &w01, &w23); fp_mul(F, &t, &t, &w[4]); fp_mul(F, &t, &t, COL(12, p)); fp_add(F, &num, &num, &t);
        fp_mul(F, &t, COL(13, p), &w[4]); fp_sub(F, &num, &num, &t);
        /* terms 2, 3: the permutation argument */
        const fe *z = COL(15, p), *zn = COL(15, pn), *cq = COL(21, p);
        fp_mul(F, &term2, alpha, z);
        fp_mul(F, &term3, alpha, zn);
        for (int j = 0; j < 5; j++) {
            fp_mul(F, &t, beta, &k[j]); fp_mul(F, &t, &t, cq); fp_add(F, &t, &t, &w[j]); fp_add(F, &t, &t, gamma);
            fp_mul(F, &term2, &term2, &t);
            fp_mul(F, &t, beta, COL(16 + j, p)); fp_add(F, &t, &t, &w[j]); fp_add(F, &t, &t, gamma);
            fp_mul(F, &term3, &term3, &t);
        }
        fp_add(F, &num, &num, &term2);
        fp_sub(F, &num, &num, &term3);
        /* term 4: L_1 (z - 1) */
        fp_sub(F, &t, z, &F->r); fp_mul(F, &t, &t, COL(22, p)); fp_mul(F, &t, &t, &a[2]); fp_add(F, &num, &num, &t);
        /* terms 5-7: boolean constraints on wires 1..3 */
        for (int j = 1; j <= 3; j++) {
            fp_sub(F, &t, &w[j], &F->r); fp_mul(F, &t, &t, &w[j]); fp_mul(F, &t, &t, COL(23, p)); fp_mul(F, &t, &t, &a[2 + j]);
            fp_add(F, &num, &num, &t);
        }
        /* terms 8-11: Anemoi round */
        const fe *w0n = COL(0, pn), *w1n = COL(1, pn), *w2n = COL(2, pn);
        const fe *prk1 = COL(24, p), *prk2 = COL(25, p), *prk3 = COL(26, p), *prk4 = COL(27, p);
        fe w30, w21, w320, w221, tmp, p5, e;
        fp_add(F, &w30, &w[0], &w[3]); fp_add(F, &w21, &w[1], &w[2]);
        fp_add(F, &w320, &w[0], &w30); fp_add(F, &w221, &w[1], &w21);
        fp_mul(F, &tmp, g, &w21); fp_add(F, &tmp, &tmp, &w30); fp_add(F, &tmp, &tmp, prk3);
        fp_sub(F, &t, &tmp, w2n); fp_pow5(F, &p5, &t);
        /* 8 */
        fp_sqr(F, &u, &tmp); fp_mul(F, &u, &u, g); fp_add(F, &e, &p5, &u);
        fp_mul(F, &u, g, &w221); fp_add(F, &u, &u, &w320); fp_add(F, &u, &u, prk1); fp_sub(F, &e, &e, &u);
        fp_mul(F, &e, &e, prk3); fp_mul(F, &e, &e, &a[6]); fp_sub(F, &num, &num, &e);
        /* 10 */
        fp_sqr(F, &u, w2n); fp_mul(F, &u, &u, g); fp_add(F, &e, &p5, &u); fp_add(F, &e, &e, g_inv); fp_sub(F, &e, &e, w0n);
        fp_mul(F, &e, &e, prk3); fp_mul(F, &e, &e, &a[8]); fp_sub(F, &num, &num, &e);
        fp_mul(F, &tmp, g, &w30); fp_mul(F, &u, &g2p1, &w21); fp_add(F, &tmp, &tmp, &u); fp_add(F, &tmp, &tmp, prk4);
        fp_sub(F, &t, &tmp, &w[4]); fp_pow5(F, &p5, &t);
        /* 9 */
        fp_sqr(F, &u, &tmp); fp_mul(F, &u, &u, g); fp_add(F, &e, &p5, &u);
        fp_mul(F, &u, g, &w320); fp_mul(F, &t, &g2p1, &w221); fp_add(F, &u, &u, &t); fp_add(F, &u, &u, prk2); fp_sub(F, &e, &e, &u);
        fp_mul(F, &e, &e, prk3); fp_mul(F, &e, &e, &a[7]); fp_sub(F, &num, &num, &e);
        /* 11 */
        fp_sqr(F, &u, &w[4]); fp_mul(F, &u, &u, g); fp_add(F, &e, &p5, &u); fp_add(F, &e, &e, g_inv); fp_sub(F, &e, &e, w1n);
        fp_mul(F, &e, &e, prk3); fp_mul(F, &e, &e, &a[9]); fp_sub(F, &num, &num, &e);
        fp_mul(F, (fe *)out + p, &num, (const fe *)z_h_inv + (p % factor));
    }
#undef COL
}
/* z_poly's evaluations, helpers.rs:160-220: z[0] = 1, z[i+1] = z[i] * prod_j (w_j[i] + gamma + beta k_j g^i)
 *                                                                  / prod_j (w_j[i] + gamma + beta * sigma(j, i))
 * with one batch inversion of the n - 1 denominators.  perm holds the 5n targets of compute_permutation. */
int oracle_plonk_z_evals(const uint64_t *w_ext, const uint64_t *perm, const uint64_t *k5, const uint64_t *group,
                         const uint64_t *beta, const uint64_t *gamma, size_t n, uint64_t *out) {
    oracle_init();
    const field_t *F = &FR;
    const fe *w = (const fe *)w_ext, *k = (const fe *)k5, *grp = (const fe *)group;
    fe *z = (fe *)out;
    fe *num = (fe *)malloc(sizeof(fe) * n), *den = (fe *)malloc(sizeof(fe) * n), *pre = (fe *)malloc(sizeof(fe) * n);
    if (!num || !den || !pre) { free(num); free(den); free(pre); return 1; }
    fe bk[5];
    for (int j = 0; j < 5; j++) fp_mul(F, &bk[j], (const fe *)beta, &k[j]);
    const size_t n1 = n ? n - 1 : 0;
#pragma omp parallel for schedule(static) if (n >= 4096)
    for (size_t i = 0; i < n1; i++) {
        fe a = F->r, b = F->r, t;
        for (int j = 0; j < 5; j++) {
            const fe *f = &w[j * n + i];
            fp_mul(F, &t, &bk[j], &grp[i]); fp_add(F, &t, &t, f); fp_add(F, &t, &t, (const fe *)gamma);
            fp_mul(F, &a, &a, &t);
            const uint64_t pp = perm[j * n + i];
            fp_mul(F, &t, &bk[pp / n], &grp[pp % n]); fp_add(F, &t, &t, f); fp_add(F, &t, &t, (const fe *)gamma);
            fp_mul(F, &b, &b, &t);
        }
        num[i] = a;
        den[i] = b;
    }
    /* batch inversion (ark_ff::batch_inversion skips zeros) */
    fe acc = F->r;
    for (size_t i = 0; i + 1 < n; i++) {
        pre[i] = acc;
        if (!fe_is_zero(&den[i])) fp_mul(F, &acc, &acc, &den[i]);
    }
    fp_inv(F, &acc, &acc);
    for (size_t i = n - 1; i-- > 0;) {
        if (fe_is_zero(&den[i])) continue;
        fe inv;
        fp_mul(F, &inv, &acc, &pre[i]);
        fp_mul(F, &acc, &acc, &den[i]);
        den[i] = inv;
    }
    z[0] = F->r;
    for (size_t i = 0; i + 1 < n; i++) {
        fe t;
        fp_mul(F, &t, &z[i], &num[i]);
        fp_mul(F, &z[i + 1], &t, &den[i]);
    }
    free(num); free(den); free(pre);
    return 0;
}
/* out[i] = sum_j scalars[j] * polys[j][i], i < L (coefficients beyond lens[j] are zero): r_poly / batch_prove's h */
void oracle_fr_lincomb(const uint64_t *const *polys, const uint64_t *lens, const uint64_t *scalars, size_t cnt, uint64_t *out, size_t L) {
    oracle_init();
    const field_t *F = &FR;
#pragma omp parallel for schedule(static) if (L >= 4096)
    for (size_t i = 0; i < L; i++) {
        fe acc, t;
        memset(&acc, 0, sizeof(fe));
        for (size_t j = 0; j < cnt; j++) {
            if (i >= lens[j]) continue;
            fp_mul(F, &t, (const fe *)scalars + j, (const fe *)polys[j] + i);
            fp_add(F, &acc, &acc, &t);
        }
        ((fe *)out)[i] = acc;
    }
}
/* c(X) = q(X) (X - z) + rem: synthetic division, q has n - 1 coefficients */
void oracle_fr_div_linear(const uint64_t *coefs, size_t n, const uint64_t *z, uint64_t *quot, uint64_t *rem) {
    oracle_init();
    const field_t *F = &FR;
    const fe *c = (const fe *)coefs;
    fe acc;
    memset(&acc, 0, sizeof(fe));
    for (size_t i = n; i-- > 0;) {
        fp_mul(F, &acc, &acc, (const fe *)z);
        fp_add(F, &acc, &acc, &c[i]);
        if (i > 0) ((fe *)quot)[i - 1] = acc;
    }
    memcpy(rem, &acc, 32);
}
/* torchrun exports OMP_NUM_THREADS=1; the CPU arm of bench.py asks for every host core explicitly */
void oracle_set_num_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
