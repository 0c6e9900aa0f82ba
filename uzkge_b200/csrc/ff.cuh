// ff.cuh -- BN254 Fq / Fr arithmetic on 8 x u32 limbs, Montgomery form R = 2^256.
//
// This is the device-side replacement of ark_ff::Fp<MontBackend<_, 4>> (ark-ff-zypher 0.4 with the `asm`
// feature, /root/reference/Cargo.toml:29) that every call on the hot path bottoms out in.  In-memory
// layout is identical (4 x u64 little-endian limbs == 8 x u32 little-endian limbs), so buffers cross the
// C ABI without conversion.
//
// Multiplication is an operand-scanning Montgomery product built on 32-bit IMAD carry chains
// (mad.lo.cc / madc.hi.cc pairs, which ptxas fuses into IMAD.WIDE.U32(.X)): the partial products of the
// even and the odd limbs of the multiplicand are accumulated in two separate 8-limb rows so that each row
// is ONE uninterrupted carry chain; the rows swap roles after every 32-bit reduction step instead of being
// shifted.  8 x (8 + 8 + 1) = 136 IMAD per product (SASS: 120 IMAD.WIDE.U32[.X] + 8 IMAD.HI + 8 IMAD).
//
// Tried and rejected on B200 (measured): replacing the low-limb product of the Fr reduction steps
// (r = 1 mod 2^28) by shifts saves 8 IMAD.HI but lengthens the dependency chain: the 2^22 NTT got 3.5 % slower;
// forming the reduction multiplier with shifts made ptxas stop fusing mad.lo/madc.hi pairs into IMAD.WIDE.
//
// Every primitive has a plain-C twin for host compilation (tests/host, and the few host-side constants the
// library derives); the twin is bit-identical by construction and is what the CPU test-suite exercises.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define UZ_HD __host__ __device__ __forceinline__
#define UZ_D __device__ __forceinline__
#else
#define UZ_HD inline
#define UZ_D inline
#endif

namespace uz {

struct alignas(16) fe {
    uint32_t l[8];
};

// ------------------------------------------------------------------ field parameters
// Moduli: SURVEY 8c-S1.  M0 = -p^-1 mod 2^32.  ONE = R mod p, R2 = R^2 mod p (checked in tests/host).
struct FqP {
    static constexpr uint32_t M0 = 0xe4866389u;
    UZ_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xd87cfd47u, 0x3c208c16u, 0x6871ca8du, 0x97816a91u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    UZ_HD static constexpr uint32_t one(int i) {
        constexpr uint32_t m[8] = {0xc58f0d9du, 0xd35d438du, 0xf5c70b3du, 0x0a78eb28u,
                                   0x7879462cu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    UZ_HD static constexpr uint32_t r2(int i) {
        constexpr uint32_t m[8] = {0x538afa89u, 0xf32cfc5bu, 0xd44501fbu, 0xb5e71911u,
                                   0x0a417ff6u, 0x47ab1effu, 0xcab8351fu, 0x06d89f71u};
        return m[i];
    }
};
struct FrP {
    static constexpr uint32_t M0 = 0xefffffffu;
    UZ_HD static constexpr uint32_t mod(int i) {
        constexpr uint32_t m[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                                   0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
        return m[i];
    }
    UZ_HD static constexpr uint32_t one(int i) {
        constexpr uint32_t m[8] = {0x4ffffffbu, 0xac96341cu, 0x9f60cd29u, 0x36fc7695u,
                                   0x7879462eu, 0x666ea36fu, 0x9a07df2fu, 0x0e0a77c1u};
        return m[i];
    }
    UZ_HD static constexpr uint32_t r2(int i) {
        constexpr uint32_t m[8] = {0xae216da7u, 0x1bb8e645u, 0xe35c59e3u, 0x53fe3ab1u,
                                   0x53bb8085u, 0x8c49833du, 0x7f4e44a5u, 0x0216d0b1u};
        return m[i];
    }
};

// ------------------------------------------------------------------ carry-chain primitives
// acc[0..7] += {x0, x2, x4, x6} * b  laid out lo,hi,lo,hi,...; the carry out of limb 7 is added to `top`.
UZ_HD void mad_row(uint32_t* acc, uint32_t& top, uint32_t x0, uint32_t x2, uint32_t x4, uint32_t x6, uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("mad.lo.cc.u32  %0, %9,  %13, %0;\n\t"
        "madc.hi.cc.u32 %1, %9,  %13, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %4;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %5;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, %6;\n\t"
        "madc.hi.cc.u32 %7, %12, %13, %7;\n\t"
        "addc.u32       %8, %8, 0;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7]), "+r"(top)
        : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(b));
#else
    const uint32_t x[4] = {x0, x2, x4, x6};
    uint64_t carry = 0;
    for (int k = 0; k < 4; k++) {
        uint64_t prod = (uint64_t)x[k] * b;
        uint64_t lo = (uint64_t)acc[2 * k] + (uint32_t)prod + carry;
        acc[2 * k] = (uint32_t)lo;
        uint64_t hi = (uint64_t)acc[2 * k + 1] + (uint32_t)(prod >> 32) + (lo >> 32);
        acc[2 * k + 1] = (uint32_t)hi;
        carry = hi >> 32;
    }
    top += (uint32_t)carry;
#endif
}

// Same chain without the trailing carry capture (caller guarantees limb 7 cannot overflow).
UZ_HD void mad_row_nc(uint32_t* acc, uint32_t x0, uint32_t x2, uint32_t x4, uint32_t x6, uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("mad.lo.cc.u32  %0, %8,  %12, %0;\n\t"
        "madc.hi.cc.u32 %1, %8,  %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %9,  %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %9,  %12, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %5;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, %6;\n\t"
        "madc.hi.u32    %7, %11, %12, %7;"
        : "+r"(acc[0]), "+r"(acc[1]), "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]),
          "+r"(acc[7])
        : "r"(x0), "r"(x2), "r"(x4), "r"(x6), "r"(b));
#else
    uint32_t top = 0;
    mad_row(acc, top, x0, x2, x4, x6, b);
#endif
}

// The stale row `o` (weights -1..6 after the implicit >> 32, o[0] == 0) becomes the odd row (weights 1..8):
//   e0 += o[1]            (weight 0; its carry enters the chain below)
//   o'[k] = o[k + 2] + {x1, x3, x5, x7} * b   (lo,hi,lo,hi,...), o[8] = o[9] = 0
UZ_HD void mad_row_shift(uint32_t* o, uint32_t& e0, uint32_t x1, uint32_t x3, uint32_t x5, uint32_t x7, uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32     %8, %8, %1;\n\t"
        "madc.lo.cc.u32 %0, %9,  %13, %2;\n\t"
        "madc.hi.cc.u32 %1, %9,  %13, %3;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %4;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %5;\n\t"
        "madc.lo.cc.u32 %4, %11, %13, %6;\n\t"
        "madc.hi.cc.u32 %5, %11, %13, %7;\n\t"
        "madc.lo.cc.u32 %6, %12, %13, 0;\n\t"
        "madc.hi.u32    %7, %12, %13, 0;"
        : "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7]), "+r"(e0)
        : "r"(x1), "r"(x3), "r"(x5), "r"(x7), "r"(b));
#else
    const uint32_t x[4] = {x1, x3, x5, x7};
    uint64_t s = (uint64_t)e0 + o[1];
    e0 = (uint32_t)s;
    uint64_t carry = s >> 32;
    uint32_t in[10];
    for (int k = 0; k < 8; k++) in[k] = o[k];
    in[8] = in[9] = 0;
    for (int k = 0; k < 4; k++) {
        uint64_t prod = (uint64_t)x[k] * b;
        uint64_t lo = (uint64_t)in[2 * k + 2] + (uint32_t)prod + carry;
        o[2 * k] = (uint32_t)lo;
        uint64_t hi = (uint64_t)in[2 * k + 3] + (uint32_t)(prod >> 32) + (lo >> 32);
        o[2 * k + 1] = (uint32_t)hi;
        carry = hi >> 32;
    }
#endif
}

// r = a + b (8 limbs), returns nothing: inputs < 2^255 so no carry out.
UZ_HD void add8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32  %0, %8,  %16;\n\t"
        "addc.cc.u32 %1, %9,  %17;\n\t"
        "addc.cc.u32 %2, %10, %18;\n\t"
        "addc.cc.u32 %3, %11, %19;\n\t"
        "addc.cc.u32 %4, %12, %20;\n\t"
        "addc.cc.u32 %5, %13, %21;\n\t"
        "addc.cc.u32 %6, %14, %22;\n\t"
        "addc.u32    %7, %15, %23;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
          "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a[i] + b[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
#endif
}

// r = a + b (8 limbs); returns the carry out (0 or 1).
UZ_HD uint32_t add8c(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t carry;
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32  %0, %9,  %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32    %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(carry)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
          "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (uint64_t)a[i] + b[i];
        r[i] = (uint32_t)c;
        c >>= 32;
    }
    carry = (uint32_t)c;
#endif
    return carry;
}

// r = a - b (8 limbs); returns 0xffffffff if a < b (borrow), else 0.
UZ_HD uint32_t sub8(uint32_t* r, const uint32_t* a, const uint32_t* b) {
    uint32_t borrow;
#if defined(__CUDA_ARCH__)
    asm("sub.cc.u32  %0, %9,  %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32    %8, 0, 0;"
        : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]),
          "=&r"(borrow)
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]),
          "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
#else
    int64_t c = 0;
    for (int i = 0; i < 8; i++) {
        c += (int64_t)a[i] - (int64_t)b[i];
        r[i] = (uint32_t)c;
        c >>= 32;  // arithmetic shift: 0 or -1
    }
    borrow = (uint32_t)c;
#endif
    return borrow;
}

// ------------------------------------------------------------------ field operations (all values fully reduced)
template <class P>
UZ_HD void load_mod(uint32_t* m) {
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = P::mod(i);
}

template <class P>
UZ_HD fe fe_one() {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = P::one(i);
    return r;
}
UZ_HD fe fe_zero() {
    fe r;
#pragma unroll
    for (int i = 0; i < 8; i++) r.l[i] = 0;
    return r;
}
UZ_HD bool fe_is_zero(const fe& a) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t |= a.l[i];
    return t == 0;
}
UZ_HD bool fe_eq(const fe& a, const fe& b) {
    uint32_t t = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) t |= a.l[i] ^ b.l[i];
    return t == 0;
}

// r - p if r >= p  (r < 2p)
template <class P>
UZ_HD void final_sub(uint32_t* r) {
    uint32_t m[8], t[8];
    load_mod<P>(m);
    uint32_t borrow = sub8(t, r, m);
#pragma unroll
    for (int i = 0; i < 8; i++) r[i] = borrow ? r[i] : t[i];
}

template <class P>
UZ_HD fe fe_add(const fe& a, const fe& b) {
    fe r;
    add8(r.l, a.l, b.l);
    final_sub<P>(r.l);
    return r;
}

template <class P>
UZ_HD fe fe_sub(const fe& a, const fe& b) {
    fe r;
    uint32_t borrow = sub8(r.l, a.l, b.l);
    uint32_t m[8];
#pragma unroll
    for (int i = 0; i < 8; i++) m[i] = P::mod(i) & borrow;
    add8(r.l, r.l, m);
    return r;
}

template <class P>
UZ_HD fe fe_neg(const fe& a) {
    fe z = fe_zero();
    return fe_sub<P>(z, a);
}

template <class P>
UZ_HD fe fe_dbl(const fe& a) {
    return fe_add<P>(a, a);
}

// Montgomery product a * b / R mod p, inputs < p, output < p.
template <class P>
UZ_HD fe fe_mul(const fe& a, const fe& b) {
    uint32_t ev[8], od[8];
    // step 0: rows start as plain products
    {
        const uint32_t bi = b.l[0];
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint64_t pe = (uint64_t)a.l[2 * k] * bi;
            uint64_t po = (uint64_t)a.l[2 * k + 1] * bi;
            ev[2 * k] = (uint32_t)pe;
            ev[2 * k + 1] = (uint32_t)(pe >> 32);
            od[2 * k] = (uint32_t)po;
            od[2 * k + 1] = (uint32_t)(po >> 32);
        }
        const uint32_t m = ev[0] * P::M0;
        mad_row_nc(od, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
        mad_row(ev, od[7], P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);
    }
#pragma unroll
    for (int i = 1; i < 8; i++) {
        // roles alternate: the row that was "odd" is aligned at weight 0 after the implicit >> 32
        uint32_t* E = (i & 1) ? od : ev;
        uint32_t* O = (i & 1) ? ev : od;
        const uint32_t bi = b.l[i];
        mad_row_shift(O, E[0], a.l[1], a.l[3], a.l[5], a.l[7], bi);
        mad_row(E, O[7], a.l[0], a.l[2], a.l[4], a.l[6], bi);
        const uint32_t m = E[0] * P::M0;
        mad_row_nc(O, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);
        mad_row(E, O[7], P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);
    }
    // after 8 steps: E = od (E[0] == 0), O = ev;  result = ev[0..7] + od[1..7]
    fe r;
    uint32_t sh[8];
#pragma unroll
    for (int i = 0; i < 7; i++) sh[i] = od[i + 1];
    sh[7] = 0;
    add8(r.l, ev, sh);
    final_sub<P>(r.l);
    return r;
}

// ------------------------------------------------------------------ dedicated squaring
// a^2 = sum_i a_i B^i (a_i B^i + sum_{j > i} 2 a_j B^j), B = 2^32.  With d = 2 a (fits 8 limbs: a < p < 2^254) the inner sum is the part
// of d above limb i minus the bit that a_i's top bit shifted into d_{i+1}: step i of the operand scan multiplies the scalar a_i with
// the vector c = [0 .. 0, a_i, a_{i+1} << 1, d_{i+2} .. d_7] -- 8 - i products instead of 8.  Everything else is fe_mul: the same two
// alternating 8-limb rows, the same interleaved reduction; the row primitives below are mad_row / mad_row_shift with their leading
// zero products removed (the host twins simply pass zeros).  36 + 72 = 108 multiplier instructions instead of 136.
// mad_row whose first 1 multiplicand limb(s) are zero: the chain starts at limb 2.
UZ_HD void mad_row_z1(uint32_t* acc, uint32_t& top, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("mad.lo.cc.u32 %0, %7, %10, %0;\n\t"
        "madc.hi.cc.u32 %1, %7, %10, %1;\n\t"
        "madc.lo.cc.u32 %2, %8, %10, %2;\n\t"
        "madc.hi.cc.u32 %3, %8, %10, %3;\n\t"
        "madc.lo.cc.u32 %4, %9, %10, %4;\n\t"
        "madc.hi.cc.u32 %5, %9, %10, %5;\n\t"
        "addc.u32       %6, %6, 0;"
        : "+r"(acc[2]), "+r"(acc[3]), "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
        : "r"(x1), "r"(x2), "r"(x3), "r"(b));
#else
    mad_row(acc, top, 0, x1, x2, x3, b);
#endif
}
// mad_row whose first 2 multiplicand limb(s) are zero: the chain starts at limb 4.
UZ_HD void mad_row_z2(uint32_t* acc, uint32_t& top, uint32_t x2, uint32_t x3, uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("mad.lo.cc.u32 %0, %5, %7, %0;\n\t"
        "madc.hi.cc.u32 %1, %5, %7, %1;\n\t"
        "madc.lo.cc.u32 %2, %6, %7, %2;\n\t"
        "madc.hi.cc.u32 %3, %6, %7, %3;\n\t"
        "addc.u32       %4, %4, 0;"
        : "+r"(acc[4]), "+r"(acc[5]), "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
        : "r"(x2), "r"(x3), "r"(b));
#else
    mad_row(acc, top, 0, 0, x2, x3, b);
#endif
}
// mad_row whose first 3 multiplicand limb(s) are zero: the chain starts at limb 6.
UZ_HD void mad_row_z3(uint32_t* acc, uint32_t& top, uint32_t x3, uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("mad.lo.cc.u32 %0, %3, %4, %0;\n\t"
        "madc.hi.cc.u32 %1, %3, %4, %1;\n\t"
        "addc.u32       %2, %2, 0;"
        : "+r"(acc[6]), "+r"(acc[7]), "+r"(top)
        : "r"(x3), "r"(b));
#else
    mad_row(acc, top, 0, 0, 0, x3, b);
#endif
}
// mad_row_shift whose first 1 multiplicand limb(s) are zero: those slots only move the stale row and pass the carry on.
UZ_HD void mad_row_shift_z1(uint32_t* o, uint32_t& e0, uint32_t x1, uint32_t x2, uint32_t x3, uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32     %8, %8, %1;\n\t"
        "addc.cc.u32    %0, %2, 0;\n\t"
        "addc.cc.u32    %1, %3, 0;\n\t"
        "madc.lo.cc.u32 %2, %9, %12, %4;\n\t"
        "madc.hi.cc.u32 %3, %9, %12, %5;\n\t"
        "madc.lo.cc.u32 %4, %10, %12, %6;\n\t"
        "madc.hi.cc.u32 %5, %10, %12, %7;\n\t"
        "madc.lo.cc.u32 %6, %11, %12, 0;\n\t"
        "madc.hi.u32 %7, %11, %12, 0;"
        : "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7]), "+r"(e0)
        : "r"(x1), "r"(x2), "r"(x3), "r"(b));
#else
    mad_row_shift(o, e0, 0, x1, x2, x3, b);
#endif
}
// mad_row_shift whose first 2 multiplicand limb(s) are zero: those slots only move the stale row and pass the carry on.
UZ_HD void mad_row_shift_z2(uint32_t* o, uint32_t& e0, uint32_t x2, uint32_t x3, uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32     %8, %8, %1;\n\t"
        "addc.cc.u32    %0, %2, 0;\n\t"
        "addc.cc.u32    %1, %3, 0;\n\t"
        "addc.cc.u32    %2, %4, 0;\n\t"
        "addc.cc.u32    %3, %5, 0;\n\t"
        "madc.lo.cc.u32 %4, %9, %11, %6;\n\t"
        "madc.hi.cc.u32 %5, %9, %11, %7;\n\t"
        "madc.lo.cc.u32 %6, %10, %11, 0;\n\t"
        "madc.hi.u32 %7, %10, %11, 0;"
        : "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7]), "+r"(e0)
        : "r"(x2), "r"(x3), "r"(b));
#else
    mad_row_shift(o, e0, 0, 0, x2, x3, b);
#endif
}
// mad_row_shift whose first 3 multiplicand limb(s) are zero: those slots only move the stale row and pass the carry on.
UZ_HD void mad_row_shift_z3(uint32_t* o, uint32_t& e0, uint32_t x3, uint32_t b) {
#if defined(__CUDA_ARCH__)
    asm("add.cc.u32     %8, %8, %1;\n\t"
        "addc.cc.u32    %0, %2, 0;\n\t"
        "addc.cc.u32    %1, %3, 0;\n\t"
        "addc.cc.u32    %2, %4, 0;\n\t"
        "addc.cc.u32    %3, %5, 0;\n\t"
        "addc.cc.u32    %4, %6, 0;\n\t"
        "addc.cc.u32    %5, %7, 0;\n\t"
        "madc.lo.cc.u32 %6, %9, %10, 0;\n\t"
        "madc.hi.u32 %7, %9, %10, 0;"
        : "+r"(o[0]), "+r"(o[1]), "+r"(o[2]), "+r"(o[3]), "+r"(o[4]), "+r"(o[5]), "+r"(o[6]), "+r"(o[7]), "+r"(e0)
        : "r"(x3), "r"(b));
#else
    mad_row_shift(o, e0, 0, 0, 0, x3, b);
#endif
}

template <class P>
UZ_HD fe fe_sqr(const fe& a) {
    const uint32_t* x = a.l;
    uint32_t d[8];                      // d[j] = limb j of 2 a, j >= 2 (d[j] for the slot right above the diagonal is x[j] << 1)
#pragma unroll
    for (int j = 2; j < 8; j++) d[j] = (x[j] << 1) | (x[j - 1] >> 31);
    uint32_t ev[8], od[8];
    {   // step 0: c = [x0, x1 << 1, d2 .. d7], plain products
        const uint32_t bi = x[0];
        const uint32_t c[8] = {x[0], x[1] << 1, d[2], d[3], d[4], d[5], d[6], d[7]};
#pragma unroll
        for (int k = 0; k < 4; k++) {
            uint64_t pe = (uint64_t)c[2 * k] * bi;
            uint64_t po = (uint64_t)c[2 * k + 1] * bi;
            ev[2 * k] = (uint32_t)pe;
            ev[2 * k + 1] = (uint32_t)(pe >> 32);
            od[2 * k] = (uint32_t)po;
            od[2 * k + 1] = (uint32_t)(po >> 32);
        }
    }
#define UZ_SQR_REDUCE(E, O)                                                   \
    {                                                                         \
        const uint32_t m = E[0] * P::M0;                                      \
        mad_row_nc(O, P::mod(1), P::mod(3), P::mod(5), P::mod(7), m);         \
        mad_row(E, O[7], P::mod(0), P::mod(2), P::mod(4), P::mod(6), m);      \
    }
    UZ_SQR_REDUCE(ev, od)
    // step 1 (E = od, O = ev): c = [0, x1, x2 << 1, d3 .. d7]
    mad_row_shift(ev, od[0], x[1], d[3], d[5], d[7], x[1]);
    mad_row_z1(od, ev[7], x[2] << 1, d[4], d[6], x[1]);
    UZ_SQR_REDUCE(od, ev)
    // step 2 (E = ev, O = od): c = [0, 0, x2, x3 << 1, d4 .. d7]
    mad_row_shift_z1(od, ev[0], x[3] << 1, d[5], d[7], x[2]);
    mad_row_z1(ev, od[7], x[2], d[4], d[6], x[2]);
    UZ_SQR_REDUCE(ev, od)
    // step 3: c = [0, 0, 0, x3, x4 << 1, d5, d6, d7]
    mad_row_shift_z1(ev, od[0], x[3], d[5], d[7], x[3]);
    mad_row_z2(od, ev[7], x[4] << 1, d[6], x[3]);
    UZ_SQR_REDUCE(od, ev)
    // step 4: c = [0 x 4, x4, x5 << 1, d6, d7]
    mad_row_shift_z2(od, ev[0], x[5] << 1, d[7], x[4]);
    mad_row_z2(ev, od[7], x[4], d[6], x[4]);
    UZ_SQR_REDUCE(ev, od)
    // step 5: c = [0 x 5, x5, x6 << 1, d7]
    mad_row_shift_z2(ev, od[0], x[5], d[7], x[5]);
    mad_row_z3(od, ev[7], x[6] << 1, x[5]);
    UZ_SQR_REDUCE(od, ev)
    // step 6: c = [0 x 6, x6, x7 << 1]
    mad_row_shift_z3(od, ev[0], x[7] << 1, x[6]);
    mad_row_z3(ev, od[7], x[6], x[6]);
    UZ_SQR_REDUCE(ev, od)
    // step 7: c = [0 x 7, x7]: the even row gets no product
    mad_row_shift_z3(ev, od[0], x[7], x[7]);
    UZ_SQR_REDUCE(od, ev)
#undef UZ_SQR_REDUCE
    // as in fe_mul: E = od (E[0] == 0), O = ev;  result = ev[0..7] + od[1..7]
    fe r;
    uint32_t sh[8];
#pragma unroll
    for (int i = 0; i < 7; i++) sh[i] = od[i + 1];
    sh[7] = 0;
    add8(r.l, ev, sh);
    final_sub<P>(r.l);
    return r;
}

template <class P>
UZ_HD fe fe_from_mont(const fe& a) {
    fe one = fe_zero();
    one.l[0] = 1;
    return fe_mul<P>(a, one);
}
template <class P>
UZ_HD fe fe_to_mont(const fe& a) {
    fe r2;
#pragma unroll
    for (int i = 0; i < 8; i++) r2.l[i] = P::r2(i);
    return fe_mul<P>(a, r2);
}

// a^e for a 64-bit exponent (table builders)
template <class P>
UZ_HD fe fe_pow_u64(fe a, uint64_t e) {
    fe acc = fe_one<P>();
    while (e) {
        if (e & 1) acc = fe_mul<P>(acc, a);
        a = fe_sqr<P>(a);
        e >>= 1;
    }
    return acc;
}

// a^(p-2): Fermat inversion (rare paths only: table setup, normalisation)
template <class P>
UZ_HD fe fe_inv(const fe& a) {
    uint32_t e[8];
    load_mod<P>(e);
    e[0] -= 2;  // p is odd and p[0] >= 2: no borrow
    fe acc = fe_one<P>();
    for (int i = 255; i >= 0; i--) {
        acc = fe_sqr<P>(acc);
        if ((e[i >> 5] >> (i & 31)) & 1) acc = fe_mul<P>(acc, a);
    }
    return acc;
}

}  // namespace uz
