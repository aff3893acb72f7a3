// modarith.cuh -- 64-bit Montgomery / Barrett device primitives (kernel family K2).
//
// Bit-exact restatements of ring/modular_reduction.go of the reference
// (Lattigo v1.3.1): every function returns the same uint64 word as the Go
// function it names, for every 64-bit input.  The 64x64 products are issued
// as PTX mul.lo.u64 / mul.hi.u64 so that NVVM cannot re-associate them;
// ptxas lowers each to the minimal 32-bit IMAD.WIDE chain (11 32x32
// multiplies per Montgomery reduction: 4 for the full product, 3 for the low
// product with q^-1, 4 for the high product with q).
#pragma once
#include <stdint.h>

typedef uint64_t u64;
typedef uint32_t u32;

#define LG_DEV __device__ __forceinline__

LG_DEV u64 mul_lo(u64 a, u64 b) {
    u64 r;
    asm("mul.lo.u64 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
LG_DEV u64 mul_hi(u64 a, u64 b) {
    u64 r;
    asm("mul.hi.u64 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// per-limb constants, loaded once per CTA
struct LimbConst {
    u64 q;     // modulus
    u64 qinv;  // q^-1 mod 2^64          (MRedParams, modular_reduction.go:53-64)
    u64 u0;    // hi word of floor(2^128/q) (BRedParams, :97-106) -- bredParams[0]
    u64 u1;    // lo word                                          -- bredParams[1]
};

// CRed, modular_reduction.go:211-216
LG_DEV u64 cred(u64 a, u64 q) { return a >= q ? a - q : a; }

// MRedConstant, modular_reduction.go:83-89: result in [0, 2q)
LG_DEV u64 mred_constant(u64 x, u64 y, u64 q, u64 qinv) {
    u64 alo = mul_lo(x, y);
    u64 ahi = mul_hi(x, y);
    u64 R = mul_lo(alo, qinv);
    u64 H = mul_hi(R, q);
    return ahi - H + q;
}
// MRed, modular_reduction.go:70-79
LG_DEV u64 mred(u64 x, u64 y, u64 q, u64 qinv) {
    u64 r = mred_constant(x, y, q, qinv);
    return r >= q ? r - q : r;
}
// BRedAddConstant / BRedAdd, modular_reduction.go:112-126 (u0 = bredParams[0])
LG_DEV u64 bred_add_constant(u64 x, u64 q, u64 u0) { return x - mul_hi(x, u0) * q; }
LG_DEV u64 bred_add(u64 x, u64 q, u64 u0) {
    u64 r = x - mul_hi(x, u0) * q;
    return r >= q ? r - q : r;
}
// BRedConstant / BRed, modular_reduction.go:133-207
LG_DEV u64 bred_constant(u64 x, u64 y, u64 q, u64 u0, u64 u1) {
    u64 alo = mul_lo(x, y), ahi = mul_hi(x, y);
    u64 lhi = mul_hi(alo, u1);
    u64 mhi = mul_hi(alo, u0), mlo = mul_lo(alo, u0);
    u64 s0 = mlo + lhi;
    u64 s1 = mhi + (s0 < mlo ? 1ull : 0ull);
    mhi = mul_hi(ahi, u1);
    mlo = mul_lo(ahi, u1);
    u64 t = mlo + s0;
    lhi = mhi + (t < mlo ? 1ull : 0ull);
    s0 = mul_lo(ahi, u0) + s1 + lhi;
    return alo - mul_lo(s0, q);
}
LG_DEV u64 bred(u64 x, u64 y, u64 q, u64 u0, u64 u1) {
    u64 r = bred_constant(x, y, q, u0, u1);
    return r >= q ? r - q : r;
}
// MFormConstant / MForm, modular_reduction.go:15-30
LG_DEV u64 mform_constant(u64 a, u64 q, u64 u0, u64 u1) {
    u64 mhi = mul_hi(a, u1);
    return (0ull - (mul_lo(a, u0) + mhi)) * q;
}
LG_DEV u64 mform(u64 a, u64 q, u64 u0, u64 u1) {
    u64 r = mform_constant(a, q, u0, u1);
    return r >= q ? r - q : r;
}
// InvMFormConstant / InvMForm, modular_reduction.go:34-49
LG_DEV u64 invmform(u64 a, u64 q, u64 qinv) {
    u64 r = q - mul_hi(mul_lo(a, qinv), q);
    return r >= q ? r - q : r;
}
// PowerOf2, ring/utils.go:8-17 (x in Montgomery form; n in [0,63])
LG_DEV u64 power_of_2(u64 x, u32 n, u64 q, u64 qinv) {
    u64 ahi = n ? (x >> (64 - n)) : 0ull, alo = x << n;
    u64 R = mul_lo(alo, qinv);
    u64 H = mul_hi(R, q);
    u64 r = ahi - H + q;
    return r >= q ? r - q : r;
}

// Butterfly, ring/ntt.go:32-40 (strict '>' as in the reference)
LG_DEV void butterfly_fwd(u64& U, u64& V, u64 w, u64 q, u64 qinv, u64 twoq) {
    u64 u = U;
    if (u > twoq) u -= twoq;
    u64 v = mred_constant(V, w, q, qinv);
    U = u + v;
    V = u + twoq - v;
}
// ---- fast butterflies ----------------------------------------------------------------------------
// The forward transform of the reference never wraps 64 bits: in Butterfly (ntt.go:32-40) the Montgomery
// product is in [1,2q-1] for ANY 64-bit V (its high word is < psi < q), so X = U'+V <= max(U,4q) and
// Y = U'+2q-V <= max(U,4q) stay below 2^64 and every value remains congruent to the true transform; the
// final BRedAdd is canonical for any 64-bit word.  Hence NTT(x) of the reference equals the canonical
// negacyclic transform of (x mod q) for every input, and ANY exact lazy butterfly followed by a canonical
// reduction is bit-identical.  The inverse transform (InvButterfly, ntt.go:43-50) has the same property
// only while no sum wraps, which holds when every input word is <= 2q (values then stay in [0,2q] and the
// final MRed by N^-1 is canonical); the kernels check that per limb and fall back to the literal
// butterflies otherwise.
//
// The fast butterflies multiply by the twiddle w in plain (non-Montgomery) form with the Shoup constant
// ws = floor(w * 2^64 / q): T = w*Y - Qh*q where Qh ~ floor(ws*Y / 2^64).
LG_DEV u64 mul_wide(u32 a, u32 b) {
    u64 r;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    return r;
}
LG_DEV u64 mad_wide(u32 a, u32 b, u64 c) {
    u64 r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
}
LG_DEV u32 mad_lo32(u32 a, u32 b, u32 c) {
    u32 r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
// Quotient estimate from the three high partial products of ws*y (the low x low product and the carries of
// the low halves of the cross products are dropped): at most 2 below floor(ws*y/2^64), for any 64-bit y.
LG_DEV u64 qhat3(u64 ws, u64 y) {
    const u32 a0 = (u32)ws, a1 = (u32)(ws >> 32), b0 = (u32)y, b1 = (u32)(y >> 32);
    const u64 m1 = mul_wide(a1, b0), m2 = mul_wide(a0, b1), p = mul_wide(a1, b1);
    u32 lo, hi;
    asm("{\n\t.reg .u32 t;\n\tadd.cc.u32 t, %2, %3;\n\taddc.u32 %1, %5, 0;\n\tadd.cc.u32 %0, t, %4;\n\taddc.u32 %1, %1, 0;\n\t}"
        : "=r"(lo), "=&r"(hi)
        : "r"((u32)(m1 >> 32)), "r"((u32)p), "r"((u32)(m2 >> 32)), "r"((u32)(p >> 32)));
    return ((u64)hi << 32) | lo;
}
// w*y + qh*nq (mod 2^64) with nq = -q: one multiply-add chain, no separate subtraction
LG_DEV u64 shoup_tail(u64 w, u64 y, u64 qh, u64 nq) {
    const u32 w0 = (u32)w, w1 = (u32)(w >> 32), b0 = (u32)y, b1 = (u32)(y >> 32);
    const u32 h0 = (u32)qh, h1 = (u32)(qh >> 32), n0 = (u32)nq, n1 = (u32)(nq >> 32);
    u64 t = mul_wide(w0, b0);
    t = mad_wide(h0, n0, t);
    u32 th = (u32)(t >> 32);
    th = mad_lo32(w0, b1, th);
    th = mad_lo32(w1, b0, th);
    th = mad_lo32(h0, n1, th);
    th = mad_lo32(h1, n0, th);
    return ((u64)th << 32) | (u32)t;
}
// w*y mod q up to a multiple of q: in [0,4q) for any 64-bit y (q < 2^62)
LG_DEV u64 shoup3(u64 w, u64 ws, u64 y, u64 nq) { return shoup_tail(w, y, qhat3(ws, y), nq); }
// exact quotient: in [0,2q)
LG_DEV u64 shoup_exact(u64 w, u64 ws, u64 y, u64 nq) { return shoup_tail(w, y, mul_hi(ws, y), nq); }

// q < 2^56, no conditional subtraction at all: a value grows by at most 4q per stage, 16 stages add
// < 64q < 2^62 to inputs that the load clamps below 2^63.  Y may be any 64-bit word.
LG_DEV void butterfly_fwd_free(u64& X, u64& Y, u64 w, u64 ws, u64 nq, u64 fourq) {
    const u64 t = shoup3(w, ws, Y, nq);
    const u64 x = X;
    X = x + t;
    Y = x + fourq - t;
}
// q < 2^61, values kept in [0,8q): X is brought below 4q, T is in [0,4q)
LG_DEV void butterfly_fwd_8q(u64& X, u64& Y, u64 w, u64 ws, u64 nq, u64 fourq) {
    const u64 t = shoup3(w, ws, Y, nq);
    u64 x = X;
    if (x >= fourq) x -= fourq;
    X = x + t;
    Y = x + fourq - t;
}
// Gentleman-Sande, q < 2^46 and inputs <= 2q: the sum path is never reduced (values double per stage,
// <= 4q*2^s after stage s < 2^63), m = q << (s+1) bounds Y before stage s so the difference stays positive.
LG_DEV void butterfly_inv_free(u64& X, u64& Y, u64 w, u64 ws, u64 nq, u64 m) {
    const u64 s = X + Y;
    const u64 d = X + m - Y;
    X = s;
    Y = shoup3(w, ws, d, nq);
}
// Gentleman-Sande, q < 2^61, values kept in [0,4q)
LG_DEV void butterfly_inv_4q(u64& X, u64& Y, u64 w, u64 ws, u64 nq, u64 fourq) {
    u64 s = X + Y;
    const u64 d = X + fourq - Y;
    if (s >= fourq) s -= fourq;
    X = s;
    Y = shoup3(w, ws, d, nq);
}

// ---- FP64-assisted quotient (moduli below 3*2^44, values below 2^52) ------------------------------------
// B200 has a full-rate FP64 pipe next to the integer multiplier, so the Shoup quotient can be taken there:
// with wd = RD(floor(w*2^64/q)) * 2^-64 <= w/q (absolute deficit < 2^-52) and cw = RD(2^52 - 2^52*wd),
//   qd = RD((2^52 + y) * wd + cw) = 2^52 + floor(y*wd - e),  0 <= e < 1      (one DFMA, round-down)
// holds the quotient estimate Qh = floor(y*w/q) - {0..3} in its mantissa for any y < 2^52, so
// T = w*y - Qh*q is in [0,4q).  Qh = -1 (y*wd < e, e.g. y = 0) shows as the double just below 2^52, i.e.
// mantissa all ones under exponent 0x432: subtracting 0x43300000 from the high word gives the two's complement
// high word of Qh in both cases.  That leaves 2 wide + 4 narrow integer multiplies, one LOP3, one IADD and the
// DFMA per product (13 instructions per butterfly with the two 64-bit adds).
LG_DEV double shoup_cw(double wd) { return __fma_rd(-4503599627370496.0, wd, 4503599627370496.0); }
LG_DEV u64 shoup_f64(u64 w, double wd, double cw, u64 y, u64 nq) {
    const u32 b0 = (u32)y, b1 = (u32)(y >> 32);
    const double qd = __fma_rd(__hiloint2double((int)(b1 | 0x43300000u), (int)b0), wd, cw);
    const u32 h0 = (u32)__double2loint(qd), h1 = (u32)__double2hiint(qd) - 0x43300000u;
    const u32 w0 = (u32)w, w1 = (u32)(w >> 32), n0 = (u32)nq, n1 = (u32)(nq >> 32);
    u64 t = mul_wide(w0, b0);
    t = mad_wide(h0, n0, t);
    u32 th = (u32)(t >> 32);
    th = mad_lo32(w0, b1, th);
    th = mad_lo32(w1, b0, th);
    th = mad_lo32(h0, n1, th);
    th = mad_lo32(h1, n0, th);
    return ((u64)th << 32) | (u32)t;
}
// forward, no conditional subtraction: inputs below 2^50, 16 stages add < 64q < 2^52 - 2^50
LG_DEV void butterfly_fwd_f64(u64& X, u64& Y, u64 w, double wd, double cw, u64 nq, u64 fourq) {
    const u64 t = shoup_f64(w, wd, cw, Y, nq);
    const u64 x = X;
    X = x + t;
    Y = x + fourq - t;
}
// Gentleman-Sande, values kept in [0,4q): the difference is below 8q < 2^52 (microbenchmark only: the
// reduction-free integer butterfly is faster)
LG_DEV void butterfly_inv_f64(u64& X, u64& Y, u64 w, double wd, double cw, u64 nq, u64 fourq) {
    u64 s = X + Y;
    const u64 d = X + fourq - Y;
    if (s >= fourq) s -= fourq;
    X = s;
    Y = shoup_f64(w, wd, cw, d, nq);
}

// ---- FP64-only butterflies (moduli below 3*2^44) ---------------------------------------------------------------
// On B200 every pipe these kernels use (integer multiply, integer ALU, FP64) is fed by one dispatch port per SM
// sub-partition that sustains about one warp instruction every two cycles: throughput follows the INSTRUCTION COUNT,
// whatever the mix (profiles/r02_fp64_butterfly.txt: the 13-instruction butterfly above runs at 5.48 per clock per SM,
// the 8-instruction one below at 7.13, any mix of the two in between).  So the shortest exact butterfly wins, and that
// is one that never touches the integer pipes: values are kept as DOUBLES holding exact signed integers.
//     h  = RN(w*y),  l = fma(w, y, -h)              the product as an exact two-word sum (an FMA's error term is exact)
//     qh = RD(y*wd + 1.5*2^52) - 1.5*2^52           wd = RD(w/q) (deficit below 2^-52): qh = floor(y*w/q) + {-1, 0, +1}
//     t  = fma(-qh, q, h) + l                       both steps exact: |h - qh*q| < 2^48;  t = w*y - qh*q in [-q, 2q)
// for |y| < 2^51 (the magic-number floor needs |y*wd| < 2^51).  A Cooley-Tukey stage adds at most 2q to the magnitude:
// 16 stages from an input below 2^49 stay below 2^49 + 32q < 2^51 for q < 3*2^44.  The values a kernel leaves in HBM
// for the next phase are the raw doubles; integers are converted on the way in (1 LOP3 + 1 DADD) and out.
#define LG_D64_MAGIC 6755399441055744.0  // 1.5 * 2^52
LG_DEV double bits2d(u64 b) { return __longlong_as_double((long long)b); }
LG_DEV u64 d2bits(double d) { return (u64)__double_as_longlong(d); }
// integer v < 2^52 -> the double v - off, c = 2^52 + off (exactly representable)
LG_DEV double u52_to_d(u64 v, double c) {
    return __dadd_rn(__hiloint2double((int)((u32)(v >> 32) | 0x43300000u), (int)(u32)v), -c);
}
// double v (an integer) with 0 <= v + off < 2^52 -> the integer v + off, c = 2^52 + off
LG_DEV u64 d_to_u52(double v, double c) { return d2bits(__dadd_rn(v, c)) & 0x000FFFFFFFFFFFFFull; }
// w*y - qh*q in [-q, 2q); in [0, 2q) when y >= 0
LG_DEV double d64_mul(double w, double wd, double y, double q) {
    const double h = __dmul_rn(w, y);
    const double l = __fma_rn(w, y, -h);
    const double qh = __dadd_rn(__fma_rd(y, wd, LG_D64_MAGIC), -LG_D64_MAGIC);
    return __dadd_rn(__fma_rn(-qh, q, h), l);
}
// v - floor~(v/q)*q in [-q, 2q) for |v| < 2^51 (qinvd = RD(1/q)); in [0, 2q) when v >= 0
LG_DEV double d64_red(double v, double qinvd, double q) {
    const double c = __dadd_rn(__fma_rd(v, qinvd, LG_D64_MAGIC), -LG_D64_MAGIC);
    return __fma_rn(-c, q, v);
}
// X, Y hold the bit patterns of the doubles
LG_DEV void butterfly_fwd_d64(u64& X, u64& Y, u64 wb, u64 wdb, double q) {
    const double t = d64_mul(bits2d(wb), bits2d(wdb), bits2d(Y), q);
    const double x = bits2d(X);
    X = d2bits(__dadd_rn(x, t));
    Y = d2bits(__dadd_rn(x, -t));
}
// Gentleman-Sande: the sum path doubles per stage (callers reduce every four stages), the product path ends in [-q, 2q)
LG_DEV void butterfly_inv_d64(u64& X, u64& Y, u64 wb, u64 wdb, double q) {
    const double x = bits2d(X), y = bits2d(Y);
    X = d2bits(__dadd_rn(x, y));
    Y = d2bits(d64_mul(bits2d(wb), bits2d(wdb), __dadd_rn(x, -y), q));
}

// InvButterfly, ring/ntt.go:43-50
LG_DEV void butterfly_inv(u64& U, u64& V, u64 w, u64 q, u64 qinv, u64 twoq) {
    u64 x = U + V;
    u64 d = U + twoq - V;
    if (x > twoq) x -= twoq;
    U = x;
    V = mred_constant(d, w, q, qinv);
}
