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
// ---- fast forward butterflies ------------------------------------------------------------------
// The forward transform of the reference never wraps 64 bits: in Butterfly (ntt.go:32-40) the Montgomery
// product is in [1,2q-1] for ANY 64-bit V (its high word is < psi < q), so X = U'+V <= max(U,4q) and
// Y = U'+2q-V <= max(U,4q) stay below 2^64 and every value remains congruent to the true transform; the
// final BRedAdd is canonical for any 64-bit word.  Hence NTT(x) of the reference equals the canonical
// negacyclic transform of (x mod q) for every input, and ANY exact lazy butterfly followed by a canonical
// reduction is bit-identical.  The two below use Shoup/Harvey multiplication by the twiddle w in plain
// form with ws = floor(w * 2^64 / q); they need 5-6 instead of 9 wide multiplies.
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
// q < 2^62, values kept in [0,4q): exact quotient floor(ws*Y/2^64) => T = w*Y - Q*q in [0,2q)
LG_DEV void butterfly_fwd_4q(u64& X, u64& Y, u64 w, u64 ws, u64 q, u64 twoq) {
    u64 x = X;
    if (x >= twoq) x -= twoq;
    const u64 qh = mul_hi(ws, Y);
    const u64 t = mul_lo(w, Y) - mul_lo(qh, q);
    X = x + t;
    Y = x + twoq - t;
}
// q < 2^56: the quotient is taken from the three high partial products only (at most 2 too small, so
// T = w*Y - Q*q is in [0,4q)) and no conditional subtraction is made: a value grows by at most 4q per
// stage, 16 stages add < 64q < 2^62 to inputs that the load clamps below 2^63.  Y may be any 64-bit word.
LG_DEV void butterfly_fwd_free(u64& X, u64& Y, u64 w, u64 ws, u64 q, u64 fourq) {
    const u32 a0 = (u32)ws, a1 = (u32)(ws >> 32), b0 = (u32)Y, b1 = (u32)(Y >> 32);
    const u64 m1 = mul_wide(a1, b0), m2 = mul_wide(a0, b1);
    const u64 qh = mad_wide(a1, b1, (m1 >> 32)) + (m2 >> 32);
    const u64 t = mul_lo(w, Y) - mul_lo(qh, q);
    const u64 x = X;
    X = x + t;
    Y = x + fourq - t;
}

// InvButterfly, ring/ntt.go:43-50
LG_DEV void butterfly_inv(u64& U, u64& V, u64 w, u64 q, u64 qinv, u64 twoq) {
    u64 x = U + V;
    u64 d = U + twoq - V;
    if (x > twoq) x -= twoq;
    U = x;
    V = mred_constant(d, w, q, qinv);
}
