// ntt.cu -- kernel family K1: batched per-limb negacyclic NTT / InvNTT, and the key-switch digit loop
// fused with the last NTT phase.
//
// Replaces ring/ntt.go:53-139 of the reference (Cooley-Tukey forward with lazy butterflies and a
// final BRedAdd; Gentleman-Sande inverse with a final MRed by N^-1).
//
// Bit-exactness (see modarith.cuh).  The forward transform of the reference never wraps, so its output is
// the canonical transform of (x mod q) for EVERY 64-bit input; the forward kernels use exact lazy
// Shoup butterflies (9 32x32 multiplies, 16 instructions) and still match the reference bit for bit,
// including on the unreduced words its own benchmarks feed (tests: "words" cases).  The inverse transform
// has that property only while every input word is <= 2q: callers either know that (outputs of our own
// canonical kernels) or run lg_launch_range_flags first, and flagged limbs take the literal InvButterfly.
// Moduli >= 2^61 and LATTIGPU_LITERAL_NTT=1 use the literal butterflies throughout (A/B and cross-check).
//
// Schedule (N = 2^logN, one limb = N words; the CTAs that share a limb's twiddles and key tile run together and hit L2:
// limbs are the slowest grid dimension; the contiguous phases walk the batch fastest, the strided phases the tiles):
//   logN <= 11 : one CTA per limb, radix-2 stages in shared memory.
//   logN >= 12 : two phases of register-resident radix-16 blocks (16 coefficients per thread, 4 stages,
//                8 independent butterflies per stage)
//     "strided" phase : the top L = logN-8 stages; a 256-thread CTA owns all 2^L rows of
//                       W = 4096/2^L adjacent columns (coalesced 8*W-byte rows) and re-distributes
//                       through shared memory once; forward and HBM-bound: the tile moves by TMA through a ring of
//                       three buffers (ntt_fwd_strided_tma),
//     "contig"  phase : the low 8 stages on contiguous 256-word segments; 16 threads own a segment, so
//                       the exchange is warp-synchronous (no CTA barrier) and every thread ends with 16
//                       consecutive words that move with 256-bit loads/stores.
//   Twiddle index for the butterfly on (j, j+2^s): (N >> (s+1)) + (j >> (s+1))
//   in both directions (ring/ntt.go:74 and :120).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace {

// butterfly flavours: forward LAZY = values in [0,8q), inverse LAZY = values in [0,4q)
// forward M_F64 = M_FREE with the quotient taken on the FP64 pipe (moduli below 3*2^44)
// M_D64 = FP64-only butterflies on doubles (moduli below 3*2^44; modarith.cuh), both directions: the default for those
// moduli, M_F64 / M_FREE remain behind the "no_d64_ntt" switch
enum { M_LITERAL = 0, M_FREE = 1, M_LAZY = 2, M_F64 = 3, M_D64 = 4 };

LG_DEV int fwd_mode(u64 q, int no_d64) {
    return q < (3ull << 44) ? (no_d64 ? M_F64 : M_D64) : (q < (1ull << 56) ? M_FREE : (q < (1ull << 61) ? M_LAZY : M_LITERAL));
}
LG_DEV int inv_mode(u64 q, int no_d64) {
    return (q < (3ull << 44) && !no_d64) ? M_D64 : (q < (1ull << 46) ? M_FREE : (q < (1ull << 61) ? M_LAZY : M_LITERAL));
}

struct TwConst {
    u64 q, qinv, twoq, fourq, nq;
    const u64* tw;   // literal: nttPsi / nttPsiInv (Montgomery).  fast: the plain-domain table (M_D64: as doubles)
    const u64* tws;  // fast: Shoup constants (M_F64, M_D64: the bits of the double table psi_wd / psi_inv_wd)
    double qd, qinvd, q34;  // M_D64: q, RD(1/q) and 34*q as doubles
};

// 256-bit global access (sm_100: LDG.E.256 / STG.E.256); p must be 32-byte aligned
LG_DEV void ld256(u64 (&v)[4], const u64* p) {
    asm volatile("ld.global.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
}
LG_DEV void ld256_nc(u64 (&v)[4], const u64* p) {
    asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(v[0]), "=l"(v[1]), "=l"(v[2]), "=l"(v[3]) : "l"(p));
}
LG_DEV void prefetch_l1(const u64* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
LG_DEV void prefetch_l2(const u64* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
LG_DEV void st256(u64* p, u64 a, u64 b, u64 c, u64 d) {
    asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}

// NG consecutive twiddles; VEC = the run is NG*8-byte aligned
template <bool VEC, int NG>
LG_DEV void load_tw(u64 (&w)[8], const u64* __restrict__ t, u32 base) {
    if (VEC && NG >= 4) {
#pragma unroll
        for (int g = 0; g < NG; g += 4) {
            u64 v[4];
            ld256_nc(v, t + base + g);
            w[g] = v[0];
            w[g + 1] = v[1];
            w[g + 2] = v[2];
            w[g + 3] = v[3];
        }
    } else if (VEC && NG == 2) {
        const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(t + base));
        w[0] = v.x;
        w[1] = v.y;
    } else {
#pragma unroll
        for (int g = 0; g < NG; ++g) w[g] = __ldg(t + base + g);
    }
}

// ---- register blocks --------------------------------------------------------
// x[r] holds the coefficient at global index j0 + r*2^s (bits [s,s+4) of j0 are
// zero); twbase = (N + j0) >> s.  Stage u pairs r and r + 2^u (stride 2^(s+u)).
template <int U, bool VEC, int MODE>
LG_DEV void fwd_stage(u64 (&x)[16], const TwConst& c, u32 twbase) {
    constexpr int NG = 16 >> (U + 1);
    const u32 base = twbase >> (U + 1);
    u64 w[8], ws[8];
    load_tw<VEC, NG>(w, c.tw, base);
    if (MODE != M_LITERAL) load_tw<VEC, NG>(ws, c.tws, base);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const double wd = __longlong_as_double((long long)ws[g]);
        const double cw = (MODE == M_F64) ? shoup_cw(wd) : 0.0;
#pragma unroll
        for (int k = 0; k < (1 << U); ++k) {
            const int r = (g << (U + 1)) + k;
            if (MODE == M_LITERAL)
                butterfly_fwd(x[r], x[r + (1 << U)], w[g], c.q, c.qinv, c.twoq);
            else if (MODE == M_FREE)
                butterfly_fwd_free(x[r], x[r + (1 << U)], w[g], ws[g], c.nq, c.fourq);
            else if (MODE == M_F64)
                butterfly_fwd_f64(x[r], x[r + (1 << U)], w[g], wd, cw, c.nq, c.fourq);
            else if (MODE == M_D64)
                butterfly_fwd_d64(x[r], x[r + (1 << U)], w[g], ws[g], c.qd);
            else
                butterfly_fwd_8q(x[r], x[r + (1 << U)], w[g], ws[g], c.nq, c.fourq);
        }
    }
}
// stages UHI, UHI-1, ..., 0
template <int UHI, bool VEC, int MODE>
LG_DEV void fwd_stages(u64 (&x)[16], const TwConst& c, u32 twbase) {
    if (UHI >= 3) fwd_stage<3, VEC, MODE>(x, c, twbase);
    if (UHI >= 2) fwd_stage<2, VEC, MODE>(x, c, twbase);
    if (UHI >= 1) fwd_stage<1, VEC, MODE>(x, c, twbase);
    fwd_stage<0, VEC, MODE>(x, c, twbase);
}

// `stage` = index (0 = first stage of the inverse transform) of this block's u = 0 stage
template <int U, bool VEC, int MODE>
LG_DEV void inv_stage(u64 (&x)[16], const TwConst& c, u32 twbase, u32 stage) {
    constexpr int NG = 16 >> (U + 1);
    const u32 base = twbase >> (U + 1);
    u64 w[8], ws[8];
    load_tw<VEC, NG>(w, c.tw, base);
    if (MODE != M_LITERAL) load_tw<VEC, NG>(ws, c.tws, base);
    const u64 m = c.q << (stage + U + 1);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
#pragma unroll
        for (int k = 0; k < (1 << U); ++k) {
            const int r = (g << (U + 1)) + k;
            if (MODE == M_LITERAL)
                butterfly_inv(x[r], x[r + (1 << U)], w[g], c.q, c.qinv, c.twoq);
            else if (MODE == M_FREE)
                butterfly_inv_free(x[r], x[r + (1 << U)], w[g], ws[g], c.nq, m);
            else if (MODE == M_D64)
                butterfly_inv_d64(x[r], x[r + (1 << U)], w[g], ws[g], c.qd);
            else
                butterfly_inv_4q(x[r], x[r + (1 << U)], w[g], ws[g], c.nq, c.fourq);
        }
    }
}
// stages 0, 1, ..., UHI
template <int UHI, bool VEC, int MODE>
LG_DEV void inv_stages(u64 (&x)[16], const TwConst& c, u32 twbase, u32 stage) {
    inv_stage<0, VEC, MODE>(x, c, twbase, stage);
    if (UHI >= 1) inv_stage<1, VEC, MODE>(x, c, twbase, stage);
    if (UHI >= 2) inv_stage<2, VEC, MODE>(x, c, twbase, stage);
    if (UHI >= 3) inv_stage<3, VEC, MODE>(x, c, twbase, stage);
}

// ---- stages with twiddles staged in shared memory ------------------------------
// The 15 twiddles a register block needs (1 + 2 + 4 + 8 for stages u = 3..0) sit in heap order:
// stage u uses slots [2^(3-u) - 1, 2^(4-u) - 1); w[g] = wp[(slot + g) * STRIDE], ws likewise from wsp.
// Twiddles staged in shared memory sit as (w, ws) pairs -- pair of heap slot k at wp[2*k*STRIDE], 16-byte aligned --
// so one 128-bit load brings both words of a butterfly group.
template <int U, int STRIDE, int MODE>
LG_DEV void fwd_stage_sm(u64 (&x)[16], const TwConst& c, const u64* wp) {
    constexpr int NG = 16 >> (U + 1);
    constexpr int S0 = (8 >> U) - 1;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const ulonglong2 pr = *reinterpret_cast<const ulonglong2*>(wp + 2 * (S0 + g) * STRIDE);
        const u64 w = pr.x;
        const u64 ws = (MODE != M_LITERAL) ? pr.y : 0ull;
        const double wd = __longlong_as_double((long long)ws);
        const double cw = (MODE == M_F64) ? shoup_cw(wd) : 0.0;
#pragma unroll
        for (int k = 0; k < (1 << U); ++k) {
            const int r = (g << (U + 1)) + k;
            if (MODE == M_LITERAL)
                butterfly_fwd(x[r], x[r + (1 << U)], w, c.q, c.qinv, c.twoq);
            else if (MODE == M_FREE)
                butterfly_fwd_free(x[r], x[r + (1 << U)], w, ws, c.nq, c.fourq);
            else if (MODE == M_F64)
                butterfly_fwd_f64(x[r], x[r + (1 << U)], w, wd, cw, c.nq, c.fourq);
            else if (MODE == M_D64)
                butterfly_fwd_d64(x[r], x[r + (1 << U)], w, ws, c.qd);
            else
                butterfly_fwd_8q(x[r], x[r + (1 << U)], w, ws, c.nq, c.fourq);
        }
    }
}
// stages UHI, ..., 0
template <int UHI, int STRIDE, int MODE>
LG_DEV void fwd_stages_sm(u64 (&x)[16], const TwConst& c, const u64* wp) {
    if (UHI >= 3) fwd_stage_sm<3, STRIDE, MODE>(x, c, wp);
    if (UHI >= 2) fwd_stage_sm<2, STRIDE, MODE>(x, c, wp);
    if (UHI >= 1) fwd_stage_sm<1, STRIDE, MODE>(x, c, wp);
    fwd_stage_sm<0, STRIDE, MODE>(x, c, wp);
}
template <int U, int STRIDE, int MODE>
LG_DEV void inv_stage_sm(u64 (&x)[16], const TwConst& c, const u64* wp, u32 stage) {
    constexpr int NG = 16 >> (U + 1);
    constexpr int S0 = (8 >> U) - 1;
    const u64 m = c.q << (stage + U + 1);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const ulonglong2 pr = *reinterpret_cast<const ulonglong2*>(wp + 2 * (S0 + g) * STRIDE);
        const u64 w = pr.x;
        const u64 ws = (MODE != M_LITERAL) ? pr.y : 0ull;
#pragma unroll
        for (int k = 0; k < (1 << U); ++k) {
            const int r = (g << (U + 1)) + k;
            if (MODE == M_LITERAL)
                butterfly_inv(x[r], x[r + (1 << U)], w, c.q, c.qinv, c.twoq);
            else if (MODE == M_FREE)
                butterfly_inv_free(x[r], x[r + (1 << U)], w, ws, c.nq, m);
            else if (MODE == M_D64)
                butterfly_inv_d64(x[r], x[r + (1 << U)], w, ws, c.qd);
            else
                butterfly_inv_4q(x[r], x[r + (1 << U)], w, ws, c.nq, c.fourq);
        }
    }
}
// stages 0, ..., UHI
template <int UHI, int STRIDE, int MODE>
LG_DEV void inv_stages_sm(u64 (&x)[16], const TwConst& c, const u64* wp, u32 stage) {
    inv_stage_sm<0, STRIDE, MODE>(x, c, wp, stage);
    if (UHI >= 1) inv_stage_sm<1, STRIDE, MODE>(x, c, wp, stage);
    if (UHI >= 2) inv_stage_sm<2, STRIDE, MODE>(x, c, wp, stage);
    if (UHI >= 3) inv_stage_sm<3, STRIDE, MODE>(x, c, wp, stage);
}

// Strided phase twiddles: group 0 = the subtree under node 1 (the block on the top four stages, the same
// for every thread), group 1+g = the subtree under node 2^(L-4) + g (the block below it).  Heap slot k of
// the subtree under node m is table entry (m << lvl) + (k + 1 - 2^lvl), lvl = floor(log2(k + 1)).
template <int L, bool LIT>
LG_DEV void fill_strided_tw(u64* tws_sm, const TwConst& c) {
    constexpr int G = 1 << (L - 4);
    for (int e = threadIdx.x; e < 15 * (G + 1); e += 256) {
        const int grp = e / 15, k = e - grp * 15;
        const u32 lvl = 31 - __clz(k + 1);
        const u32 node = grp == 0 ? 1u : (u32)(G + grp - 1);
        const u32 idx = (node << lvl) + (k + 1 - (1u << lvl));
        tws_sm[grp * 32 + 2 * k] = __ldg(c.tw + idx);
        if (!LIT) tws_sm[grp * 32 + 2 * k + 1] = __ldg(c.tws + idx);
    }
}

// CTAs are dispatched in blockIdx order (x fastest, z slowest).  A launch that reads what the previous launch wrote can
// walk its grid backwards (NttArgs::rev): the data written last -- still in L2 -- is then read first.
// NttArgs::tfast (strided phases): the tile index is the fastest grid dimension instead of the batch index, so the CTAs
// in flight together read and write ADJACENT 128-byte columns of the same rows (whole 2 KiB rows of the limb): DRAM page
// locality for a launch that is HBM-bound.
LG_DEV int cta_x(const NttArgs& a) {  // batch index
    if (a.tfast) return a.rev ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
    return a.rev ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
}
LG_DEV int cta_y(const NttArgs& a) {  // tile index
    if (a.tfast) return a.rev ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
    return a.rev ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
}
LG_DEV int cta_z(const NttArgs& a) { return a.rev ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z; }

struct LimbSetup {
    LimbConst c;
    const u64* in;
    u64* out;
    int tl;
    bool skip;
};

LG_DEV LimbSetup setup_limb(const NttArgs& a) {
    LimbSetup s;
    const int j = cta_z(a), b = cta_x(a);
    s.tl = a.map(j);
    if (a.skip_alpha > 0) {
        const int dg = b / a.skip_div;
        s.skip = (s.tl < a.skip_nl) && (s.tl >= dg * a.skip_alpha) && (s.tl < (dg + 1) * a.skip_alpha);
    } else {
        s.skip = (j >= a.skip0 && j < a.skip1);
    }
    s.c = load_limb_const(a.T, s.tl);
    s.in = a.in + (size_t)b * a.in_bstride + (a.bcast.enabled ? 0 : (size_t)j * (a.in_ls ? a.in_ls : a.T.N));
    s.out = a.out + (size_t)b * a.out_bstride + (size_t)j * (a.out_ls ? a.out_ls : a.T.N);
    return s;
}

template <bool FWD, int MODE>
LG_DEV TwConst tw_const(const RingTables& T, const LimbConst& lc, int tl) {
    TwConst c;
    c.q = lc.q;
    c.qinv = lc.qinv;
    c.twoq = 2 * lc.q;
    c.fourq = 4 * lc.q;
    c.nq = 0ull - lc.q;
    const size_t off = (size_t)tl * T.N;
    if (MODE == M_LITERAL) {
        c.tw = (FWD ? T.psi : T.psi_inv) + off;
        c.tws = nullptr;
    } else if (MODE == M_F64) {
        c.tw = T.psi_w + off;
        c.tws = T.psi_wd + off;
    } else if (MODE == M_D64) {
        c.tw = (FWD ? T.psi_wf : T.psi_inv_wf) + off;
        c.tws = (FWD ? T.psi_wd : T.psi_inv_wd) + off;
        c.qd = __ull2double_rn(lc.q);  // exact: q < 2^46
        c.qinvd = __ddiv_rd(1.0, c.qd);
        c.q34 = 34.0 * c.qd;           // exact: an integer below 2^52
    } else {
        c.tw = (FWD ? T.psi_w : T.psi_inv_w) + off;
        c.tws = (FWD ? T.psi_ws : T.psi_inv_ws) + off;
    }
    return c;
}

// M_D64 helpers on the bit patterns held in x[]
// signed lazy double (|v| < 34q) -> canonical integer
LG_DEV u64 d64_canon(u64 xb, const TwConst& c) {
    const double r = d64_red(__dadd_rn(bits2d(xb), c.q34), c.qinvd, c.qd);  // v + 34q >= 0: r in [0, 2q)
    return cred(d_to_u52(r, 4503599627370496.0), c.q);
}
LG_DEV void d64_reduce_all(u64 (&x)[16], const TwConst& c) {  // -> [-q, 2q)
#pragma unroll
    for (int r = 0; r < 16; ++r) x[r] = d2bits(d64_red(bits2d(x[r]), c.qinvd, c.qd));
}

// ---- forward, strided phase: stages 1..L -----------------------------------
// (rows are 2^8 words apart: the contiguous phase always takes the low 8 stages)
template <int L, int MODE>
LG_DEV void fwd_strided_body(const NttArgs& a, const LimbSetup& s, u64* sm, u64* tws_sm) {
    constexpr int G = 1 << (L - 4);  // threads per column
    constexpr int W = 256 / G;       // columns per CTA
    constexpr int N2 = L - 4;        // stages of the second register block
    const TwConst c = tw_const<true, MODE>(a.T, s.c, s.tl);
    const int t = threadIdx.x, col = t % W, g = t / W;
    const u32 colg = cta_y(a) * W + col;
    u64 x[16];
    const u64* in = s.in + colg + g * 256;
#pragma unroll
    for (int r = 0; r < 16; ++r) x[r] = in[r * G * 256];
    if (a.bcast.enabled && a.bcast.add != nullptr) {  // ring_scaling.go:99-103: + (q_j - pHalf mod q_j), unreduced
        const u64 add = __ldg(a.bcast.add + s.tl);
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] += add;
    }
    fill_strided_tw<L, MODE == M_LITERAL>(tws_sm, c);
    if (MODE != M_LITERAL) {
        // headroom of the lazy butterflies (canonical inputs never take the slow branch): M_FREE keeps
        // everything below 2^63, M_F64 below 2^50, M_LAZY below 2^(bits(q)+2) <= 8q
        // (M_D64: below 2^49, so that 16 stages stay below 2^49 + 32q < 2^51)
        const u32 sh = MODE == M_FREE ? 63u : (MODE == M_F64 ? 50u : (MODE == M_D64 ? 49u : 66u - (u32)__clzll((long long)c.q)));
        u64 o = 0;
#pragma unroll
        for (int r = 0; r < 16; ++r) o |= x[r];
        if (o >> sh) {
#pragma unroll
            for (int r = 0; r < 16; ++r)
                if (x[r] >> sh) x[r] = bred_add(x[r], c.q, s.c.u0);
        }
    }
    if (MODE == M_D64) {  // integers -> doubles; the phase leaves raw doubles for the contiguous phase
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] = d2bits(u52_to_d(x[r], 4503599627370496.0));
    }
    __syncthreads();
    fwd_stages_sm<3, 1, MODE>(x, c, tws_sm);
    if (N2 > 0) {
#pragma unroll
        for (int r = 0; r < 16; ++r) sm[(g + r * G) * W + col] = x[r];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] = sm[(16 * g + r) * W + col];
        const u64* twg = tws_sm + (1 + g) * 32;
        fwd_stages_sm<(N2 > 0 ? N2 - 1 : 0), 1, MODE>(x, c, twg);
        u64* out = s.out + colg + g * 16 * 256;
#pragma unroll
        for (int r = 0; r < 16; ++r) out[r * 256] = x[r];
    } else {
        u64* out = s.out + colg + g * 256;
#pragma unroll
        for (int r = 0; r < 16; ++r) out[r * G * 256] = x[r];
    }
}

#ifndef STRIDED_MINB
#define STRIDED_MINB 4
#endif
template <int L, bool LITERAL>
__global__ void __launch_bounds__(256, STRIDED_MINB) ntt_fwd_strided(const NttArgs a) {
    __shared__ u64 sm[(L - 4) > 0 ? 4096 : 1];
    __shared__ __align__(16) u64 tws_sm[32 * ((1 << (L - 4)) + 1)];
    const LimbSetup s = setup_limb(a);
    if (s.skip) return;
    const int mode = LITERAL ? M_LITERAL : fwd_mode(s.c.q, a.no_d64);
    if (mode == M_D64)
        fwd_strided_body<L, M_D64>(a, s, sm, tws_sm);
    else if (mode == M_F64)
        fwd_strided_body<L, M_F64>(a, s, sm, tws_sm);
    else if (mode == M_FREE)
        fwd_strided_body<L, M_FREE>(a, s, sm, tws_sm);
    else if (mode == M_LAZY)
        fwd_strided_body<L, M_LAZY>(a, s, sm, tws_sm);
    else
        fwd_strided_body<L, M_LITERAL>(a, s, sm, tws_sm);
}

// ---- contiguous phase geometry ------------------------------------------------
// 128-thread CTA = 4 warps = 8 segments of 256 words (tile of 2048 words); the 16 threads of a segment
// exchange through shared memory warp-synchronously.
#define CONTIG_THREADS 128
#define CONTIG_TILE 2048u

// ---- key-switch digit loop fused with the contiguous phase ---------------------
// Everything the digit loop re-reads is staged once: the tile's twiddles live in shared memory for all
// digits (thread-private slots for the second register block, one copy per segment for the first), and the
// next digit's tile is fetched with cp.async into the other half of a double buffer while the current one
// is transformed; the buffer a digit was read from then serves as its exchange space.
//   shared per CTA: 2 x 16 KiB tiles + 30 KiB private twiddles + 2 KiB segment twiddles = 64 KiB (3 CTAs/SM)
#define KS_SMEM_WORDS (2 * 2048 + 30 * CONTIG_THREADS + 8 * 32)

LG_DEV void cp_async16(u64* smem_dst, const u64* gsrc) {
    const u32 d = (u32)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
LG_DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
LG_DEV void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// a warp fetches its own two segments (512 words) of a tile
LG_DEV void prefetch_warp_tile(u64* buf, const u64* __restrict__ src_tile) {
    const u32 lane = threadIdx.x & 31, wbase = (threadIdx.x >> 5) * 512u;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const u32 o = wbase + 2u * (lane + 32u * k);
        cp_async16(buf + o, src_tile + o);
    }
    cp_async_commit();
}

// Twiddles of one 2048-word tile: thread-private slots for the register block on the natural layout
// (words 16cc + r), one copy per segment for the block on the column layout (words cc + 16r).
template <int MODE>
LG_DEV void contig_fill_tw(const TwConst& c, u32 N, u32 segbase, u32 cc, u64* twp, u64* twseg) {
    u64 w[8];
    const u32 tb = N + segbase + 16 * cc;
#define LG_FILL(U, SLOT)                                                    \
    load_tw<true, (16 >> (U + 1))>(w, c.tw, tb >> (U + 1));                 \
    _Pragma("unroll") for (int g = 0; g < (16 >> (U + 1)); ++g) twp[2 * (SLOT + g) * CONTIG_THREADS] = w[g]; \
    if (MODE != M_LITERAL) {                                                \
        load_tw<true, (16 >> (U + 1))>(w, c.tws, tb >> (U + 1));            \
        _Pragma("unroll") for (int g = 0; g < (16 >> (U + 1)); ++g) twp[2 * (SLOT + g) * CONTIG_THREADS + 1] = w[g]; \
    }
    LG_FILL(3, 0)
    LG_FILL(2, 1)
    LG_FILL(1, 3)
    LG_FILL(0, 7)
#undef LG_FILL
    if (cc < 15) {  // node m = (N + segbase) >> 8 and its 15 descendants in heap order
        const u32 lvl = 31 - __clz(cc + 1);
        const u32 idx = (((N + segbase) >> 8) << lvl) + (cc + 1 - (1u << lvl));
        twseg[2 * cc] = __ldg(c.tw + idx);
        if (MODE != M_LITERAL) twseg[2 * cc + 1] = __ldg(c.tws + idx);
    }
}

// same fetch with the 16-byte chunks of every 16-word row rotated by the row index (chunk p of row r at
// r*8 + (p ^ (r & 7))), so that threads reading whole rows with 128-bit loads do not collide
LG_DEV void prefetch_warp_tile_rows(u64* buf, const u64* __restrict__ src_tile) {
    const u32 lane = threadIdx.x & 31, wbase = (threadIdx.x >> 5) * 512u;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const u32 u = lane + 32u * k, seg = u >> 7, v = u & 127u, row = v >> 3, pp = v & 7u;
        cp_async16(buf + wbase + 2u * (seg * 128u + row * 8u + (pp ^ (row & 7u))), src_tile + wbase + 2u * u);
    }
    cp_async_commit();
}

// ---- contiguous phase of the plain transforms, pipelined over batch entries ------------------------
// One CTA takes the same tile of up to `bpc` batch entries: the twiddles are staged once, entry i+1 is
// fetched with cp.async while entry i is transformed (same structure as the key-switch digit loop).
// Single tile buffer (48 KiB of shared memory per CTA, 4 CTAs/SM): the fetch of entry i+1 is issued as soon
// as entry i has left the buffer for good (after the exchange), and lands during the second register block.
// Build-time variants of the fused digit loop (A/B numbers in profiles/README.md): resident CTAs per SM, and where the
// 32 key words of a digit are loaded -- 0: in the multiply-accumulate, 1: before the last register block (needs
// KS_MINB 2: 254 registers), 2: at the top of the iteration (spills).  The shipped build is 3 / 0.
#ifndef KS_MINB
#define KS_MINB 3
#endif
#ifndef KS_KEYPREFETCH
#define KS_KEYPREFETCH 0
#endif
#define PIPE_SMEM_WORDS (2048 + 30 * CONTIG_THREADS + 8 * 32)
template <bool FWD, int MODE, bool TAIL>
LG_DEV void contig_pipe_body(const NttArgs& a, const LimbConst& lc, int tl, int b0, int nb, u64* smem) {
    const u32 N = a.T.N;
    const int j = cta_z(a);
    const TwConst c = tw_const<FWD, MODE>(a.T, lc, tl);
    const u32 t = threadIdx.x, sg = t >> 4, cc = t & 15;
    const u32 tile0 = cta_y(a) * CONTIG_TILE;
    const u32 segbase = tile0 + sg * 256u;
    const u32 e0 = segbase + 16 * cc, j0 = segbase + cc;
    u64* const tilebuf = smem;
    u64* const buf = smem + sg * 256;
    u64* const twp = smem + 2048 + 2 * t;  // private (w, ws) pairs: slot k at twp[2*k*CONTIG_THREADS]
    u64* const twseg = smem + 2048 + 30 * CONTIG_THREADS + sg * 32;
    const u64* src = a.in + (size_t)b0 * a.in_bstride + (size_t)j * (a.in_ls ? a.in_ls : N) + tile0;
    u64* dst = a.out + (size_t)b0 * a.out_bstride + (size_t)j * (a.out_ls ? a.out_ls : N);
    if (FWD)
        prefetch_warp_tile(tilebuf, src);
    else
        prefetch_warp_tile_rows(tilebuf, src);
    contig_fill_tw<MODE>(c, N, segbase, cc, twp, twseg);
#pragma unroll 1
    for (int i = 0; i < nb; ++i, dst += a.out_bstride) {
        u64 x[16];
        cp_async_wait_all();
        __syncwarp();
        if (FWD) {
#pragma unroll
            for (int r = 0; r < 16; ++r) x[r] = buf[cc + 16 * r];
            fwd_stages_sm<3, 1, MODE>(x, c, twseg);
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 16; ++r) buf[16 * r + (cc ^ r)] = x[r];
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 16; ++r) x[r] = buf[16 * cc + (r ^ cc)];
            __syncwarp();
            if (i + 1 < nb) prefetch_warp_tile(tilebuf, src + (size_t)(i + 1) * a.in_bstride);
            if (TAIL && a.pf) {  // the thread's 16 words of a[] (and of out[] when it is added to) = one 128-byte line each
                const int bi = a.batch0 + b0 + i, set = bi >= a.tail.split ? 1 : 0;
                const size_t bb = (size_t)(bi - (set ? a.tail.split : 0));
                const u64* ta = a.tail.a[set] + bb * a.tail.a_bs[set] + (size_t)j * (a.tail.a_ls ? a.tail.a_ls : N) + e0;
                if (a.pf == 2) prefetch_l2(ta); else prefetch_l1(ta);
                if (a.tail.add[set]) {
                    const u64* to = a.tail.out[set] + bb * a.tail.out_bs[set] + (size_t)j * (a.tail.out_ls ? a.tail.out_ls : N) + e0;
                    if (a.pf == 2) prefetch_l2(to); else prefetch_l1(to);
                }
            }
            fwd_stages_sm<3, CONTIG_THREADS, MODE>(x, c, twp);
            if (TAIL) {
                // the caller's (x - NTT(y)) * s_j tail (+ add) on the canonical transform, straight from the registers
                const int bi = a.batch0 + b0 + i, set = bi >= a.tail.split ? 1 : 0;
                const size_t bb = (size_t)(bi - (set ? a.tail.split : 0));
                const u64* ta = a.tail.a[set] + bb * a.tail.a_bs[set] + (size_t)j * (a.tail.a_ls ? a.tail.a_ls : N) + e0;
                u64* to = a.tail.out[set] + bb * a.tail.out_bs[set] + (size_t)j * (a.tail.out_ls ? a.tail.out_ls : N) + e0;
                const u64 sj = __ldg(a.tail.s + tl);
                const bool add = a.tail.add[set] != 0;
                const bool canon = a.tail.a_canon != 0;
                const u64 kq = ((lc.u0 >> 12) + 1) * c.q;  // the first multiple of q above 2^52 (u0 = floor(2^64/q))
#pragma unroll
                for (int h = 0; h < 4; ++h) {
                    u64 va[4], r[4];
                    ld256(va, ta + 4 * h);
                    if (MODE == M_D64 && canon) {
                        // va < q and |x| < 34q: va + (34q - x) is positive, below 2^52 and congruent
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            r[e] = mred(va[e] + d_to_u52(-bits2d(x[4 * h + e]), 4503599627370496.0 + c.q34), sj, c.q, c.qinv);
                    } else if (MODE == M_D64) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) r[e] = mred(va[e] + (c.q - d64_canon(x[4 * h + e], c)), sj, c.q, c.qinv);
                    } else if (MODE == M_F64 && canon) {
                        // va < q and x < 2^52 (FP64 transform): va + (kq - x) is positive, below 2^53 and congruent, so the
                        // one canonical word MRed + CRed returns is the reference's
#pragma unroll
                        for (int e = 0; e < 4; ++e) r[e] = mred(va[e] + (kq - x[4 * h + e]), sj, c.q, c.qinv);
                    } else {
#pragma unroll
                        for (int e = 0; e < 4; ++e)
                            r[e] = mred(va[e] + (c.q - bred_add(x[4 * h + e], c.q, lc.u0)), sj, c.q, c.qinv);
                    }
                    if (add) {
                        u64 vo[4];
                        ld256(vo, to + 4 * h);
#pragma unroll
                        for (int e = 0; e < 4; ++e) r[e] = cred(vo[e] + r[e], c.q);
                    }
                    st256(to + 4 * h, r[0], r[1], r[2], r[3]);
                }
                continue;
            }
            // ring/ntt.go:83-85
            if (MODE == M_D64) {
#pragma unroll
                for (int h = 0; h < 4; ++h)
                    st256(dst + e0 + 4 * h, d64_canon(x[4 * h], c), d64_canon(x[4 * h + 1], c), d64_canon(x[4 * h + 2], c),
                          d64_canon(x[4 * h + 3], c));
                continue;
            }
#pragma unroll
            for (int h = 0; h < 4; ++h)
                st256(dst + e0 + 4 * h, bred_add(x[4 * h], c.q, lc.u0), bred_add(x[4 * h + 1], c.q, lc.u0),
                      bred_add(x[4 * h + 2], c.q, lc.u0), bred_add(x[4 * h + 3], c.q, lc.u0));
        } else {
            const ulonglong2* row = reinterpret_cast<const ulonglong2*>(buf + 16 * cc);
#pragma unroll
            for (int pp = 0; pp < 8; ++pp) {
                const ulonglong2 v = row[pp ^ (cc & 7)];
                x[2 * pp] = v.x;
                x[2 * pp + 1] = v.y;
            }
            if (MODE == M_D64) {  // in-range integers (<= 2q) -> doubles
#pragma unroll
                for (int r = 0; r < 16; ++r) x[r] = d2bits(u52_to_d(x[r], 4503599627370496.0));
            }
            inv_stages_sm<3, CONTIG_THREADS, MODE>(x, c, twp, 0u);
            if (MODE == M_D64) d64_reduce_all(x, c);  // sums of 16 inputs -> [-q, 2q) before the next four stages
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 16; ++r) buf[16 * cc + (r ^ cc)] = x[r];
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 16; ++r) x[r] = buf[16 * r + (cc ^ r)];
            __syncwarp();
            if (i + 1 < nb) prefetch_warp_tile_rows(tilebuf, src + (size_t)(i + 1) * a.in_bstride);
            inv_stages_sm<3, 1, MODE>(x, c, twseg, 4u);
#pragma unroll
            for (int r = 0; r < 16; ++r) dst[j0 + 16 * r] = x[r];
        }
    }
}

// TAIL: the store is the caller's (x - NTT(y)) * s_j epilogue (NttTail); a separate instantiation, so the plain
// transform keeps its register budget
template <bool FWD, bool LITERAL, bool TAIL = false>
__global__ void __launch_bounds__(CONTIG_THREADS, 4) ntt_contig_pipe(const NttArgs a, int batch, int bpc) {
    extern __shared__ __align__(16) u64 ks_smem[];
    const int j = cta_z(a);
    const int b0 = cta_x(a) * bpc, nb = (batch - b0) < bpc ? (batch - b0) : bpc;
    const int tl = a.map(j);
    if (a.skip_alpha > 0) {  // digit-batched launch: bpc divides skip_div, so a group never straddles two digits
        const int dg = b0 / a.skip_div;
        if (tl < a.skip_nl && tl >= dg * a.skip_alpha && tl < (dg + 1) * a.skip_alpha) return;
    } else if (j >= a.skip0 && j < a.skip1) {
        return;
    }
    const LimbConst lc = load_limb_const(a.T, tl);
    int mode;
    if (FWD) {
        mode = LITERAL ? M_LITERAL : fwd_mode(lc.q, a.no_d64);
    } else {
        bool flagged = LITERAL;  // one flagged entry makes the whole group literal (always exact)
        if (a.flags != nullptr) flagged |= a.flags[j] != 0;
        mode = flagged ? M_LITERAL : inv_mode(lc.q, a.no_d64);
    }
    if (mode == M_D64)
        contig_pipe_body<FWD, M_D64, TAIL>(a, lc, tl, b0, nb, ks_smem);
    else if (mode == M_F64)
        contig_pipe_body<FWD, FWD ? M_F64 : M_FREE, TAIL>(a, lc, tl, b0, nb, ks_smem);
    else if (mode == M_FREE)
        contig_pipe_body<FWD, M_FREE, TAIL>(a, lc, tl, b0, nb, ks_smem);
    else if (mode == M_LAZY)
        contig_pipe_body<FWD, M_LAZY, TAIL>(a, lc, tl, b0, nb, ks_smem);
    else
        contig_pipe_body<FWD, M_LITERAL, TAIL>(a, lc, tl, b0, nb, ks_smem);
}

// ---- key-switch accumulators ------------------------------------------------------------------------------
// The reference adds MRed(evk, d) terms lazily and ends canonical (ckks/evaluator.go:1515-1552), i.e. it returns
// (sum_i evk_i * d_i) * 2^-64 mod q.  Montgomery reduction is linear, so the same word comes out of ONE reduction
// of the exact integer sum of the products.  For the FP64-butterfly limbs (q < 3*2^44, transform values below 2^52)
// the digit value is first brought below 3q with one round-down DFMA (quotient by q as in shoup_f64, w = 1), the
// product then stays below 3q^2 and beta of them fit a 96-bit accumulator: 3 wide + 1 narrow multiply and two adds
// per term instead of the 11 multiplies of a Montgomery reduction per term.  Key words may be any 64-bit value (the
// reference's MRed is total): the kernel ORs their high halves on the way and a CTA that saw a word of more bits
// than q repeats its tile on the 64-bit path, which is exact for every word.
// ACC_EXACT is the always-valid form: the digit value is made canonical first (as the reference's c2QiQ / c2QiP are) and
// every term is a full MRed + CRed, exact for ANY 64-bit key word.  The two lazy forms assume key words of at most
// bits(q) bits; they watch the high halves of the key words and the CTA repeats its tile with ACC_EXACT otherwise.
// ACC_FP (M_D64 limbs only): the multiply-accumulate on the FP64 pipe as well -- the key limb out of Montgomery form as
// doubles (lg_launch_swk_prepare), every term k*x reduced to [-q, 2q) by d64_mul (the digit value is already a double
// below 34q: no conversion), the accumulators two doubles per coefficient (64 registers instead of 96).  Sum of beta
// terms below 2q: exact.  8 instructions per term instead of 12.5; limbs with a non-canonical key word stay integer.
enum { ACC_EXACT = 0, ACC_LAZY64 = 1, ACC_WIDE96 = 2, ACC_FP = 3 };

// (a2 : A) += k * x, caller guarantees the running sum stays below 2^96
LG_DEV void mac96(u64& A, u32& a2, u64 k, u64 x) {
    const u32 k0 = (u32)k, k1 = (u32)(k >> 32), x0 = (u32)x, x1 = (u32)(x >> 32);
    u64 h = mul_wide(k1, x0);
    h = mad_wide(k0, x1, h);
    const u32 hl = (u32)h, hh = mad_lo32(k1, x1, (u32)(h >> 32));
    asm("{\n\t.reg .u32 l, h;\n\tmov.b64 {l, h}, %0;\n\t"
        "mad.lo.cc.u32 l, %2, %3, l;\n\tmadc.hi.cc.u32 h, %2, %3, h;\n\taddc.u32 %1, %1, 0;\n\t"
        "add.cc.u32 h, h, %4;\n\taddc.u32 %1, %1, %5;\n\tmov.b64 %0, {l, h};\n\t}"
        : "+l"(A), "+r"(a2)
        : "r"(k0), "r"(x0), "r"(hl), "r"(hh));
}
// x < 2^52 -> x - Qh*q in [0,3q), Qh = floor(x/q) - {0,1,2} from one round-down DFMA (see shoup_f64; Qh = -1 for
// tiny x arrives in two's complement and adds q)
LG_DEV u64 reduce_f64(u64 x, double qd1, double cq1, u64 nq) {
    const u32 b0 = (u32)x, b1 = (u32)(x >> 32);
    const double qd = __fma_rd(__hiloint2double((int)(b1 | 0x43300000u), (int)b0), qd1, cq1);
    const u32 h0 = (u32)__double2loint(qd), h1 = (u32)__double2hiint(qd) - 0x43300000u;
    const u32 n0 = (u32)nq, n1 = (u32)(nq >> 32);
    const u64 t = mad_wide(h0, n0, x);
    u32 th = (u32)(t >> 32);
    th = mad_lo32(h0, n1, th);
    th = mad_lo32(h1, n0, th);
    return ((u64)th << 32) | (u32)t;
}
// Montgomery reduction of the 96-bit sum (a2 < 2^32 < q): canonical
LG_DEV u64 mred96(u64 A, u32 a2, u64 q, u64 qinv) {
    const u64 H = mul_hi(mul_lo(A, qinv), q);
    return cred((u64)a2 - H + q, q);
}

// the thread's 16 words of evk[i][0] and evk[i][1]
LG_DEV void ks_load_keys(u64 (&k0)[16], u64 (&k1)[16], const u64* key, size_t hs) {
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        u64 v[4], w[4];
        ld256_nc(v, key + 4 * h);
        ld256_nc(w, key + hs + 4 * h);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            k0[4 * h + e] = v[e];
            k1[4 * h + e] = w[e];
        }
    }
}

// data limb of this CTA: blockIdx.z, through the launcher's list when the limbs are split between two kernels
LG_DEV int ks_data_limb(const KsFusedArgs& a) {
    const int z = a.rev ? (int)(gridDim.z - 1 - blockIdx.z) : (int)blockIdx.z;
    return a.use_zl ? (int)a.zl[z] : z;
}

// ACC_LAZY64: beta * 2q fits 64 bits, so the products are accumulated unreduced (MRedConstant, in (0,2q)).
template <int MODE, int ACC>
LG_DEV u32 ks_fused_body(const KsFusedArgs& a, const LimbConst& lc, int tl, u64* smem) {
    const u32 N = a.T.N;
    const int j = ks_data_limb(a);
    const int b = a.rev ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
    const u64 q = lc.q, qinv = lc.qinv;
    const TwConst c = tw_const<true, MODE>(a.T, lc, tl);
    const u32 t = threadIdx.x, sg = t >> 4, cc = t & 15;
    const u32 tile0 = (a.rev ? gridDim.y - 1 - blockIdx.y : blockIdx.y) * CONTIG_TILE;  // first word of the CTA's tile within the limb
    const u32 segbase = tile0 + sg * 256u;
    const u32 e0 = segbase + 16 * cc;
    u64* const tilebuf = smem;                        // [2][2048]
    u64* const twp = smem + 2 * 2048 + 2 * t;         // private (w, ws) pairs: slot k at twp[2*k*CONTIG_THREADS]
    u64* const twseg = smem + 2 * 2048 + 30 * CONTIG_THREADS + sg * 32;  // segment pairs: slot k at twseg[2*k]

    const u64* din = a.D + (size_t)b * a.d_bs + (size_t)j * N + tile0;
    const int own_i = (tl < a.nl) ? tl / a.alpha : -1;  // the digit whose own limb this is
    if (own_i != 0) prefetch_warp_tile(tilebuf, din);

    contig_fill_tw<MODE>(c, N, segbase, cc, twp, twseg);  // twiddles of the tile, once for all digits

    // floor(2^64/q) is the high Barrett word (modular_reduction.go:97-106); rounded down it is the Shoup double of w = 1
    const double qd1 = (ACC == ACC_WIDE96 && MODE != M_D64) ? __ull2double_rd(lc.u0) * 5.421010862427522170037e-20 : 0.0;
    const double cq1 = (ACC == ACC_WIDE96 && MODE != M_D64) ? shoup_cw(qd1) : 0.0;

    const u64* key = (ACC == ACC_FP ? a.evk_f : a.evk) + (size_t)tl * N + e0;
    u64 acc0[16], acc1[16];  // ACC_FP: bit patterns of doubles (+0.0 = 0)
    u32 top0[16], top1[16], keyhi = 0;
#pragma unroll
    for (int r = 0; r < 16; ++r) {
        acc0[r] = acc1[r] = 0;
        top0[r] = top1[r] = 0;
    }
#pragma unroll 1
    for (int i = 0; i < a.beta; ++i, key += a.evk_ds) {
        u64 x[16];
#if KS_KEYPREFETCH == 1
        u64 kk0[16], kk1[16];
#endif
        u64* const buf = tilebuf + (i & 1) * 2048 + sg * 256;
        cp_async_wait_all();
        __syncwarp();
        // fetch the next digit's tile into the other buffer (all its readers passed the barrier above)
        if (i + 1 < a.beta && i + 1 != own_i)
            prefetch_warp_tile(tilebuf + ((i + 1) & 1) * 2048, din + (size_t)(i + 1) * a.d_ds);
#if KS_KEYPREFETCH == 2
        u64 kk0[16], kk1[16];
        ks_load_keys(kk0, kk1, key, a.evk_hs);
#endif
        if (a.pf == 2) {  // at the top of the iteration instead
            prefetch_l1(key);
            prefetch_l1(key + a.evk_hs);
        }
        if (i == own_i) {  // ckks/evaluator.go:1579-1584, bfv/evaluator.go:776-780
#if KS_KEYPREFETCH == 1
            ks_load_keys(kk0, kk1, key, a.evk_hs);
#endif
            const u64* cx = a.cx + (size_t)b * a.cx_bs + (size_t)j * (a.cx_ls ? a.cx_ls : N) + e0;
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                u64 v[4];
                ld256(v, cx + 4 * h);
#pragma unroll
                for (int e = 0; e < 4; ++e) x[4 * h + e] = v[e];
            }
            if (ACC == ACC_WIDE96 || ACC == ACC_EXACT || ACC == ACC_FP) {  // caller data, any 64-bit word: canonical
#pragma unroll
                for (int r = 0; r < 16; ++r) x[r] = bred_add(x[r], q, lc.u0);
            }
            if (ACC == ACC_FP) {
#pragma unroll
                for (int r = 0; r < 16; ++r) x[r] = d2bits(u52_to_d(x[r], 4503599627370496.0));
            }
        } else {
#pragma unroll
            for (int r = 0; r < 16; ++r) x[r] = buf[cc + 16 * r];
            fwd_stages_sm<3, 1, MODE>(x, c, twseg);
            __syncwarp();
            // exchange in place, XOR-swizzled (word 16r+cc at 16r + (cc^r)): conflict-free both ways
#pragma unroll
            for (int r = 0; r < 16; ++r) buf[16 * r + (cc ^ r)] = x[r];
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 16; ++r) x[r] = buf[16 * cc + (r ^ cc)];
#if KS_KEYPREFETCH == 1
            ks_load_keys(kk0, kk1, key, a.evk_hs);  // in flight during the second register block
#endif
            if (a.pf == 1) {  // the thread's 16 words of evk[i][0] and evk[i][1]: one 128-byte line each
                prefetch_l1(key);
                prefetch_l1(key + a.evk_hs);
            }
            fwd_stages_sm<3, CONTIG_THREADS, MODE>(x, c, twp);
            if (MODE == M_D64) {  // doubles (|v| < 34q) -> the integers the multiply-accumulate takes
#pragma unroll
                for (int r = 0; r < 16; ++r) {
                    if (ACC == ACC_WIDE96)  // [0, 3q)
                        x[r] = d_to_u52(d64_red(bits2d(x[r]), c.qinvd, c.qd), 4503599627370496.0 + c.qd);
                    else if (ACC == ACC_LAZY64)  // (0, 68q), congruent
                        x[r] = d_to_u52(bits2d(x[r]), 4503599627370496.0 + c.q34);
                    else if (ACC == ACC_EXACT)
                        x[r] = d64_canon(x[r], c);
                }
            } else if (ACC == ACC_WIDE96) {
#pragma unroll
                for (int r = 0; r < 16; ++r) x[r] = reduce_f64(x[r], qd1, cq1, c.nq);
            } else if (ACC == ACC_EXACT) {
#pragma unroll
                for (int r = 0; r < 16; ++r) x[r] = bred_add(x[r], q, lc.u0);
            }
        }
#if KS_KEYPREFETCH == 0
        u64 kk0[16], kk1[16];
        ks_load_keys(kk0, kk1, key, a.evk_hs);
#endif
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            if (ACC == ACC_FP) {
                const double xd = bits2d(x[r]), k0 = bits2d(kk0[r]), k1 = bits2d(kk1[r]);
                acc0[r] = d2bits(__dadd_rn(bits2d(acc0[r]), d64_mul(k0, __dmul_rd(k0, c.qinvd), xd, c.qd)));
                acc1[r] = d2bits(__dadd_rn(bits2d(acc1[r]), d64_mul(k1, __dmul_rd(k1, c.qinvd), xd, c.qd)));
            } else if (ACC == ACC_WIDE96) {
                keyhi |= (u32)(kk0[r] >> 32) | (u32)(kk1[r] >> 32);
                mac96(acc0[r], top0[r], kk0[r], x[r]);
                mac96(acc1[r], top1[r], kk1[r], x[r]);
            } else if (ACC == ACC_LAZY64) {
                keyhi |= (u32)(kk0[r] >> 32) | (u32)(kk1[r] >> 32);
                acc0[r] += mred_constant(kk0[r], x[r], q, qinv);
                acc1[r] += mred_constant(kk1[r], x[r], q, qinv);
            } else {
                acc0[r] = cred(acc0[r] + mred(kk0[r], x[r], q, qinv), q);
                acc1[r] = cred(acc1[r] + mred(kk1[r], x[r], q, qinv), q);
            }
        }
    }
    u64* o0 = a.acc0 + (size_t)b * a.acc_bs + (size_t)j * N + e0;
    u64* o1 = a.acc1 + (size_t)b * a.acc_bs + (size_t)j * N + e0;
    if (ACC == ACC_FP) {  // |sum| < 18q
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            acc0[r] = d64_canon(acc0[r], c);
            acc1[r] = d64_canon(acc1[r], c);
        }
    } else if (ACC == ACC_WIDE96) {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            acc0[r] = mred96(acc0[r], top0[r], q, qinv);
            acc1[r] = mred96(acc1[r], top1[r], q, qinv);
        }
    } else {
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            acc0[r] = bred_add(acc0[r], q, lc.u0);
            acc1[r] = bred_add(acc1[r], q, lc.u0);
        }
    }
#pragma unroll
    for (int h = 0; h < 4; ++h) {
        st256(o0 + 4 * h, acc0[4 * h], acc0[4 * h + 1], acc0[4 * h + 2], acc0[4 * h + 3]);
        st256(o1 + 4 * h, acc1[4 * h], acc1[4 * h + 1], acc1[4 * h + 2], acc1[4 * h + 3]);
    }
    return keyhi;
}

template <bool LITERAL>
__global__ void __launch_bounds__(CONTIG_THREADS, KS_MINB) ks_fused_kernel(const KsFusedArgs a) {
    extern __shared__ __align__(16) u64 ks_smem[];
    const int tl = a.map(ks_data_limb(a));
    const LimbConst lc = load_limb_const(a.T, tl);
    const int mode = LITERAL ? M_LITERAL : fwd_mode(lc.q, a.no_d64);
    // beta lazy terms below 2q fit 64 bits (a term is below 2q when the key word has at most bits(q) bits: its product
    // with a transform value below 2^63 (M_FREE), 8q (M_LAZY) or 2^52 (M_F64) then has a high word below q)
    const bool lazyacc = (2 * lc.q) <= (~0ull) / (u64)a.beta;
    // key words of at most kb bits are what the lazy forms assume (kb = bits of q, at least 32: only the high halves
    // are watched -- below 2^32 a key word is harmless for any q)
    const int qbits = 64 - __clzll((long long)lc.q), kb = qbits < 32 ? 32 : qbits;
    u32 keyhi = 0;
    bool fpmac = mode == M_D64 && a.evk_f != nullptr && !a.acc64 && a.beta <= 32;
    if (fpmac) {  // every key limb of this table limb canonical?
        const int nqp = (int)(a.evk_hs / a.T.N);
        u32 bad = 0;
        for (int i = 0; i < 2 * a.beta; ++i) bad |= a.key_bad[(size_t)i * nqp + tl];
        fpmac = bad == 0;
    }
    if (fpmac) {
        ks_fused_body<M_D64, ACC_FP>(a, lc, tl, ks_smem);
    } else if (mode == M_D64) {
        const int sh = 96 - kb;
        if (!a.acc64 && (sh >= 64 || ((3ull * (u64)a.beta * lc.q) >> sh) == 0))
            keyhi = ks_fused_body<M_D64, ACC_WIDE96>(a, lc, tl, ks_smem);
        else
            keyhi = ks_fused_body<M_D64, ACC_LAZY64>(a, lc, tl, ks_smem);
        if (__syncthreads_or((keyhi >> (kb - 32)) != 0)) ks_fused_body<M_D64, ACC_EXACT>(a, lc, tl, ks_smem);
    } else if (mode == M_F64) {
        // beta products of a key word below 2^kb and a digit value below 3q fit 96 bits
        const int sh = 96 - kb;  // 50..64; a shift by 64 is not defined in C++: every product fits then
        if (!a.acc64 && (sh >= 64 || ((3ull * (u64)a.beta * lc.q) >> sh) == 0))
            keyhi = ks_fused_body<M_F64, ACC_WIDE96>(a, lc, tl, ks_smem);
        else  // q < 2^56: beta <= 64 terms below 2q always fit
            keyhi = ks_fused_body<M_F64, ACC_LAZY64>(a, lc, tl, ks_smem);
        if (__syncthreads_or((keyhi >> (kb - 32)) != 0)) ks_fused_body<M_F64, ACC_EXACT>(a, lc, tl, ks_smem);
    } else if (mode == M_FREE) {
        keyhi = ks_fused_body<M_FREE, ACC_LAZY64>(a, lc, tl, ks_smem);
        if (__syncthreads_or((keyhi >> (kb - 32)) != 0)) ks_fused_body<M_FREE, ACC_EXACT>(a, lc, tl, ks_smem);
    } else if (mode == M_LAZY) {
        if (lazyacc) {
            keyhi = ks_fused_body<M_LAZY, ACC_LAZY64>(a, lc, tl, ks_smem);
            if (__syncthreads_or((keyhi >> (kb - 32)) != 0)) ks_fused_body<M_LAZY, ACC_EXACT>(a, lc, tl, ks_smem);
        } else {
            ks_fused_body<M_LAZY, ACC_EXACT>(a, lc, tl, ks_smem);
        }
    } else {
        ks_fused_body<M_LITERAL, ACC_EXACT>(a, lc, tl, ks_smem);
    }
}

// ---- fused digit loop, FP64-class limbs: two batch entries per CTA, key tiles by TMA ---------------------------------
// ks_fused_kernel keeps the 32 key words of a digit in 64 registers (168 in all: 3 CTAs of 4 warps per SM) and waits for
// them where the multiply-accumulate starts (ncu: 15 % of its stall samples sit on the first instruction that reads a key
// word; software prefetch into L1 does not help, profiles/README.md).  This kernel is the ACC_FP path rebuilt around a
// shared-memory key tile:
//   * a 256-thread CTA takes the same (limb, tile) of TWO batch entries (threads 0..127 and 128..255): the tile's twiddles
//     (32 KiB) and the digit's key tile (2 x 16 KiB) are staged once for both, so their L2 -> SM traffic halves;
//   * the key tile of digit i arrives by TMA (cp.async.bulk.tensor.2d through a CUtensorMap of the key seen as rows of
//     16 words, 128-byte swizzle: thread t then reads its 16 consecutive words -- row t of the tile -- with eight
//     conflict-free 128-bit loads) while the digit's transform runs, in boxes of 32 rows: the two warps that read the same
//     rows (warp w of either entry) own a full / empty mbarrier pair, so no warp waits for more than its partner;
//   * each warp fetches its 4 KiB of the next digit tile with one cp.async.bulk (instead of 256 cp.async of 16 bytes)
//     onto its own mbarrier; single tile buffer per entry, re-filled as soon as the exchange has left it;
//   * no key registers: 128 registers, 2 CTAs (16 warps) per SM instead of 12 warps.
// shared memory from a 1024-byte aligned base: key tile 32 KiB | digit tiles 2 x 17 KiB (padded rows) | private twiddles 30 KiB |
// segment twiddles 2 KiB | mbarriers
#define KSF_GROUPS 2
#define KSF_THREADS (KSF_GROUPS * CONTIG_THREADS)
// A segment's 256 words sit in KSF_SEG words: the tile arrives in natural order (the first register block reads cc + 16r),
// the exchange between the register blocks writes word 16r + cc to KSF_ROW*r + cc and reads KSF_ROW*cc + r.  Rows padded to
// 17 words are conflict-free both ways for 64-bit accesses and every address is the thread's base plus a compile-time offset
// (an XOR swizzle costs about 50 address instructions per exchange): step -0.9 % in an ABAB of two libraries.  This kernel's
// shared memory takes the largest carve-out either way, which is what made the same padding lose in the plain transforms.
// Rows of 18 words with eight 128-bit loads on the read side (KSF_ROW 18) measured slower than 17 (profiles/README.md).
#ifndef KSF_ROW
#define KSF_ROW 17u
#endif
#define KSF_SEG (16u * KSF_ROW)
#define KSF_TILE (8u * KSF_SEG)
#define KSF_OFF_TILE 32768u
#define KSF_OFF_TWP (KSF_OFF_TILE + KSF_GROUPS * KSF_TILE * 8u)
#define KSF_OFF_TWSEG (KSF_OFF_TWP + 30u * CONTIG_THREADS * 8u)
#define KSF_OFF_BAR (KSF_OFF_TWSEG + 8u * 32u * 8u)
#define KSF_SMEM_BYTES (KSF_OFF_BAR + 128u + 1024u)

LG_DEV void mbar_init(u32 bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory"); }
LG_DEV void mbar_expect_tx(u32 bar, u32 bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
LG_DEV void mbar_arrive(u32 bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory"); }
LG_DEV bool mbar_try_wait(u32 bar, u32 parity) {
    u32 ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity)
                 : "memory");
    return ok != 0;
}
// with a suspend-time hint (ns): the thread may sleep in the instruction instead of spinning through the loop around it
LG_DEV bool mbar_try_wait_hint(u32 bar, u32 parity, u32 ns) {
    u32 ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok)
                 : "r"(bar), "r"(parity), "r"(ns)
                 : "memory");
    return ok != 0;
}
// bounded: a transfer that never completes traps instead of hanging the device
LG_DEV void mbar_wait(u32 bar, u32 parity, bool hint = false) {
    u32 n = 0;
    while (!(hint ? mbar_try_wait_hint(bar, parity, 2000u) : mbar_try_wait(bar, parity)))
        if (++n > (1u << 26)) __trap();
}
LG_DEV void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
LG_DEV void tma_load_rows(u32 dst, const CUtensorMap* map, int row, u32 bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(map), "r"(0), "r"(row), "r"(bar)
                 : "memory");
}
LG_DEV void bulk_load1(u32 dst, const u64* src, u32 bytes, u32 bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
                 "r"(bytes), "r"(bar)
                 : "memory");
}
// a warp's two segments (512 consecutive words) into their padded places: two copies of 2 KiB on one barrier
LG_DEV void bulk_load(u32 dst, const u64* src, u32 bytes, u32 bar) {
    bulk_load1(dst, src, bytes / 2, bar);
    bulk_load1(dst + KSF_SEG * 8u, src + 256, bytes / 2, bar);
}

__global__ void __launch_bounds__(KSF_THREADS, 2) ks_fused_tma_kernel(const KsFusedArgs a, const __grid_constant__ CUtensorMap kmap) {
    extern __shared__ __align__(16) u64 ks_smem[];
    const u32 raw = (u32)__cvta_generic_to_shared(ks_smem);
    const u32 sbase = (raw + 1023u) & ~1023u;  // the swizzled key boxes want 1024-byte alignment
    u64* const base = ks_smem + ((sbase - raw) >> 3);
    const u32 N = a.T.N;
    const int j = ks_data_limb(a);
    const int tl = a.map(j);
    const LimbConst lc = load_limb_const(a.T, tl);
    const TwConst c = tw_const<true, M_D64>(a.T, lc, tl);
    const u32 t = threadIdx.x, g = t >> 7, tt = t & 127u, warp = t >> 5, lane = t & 31u;
    const u32 sg = tt >> 4, cc = tt & 15u;
    const int batch = a.use_zl >> 8;  // the launcher packs the batch size above the flag
    int b = (int)blockIdx.x * KSF_GROUPS + (int)g;
    const bool active = b < batch;
    if (!active) b = batch - 1;  // an odd batch: the second half of the last CTA repeats the last entry and does not store
    const u32 tile0 = blockIdx.y * CONTIG_TILE;
    const u32 segbase = tile0 + sg * 256u;
    const u32 e0 = segbase + 16 * cc;
    u64* const tilebuf = base + (KSF_OFF_TILE >> 3) + g * KSF_TILE;
    u64* const buf = tilebuf + sg * KSF_SEG;
    u64* const twp = base + (KSF_OFF_TWP >> 3) + 2 * tt;
    u64* const twseg = base + (KSF_OFF_TWSEG >> 3) + sg * 32;
    // key hand-off per warp pair (warp wq of either entry reads rows 32wq .. 32wq+31 of both boxes): full / empty barriers
    // a.pf (the "ks_key_pf" switch, A/B): bit 0 = suspend-time hint on the key wait, bit 1 = one full / empty pair for the
    // whole CTA (every warp waits for the slowest) instead of one per warp pair
    const bool hint = (a.pf & 1) != 0, cta_wide = (a.pf & 2) != 0;
    const u32 wq = cta_wide ? 0u : (warp & 3u);
    const u32 bars = sbase + KSF_OFF_BAR, kbar_full = bars + 8 * wq, kbar_empty = bars + 32 + 8 * wq, tbar = bars + 64 + 8 * warp;
    const u32 kdst = sbase + wq * 4096u;  // the pair's rows of the evk[i][0] box; evk[i][1] 16 KiB further
    const u32 warp_tile = sbase + KSF_OFF_TILE + (g * KSF_TILE + (warp & 3u) * 2u * KSF_SEG) * 8u;  // the warp's two segments

    const u64* din = a.D + (size_t)b * a.d_bs + (size_t)j * N + tile0 + (warp & 3u) * 512u;
    const int own_i = (tl < a.nl) ? tl / a.alpha : -1;
    const size_t key_row0 = ((size_t)tl * N + tile0) >> 4;  // row of the tile's first word within evk_f[0][0]
    const u32 ds_rows = (u32)(a.evk_ds >> 4), hs_rows = (u32)(a.evk_hs >> 4);

    if (t == 0) {
        for (int w = 0; w < 4; ++w) {
            mbar_init(bars + 8 * w, 1);
            mbar_init(bars + 32 + 8 * w, cta_wide ? KSF_THREADS / 32 : KSF_GROUPS);
        }
        for (int w = 0; w < KSF_THREADS / 32; ++w) mbar_init(bars + 64 + 8 * w, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (g == 0) contig_fill_tw<M_D64>(c, N, segbase, cc, twp, twseg);  // twiddles of the tile, once for both entries and all digits
    __syncthreads();
    const int krow0 = (int)key_row0 + 32 * (int)wq;
    const int nbox = cta_wide ? 4 : 1;                                        // boxes of 32 rows per half
    const bool mover = cta_wide ? (t == 0) : (g == 0 && lane == 0);           // the first entry's warp of each pair moves its rows
    if (mover) {
        mbar_expect_tx(kbar_full, 8192u * nbox);
        for (int bx = 0; bx < nbox; ++bx) {
            tma_load_rows(kdst + 4096u * bx, &kmap, krow0 + 32 * bx, kbar_full);
            tma_load_rows(kdst + 4096u * bx + 16384u, &kmap, krow0 + 32 * bx + (int)hs_rows, kbar_full);
        }
    }
    if (lane == 0 && own_i != 0) {
        mbar_expect_tx(tbar, 4096u);
        bulk_load(warp_tile, din, 4096u, tbar);
    }
    u32 tphase = 0;
    double acc0[16], acc1[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) acc0[r] = acc1[r] = 0.0;
    // the thread's row of the key boxes: 16-byte chunk p of row tt sits at chunk p ^ (tt & 7)
    const u64* const krow = base + tt * 16u;
    const u32 ksw = tt & 7u;
#pragma unroll 1
    for (int i = 0; i < a.beta; ++i) {
        u64 x[16];
        if (i == own_i) {  // ckks/evaluator.go:1579-1584, bfv/evaluator.go:776-780
            const u64* cx = a.cx + (size_t)b * a.cx_bs + (size_t)j * (a.cx_ls ? a.cx_ls : N) + e0;
#pragma unroll
            for (int h = 0; h < 4; ++h) {
                u64 v[4];
                ld256(v, cx + 4 * h);
#pragma unroll
                for (int e = 0; e < 4; ++e) x[4 * h + e] = d2bits(u52_to_d(bred_add(v[e], lc.q, lc.u0), 4503599627370496.0));
            }
            if (i + 1 < a.beta && lane == 0) {  // the next digit's tile (nothing was in flight for this one)
                mbar_expect_tx(tbar, 4096u);
                bulk_load(warp_tile, din + (size_t)(i + 1) * a.d_ds, 4096u, tbar);
            }
        } else {
            mbar_wait(tbar, tphase);
            tphase ^= 1u;
#pragma unroll
            for (int r = 0; r < 16; ++r) x[r] = buf[cc + 16 * r];
            fwd_stages_sm<3, 1, M_D64>(x, c, twseg);
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 16; ++r) buf[KSF_ROW * r + cc] = x[r];
            __syncwarp();
            if (KSF_ROW % 2 == 0) {
                const ulonglong2* row = reinterpret_cast<const ulonglong2*>(buf + KSF_ROW * cc);
#pragma unroll
                for (int h = 0; h < 8; ++h) {
                    const ulonglong2 v = row[h];
                    x[2 * h] = v.x;
                    x[2 * h + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int r = 0; r < 16; ++r) x[r] = buf[KSF_ROW * cc + r];
            }
            __syncwarp();
            if (i + 1 < a.beta && i + 1 != own_i && lane == 0) {  // the buffer is free: fetch the next digit's tile
                fence_proxy_async();
                mbar_expect_tx(tbar, 4096u);
                bulk_load(warp_tile, din + (size_t)(i + 1) * a.d_ds, 4096u, tbar);
            }
            fwd_stages_sm<3, CONTIG_THREADS, M_D64>(x, c, twp);
        }
        mbar_wait(kbar_full, (u32)i & 1u, hint);
#pragma unroll
        for (int p = 0; p < 8; ++p) {
            const ulonglong2 k0 = *reinterpret_cast<const ulonglong2*>(krow + 2u * (p ^ ksw));
            const ulonglong2 k1 = *reinterpret_cast<const ulonglong2*>(krow + 2048u + 2u * (p ^ ksw));
            const double xa = bits2d(x[2 * p]), xb = bits2d(x[2 * p + 1]);
            const double k0a = bits2d(k0.x), k0b = bits2d(k0.y), k1a = bits2d(k1.x), k1b = bits2d(k1.y);
            acc0[2 * p] = __dadd_rn(acc0[2 * p], d64_mul(k0a, __dmul_rd(k0a, c.qinvd), xa, c.qd));
            acc1[2 * p] = __dadd_rn(acc1[2 * p], d64_mul(k1a, __dmul_rd(k1a, c.qinvd), xa, c.qd));
            acc0[2 * p + 1] = __dadd_rn(acc0[2 * p + 1], d64_mul(k0b, __dmul_rd(k0b, c.qinvd), xb, c.qd));
            acc1[2 * p + 1] = __dadd_rn(acc1[2 * p + 1], d64_mul(k1b, __dmul_rd(k1b, c.qinvd), xb, c.qd));
        }
        // hand the pair's key rows back; the first entry's warp re-fills them with the next digit's once both have done so
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(kbar_empty);
            if (mover && i + 1 < a.beta) {
                mbar_wait(kbar_empty, (u32)i & 1u);
                fence_proxy_async();
                mbar_expect_tx(kbar_full, 8192u * nbox);
                const int row = krow0 + (i + 1) * (int)ds_rows;
                for (int bx = 0; bx < nbox; ++bx) {
                    tma_load_rows(kdst + 4096u * bx, &kmap, row + 32 * bx, kbar_full);
                    tma_load_rows(kdst + 4096u * bx + 16384u, &kmap, row + 32 * bx + (int)hs_rows, kbar_full);
                }
            }
        }
    }
    if (!active) return;
    u64* o0 = a.acc0 + (size_t)b * a.acc_bs + (size_t)j * N + e0;
    u64* o1 = a.acc1 + (size_t)b * a.acc_bs + (size_t)j * N + e0;
#pragma unroll
    for (int h = 0; h < 4; ++h) {  // |sum| < 18q
        st256(o0 + 4 * h, d64_canon(d2bits(acc0[4 * h]), c), d64_canon(d2bits(acc0[4 * h + 1]), c), d64_canon(d2bits(acc0[4 * h + 2]), c),
              d64_canon(d2bits(acc0[4 * h + 3]), c));
        st256(o1 + 4 * h, d64_canon(d2bits(acc1[4 * h]), c), d64_canon(d2bits(acc1[4 * h + 1]), c), d64_canon(d2bits(acc1[4 * h + 2]), c),
              d64_canon(d2bits(acc1[4 * h + 3]), c));
    }
}

// ---- forward strided phase fed by TMA ----------------------------------------------------------------------------
// ntt_fwd_strided issues its 16 loads per thread in one burst, computes, then stores: a CTA has nothing in flight for most
// of its life, and the digit launch of a key switch (10.4 GB) runs at 4.8 TB/s although its access pattern alone reaches
// 6.3 TB/s (profiles/r02_strided_copy.txt).  Here the tile -- all 2^L rows of W adjacent columns = a 2-D box of the limb seen
// as rows of 256 words -- moves by TMA in both directions: a CTA takes the same (limb, tile) of up to `bpc` batch entries,
// loads run two entries ahead into a ring of three 32 KiB buffers (cp.async.bulk.tensor.2d + mbarrier), the register blocks
// work IN PLACE on the buffer (the exchange between them uses the slots the values were read from), and the result leaves with
// a TMA store (bulk_group) while the next entry is transformed.  No thread issues a global load or store; twiddles are staged
// once per CTA.  256 threads, up to 128 registers, 2 CTAs/SM.
#ifndef STMA_STAGES
#define STMA_STAGES 3
#endif
#ifndef STMA_MINB
#define STMA_MINB 2
#endif
#define STMA_SMEM_BYTES(L) (STMA_STAGES * 32768u + 32u * ((1u << ((L) - 4)) + 1u) * 8u + 64u + 128u)

LG_DEV void tma_load_box(u32 dst, const CUtensorMap* map, int col, int row, u32 bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(map), "r"(col), "r"(row), "r"(bar)
                 : "memory");
}
LG_DEV void tma_store_box(const CUtensorMap* map, int col, int row, u32 src) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(map), "r"(col), "r"(row), "r"(src)
                 : "memory");
}
LG_DEV void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N_>
LG_DEV void bulk_wait_read() {
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N_) : "memory");
}

// the two register blocks of the strided phase on a tile held in shared memory as [2^L rows][W columns], in place
template <int L, int MODE>
LG_DEV void fwd_strided_tile(const NttArgs& a, const LimbConst& lc, int tl, u64* tile, const u64* tws_sm) {
    constexpr int G = 1 << (L - 4);  // threads per column
    constexpr int W = 256 / G;       // columns per CTA
    constexpr int N2 = L - 4;        // stages of the second register block
    const TwConst c = tw_const<true, MODE>(a.T, lc, tl);
    const int t = threadIdx.x, col = t % W, g = t / W;
    u64 x[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) x[r] = tile[(g + r * G) * W + col];
    if (a.bcast.enabled && a.bcast.add != nullptr) {  // ring_scaling.go:99-103: + (q_j - pHalf mod q_j), unreduced
        const u64 add = __ldg(a.bcast.add + tl);
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] += add;
    }
    if (MODE != M_LITERAL) {  // headroom of the lazy butterflies, as in fwd_strided_body
        const u32 sh = MODE == M_FREE ? 63u : (MODE == M_F64 ? 50u : (MODE == M_D64 ? 49u : 66u - (u32)__clzll((long long)c.q)));
        u64 o = 0;
#pragma unroll
        for (int r = 0; r < 16; ++r) o |= x[r];
        if (o >> sh) {
#pragma unroll
            for (int r = 0; r < 16; ++r)
                if (x[r] >> sh) x[r] = bred_add(x[r], c.q, lc.u0);
        }
    }
    if (MODE == M_D64) {
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] = d2bits(u52_to_d(x[r], 4503599627370496.0));
    }
    fwd_stages_sm<3, 1, MODE>(x, c, tws_sm);
    if (N2 > 0) {
#pragma unroll
        for (int r = 0; r < 16; ++r) tile[(g + r * G) * W + col] = x[r];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] = tile[(16 * g + r) * W + col];
        const u64* twg = tws_sm + (1 + g) * 32;
        fwd_stages_sm<(N2 > 0 ? N2 - 1 : 0), 1, MODE>(x, c, twg);
#pragma unroll
        for (int r = 0; r < 16; ++r) tile[(16 * g + r) * W + col] = x[r];
    } else {
#pragma unroll
        for (int r = 0; r < 16; ++r) tile[(g + r * G) * W + col] = x[r];
    }
}

template <int L, bool LITERAL>
__global__ void __launch_bounds__(256, STMA_MINB) ntt_fwd_strided_tma(const NttArgs a, const __grid_constant__ CUtensorMap in_map,
                                                               const __grid_constant__ CUtensorMap out_map, int batch, int bpc) {
    extern __shared__ __align__(16) u64 ks_smem[];
    constexpr int W = 256 >> (L - 4);
    const u32 raw = (u32)__cvta_generic_to_shared(ks_smem);
    const u32 sbase = (raw + 127u) & ~127u;
    u64* const base = ks_smem + ((sbase - raw) >> 3);
    u64* const tws_sm = base + STMA_STAGES * 4096;
    const u32 bars = sbase + STMA_STAGES * 32768u + 32u * ((1u << (L - 4)) + 1u) * 8u;
    const int j = (int)blockIdx.z, tile_x = (int)blockIdx.x;
    const int b0 = (int)blockIdx.y * bpc, nb = (batch - b0) < bpc ? (batch - b0) : bpc;
    const int tl = a.map(j);
    if (a.skip_alpha > 0) {  // digit-batched launch: bpc divides skip_div, so a group never straddles two digits
        const int dg = b0 / a.skip_div;
        if (tl < a.skip_nl && tl >= dg * a.skip_alpha && tl < (dg + 1) * a.skip_alpha) return;
    } else if (j >= a.skip0 && j < a.skip1) {
        return;
    }
    const LimbConst lc = load_limb_const(a.T, tl);
    const int mode = LITERAL ? M_LITERAL : fwd_mode(lc.q, a.no_d64);
    const u32 t = threadIdx.x;
    // rows of 256 words: limb j of entry b starts at row (b*bstride + j*ls) / 256
    const size_t in_ls = a.bcast.enabled ? 0 : (a.in_ls ? a.in_ls : a.T.N), out_ls = a.out_ls ? a.out_ls : a.T.N;
    const int col0 = tile_x * W;
    auto in_row = [&](int i) { return (int)(((size_t)(b0 + i) * a.in_bstride + (size_t)j * in_ls) >> 8); };
    auto out_row = [&](int i) { return (int)(((size_t)(b0 + i) * a.out_bstride + (size_t)j * out_ls) >> 8); };
    if (t == 0) {
        for (int s = 0; s < STMA_STAGES; ++s) mbar_init(bars + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    {
        const TwConst c = LITERAL ? tw_const<true, M_LITERAL>(a.T, lc, tl)
                                  : (mode == M_D64 ? tw_const<true, M_D64>(a.T, lc, tl)
                                                   : (mode == M_F64 ? tw_const<true, M_F64>(a.T, lc, tl) : tw_const<true, M_FREE>(a.T, lc, tl)));
        if (LITERAL || mode == M_LITERAL)
            fill_strided_tw<L, true>(tws_sm, tw_const<true, M_LITERAL>(a.T, lc, tl));
        else
            fill_strided_tw<L, false>(tws_sm, c);
    }
    __syncthreads();
    if (t == 0) {
        for (int i = 0; i < STMA_STAGES - 1 && i < nb; ++i) {
            mbar_expect_tx(bars + 8 * i, 32768u);
            tma_load_box(sbase + 32768u * i, &in_map, col0, in_row(i), bars + 8 * i);
        }
    }
#pragma unroll 1
    for (int i = 0; i < nb; ++i) {
        const int s = i % STMA_STAGES;
        if (t == 0 && i + STMA_STAGES - 1 < nb) {
            // the buffer of entry i+STAGES-1 held entry i-1: its store must have finished reading shared memory
            bulk_wait_read<0>();
            const int s2 = (i + STMA_STAGES - 1) % STMA_STAGES;
            mbar_expect_tx(bars + 8 * s2, 32768u);
            tma_load_box(sbase + 32768u * s2, &in_map, col0, in_row(i + STMA_STAGES - 1), bars + 8 * s2);
        }
        mbar_wait(bars + 8 * s, (u32)(i / STMA_STAGES) & 1u);
        u64* tile = base + s * 4096;
        if (mode == M_D64)
            fwd_strided_tile<L, M_D64>(a, lc, tl, tile, tws_sm);
        else if (mode == M_F64)
            fwd_strided_tile<L, M_F64>(a, lc, tl, tile, tws_sm);
        else if (mode == M_FREE)
            fwd_strided_tile<L, M_FREE>(a, lc, tl, tile, tws_sm);
        else if (mode == M_LAZY)
            fwd_strided_tile<L, M_LAZY>(a, lc, tl, tile, tws_sm);
        else
            fwd_strided_tile<L, M_LITERAL>(a, lc, tl, tile, tws_sm);
        fence_proxy_async();  // every thread's writes to the tile, before the async-proxy store reads them
        __syncthreads();
        if (t == 0) {
            tma_store_box(&out_map, col0, out_row(i), sbase + 32768u * s);
            bulk_commit();
        }
    }
    if (t == 0) bulk_wait_read<0>();  // shared memory stays valid until the last store has read it
}

LG_DEV bool inv_flagged(const NttArgs& a) {
    return a.flags != nullptr && a.flags[cta_z(a)] != 0;
}

// ---- inverse, strided phase: last L stages + MRed by N^-1 --------------------
template <int L, int MODE>
LG_DEV void inv_strided_body(const NttArgs& a, const LimbSetup& s, u64* sm, u64* tws_sm) {
    constexpr int G = 1 << (L - 4);
    constexpr int W = 256 / G;
    constexpr int N2 = L - 4;
    const TwConst c = tw_const<false, MODE>(a.T, s.c, s.tl);
    const int t = threadIdx.x, col = t % W, g = t / W;
    const u32 colg = cta_y(a) * W + col;
    u64 x[16];
    if (N2 > 0) {
        const u64* in = s.in + colg + g * 16 * 256;
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] = in[r * 256];
    } else {
        const u64* in = s.in + colg + g * 256;
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] = in[r * G * 256];
    }
    fill_strided_tw<L, MODE == M_LITERAL>(tws_sm, c);
    __syncthreads();
    if (MODE == M_D64) d64_reduce_all(x, c);  // the contiguous phase left sums of 16 values: -> [-q, 2q)
    if (N2 > 0) {
        const u64* twg = tws_sm + (1 + g) * 32;
        inv_stages_sm<(N2 > 0 ? N2 - 1 : 0), 1, MODE>(x, c, twg, 8u);
        if (MODE == M_D64) d64_reduce_all(x, c);
#pragma unroll
        for (int r = 0; r < 16; ++r) sm[(16 * g + r) * W + col] = x[r];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] = sm[(g + r * G) * W + col];
    }
    inv_stages_sm<3, 1, MODE>(x, c, tws_sm, 8u + N2);
    // ring/ntt.go:136-138
    u64* out = s.out + colg + g * 256;
    if (MODE == M_LITERAL) {
        const u64 ninv = a.T.ninv[s.tl];
#pragma unroll
        for (int r = 0; r < 16; ++r) out[r * G * 256] = mred(x[r], ninv, c.q, c.qinv);
    } else if (MODE == M_D64) {
        // |v| <= 32q: reduced to [-q, 2q) and shifted to [0, 3q) (the multiplication needs |y| < 2^51 and, for a
        // result in [0, 2q), y >= 0); the product with N^-1 then lands in [0, 2q)
        const double nf = bits2d(a.T.ninv_f[2 * s.tl]), nd = bits2d(a.T.ninv_f[2 * s.tl + 1]);
#pragma unroll
        for (int r = 0; r < 16; ++r) {
            const double y = __dadd_rn(d64_red(bits2d(x[r]), c.qinvd, c.qd), c.qd);
            out[r * G * 256] = cred(d_to_u52(d64_mul(nf, nd, y, c.qd), 4503599627370496.0), c.q);
        }
    } else {
        const u64 nw = a.T.ninv_w[2 * s.tl], nws = a.T.ninv_w[2 * s.tl + 1];
#pragma unroll
        for (int r = 0; r < 16; ++r) out[r * G * 256] = cred(shoup_exact(nw, nws, x[r], c.nq), c.q);
    }
}

template <int L, bool LITERAL>
__global__ void __launch_bounds__(256, STRIDED_MINB) ntt_inv_strided(const NttArgs a) {
    __shared__ u64 sm[(L - 4) > 0 ? 4096 : 1];
    __shared__ __align__(16) u64 tws_sm[32 * ((1 << (L - 4)) + 1)];
    const LimbSetup s = setup_limb(a);
    if (s.skip) return;
    const int mode = (LITERAL || inv_flagged(a)) ? M_LITERAL : inv_mode(s.c.q, a.no_d64);
    if (mode == M_D64)
        inv_strided_body<L, M_D64>(a, s, sm, tws_sm);
    else if (mode == M_FREE)
        inv_strided_body<L, M_FREE>(a, s, sm, tws_sm);
    else if (mode == M_LAZY)
        inv_strided_body<L, M_LAZY>(a, s, sm, tws_sm);
    else
        inv_strided_body<L, M_LITERAL>(a, s, sm, tws_sm);
}

// flags[j] != 0 when some word of data limb j, in ANY batch entry, exceeds 2q: the inverse transform of that limb is then
// literal for the whole batch.  One decision per limb keeps the two phases consistent -- they exchange raw doubles
// (FP64-only butterflies) or integers (literal) through HBM, and their CTAs group the batch entries differently.
// The flags are zeroed by the launcher, every CTA scans up to 4096 words.
__global__ void __launch_bounds__(256) range_flags_kernel(const NttArgs a, u32* flags) {
    const int j = blockIdx.z, b = blockIdx.x;
    const u64 twoq = 2 * a.T.q[a.map(j)];
    const u32 half = a.T.N >> 1, lo = blockIdx.y * 2048u, hi = (lo + 2048u < half) ? lo + 2048u : half;
    const ulonglong2* in = reinterpret_cast<const ulonglong2*>(a.in + (size_t)b * a.in_bstride + (size_t)j * (a.in_ls ? a.in_ls : a.T.N));
    int bad = 0;
    for (u32 i = lo + threadIdx.x; i < hi; i += 256) {
        const ulonglong2 v = in[i];
        bad |= (v.x > twoq) | (v.y > twoq);
    }
    bad = __syncthreads_or(bad);
    if (bad && threadIdx.x == 0) atomicOr(flags + j, 1u);
}

// ---- small rings (logN <= 11): one CTA per limb, radix-2 in shared memory ----
template <bool FWD>
__global__ void ntt_small(const NttArgs a) {
    extern __shared__ u64 dsm[];
    const LimbSetup s = setup_limb(a);
    if (s.skip) return;
    const u32 N = a.T.N;
    const u64 q = s.c.q, qinv = s.c.qinv, twoq = 2 * s.c.q;
    const u64* tw = (FWD ? a.T.psi : a.T.psi_inv) + (size_t)s.tl * N;
    for (u32 i = threadIdx.x; i < N; i += blockDim.x) dsm[i] = s.in[i];
    __syncthreads();
    if (FWD) {
        u32 sh = a.T.logN - 1;  // log2(t)
        for (u32 m = 1; m < N; m <<= 1, --sh) {
            for (u32 k = threadIdx.x; k < (N >> 1); k += blockDim.x) {
                const u32 i = k >> sh, jj = k & ((1u << sh) - 1);
                const u32 j = (i << (sh + 1)) + jj;
                u64 U = dsm[j], V = dsm[j + (1u << sh)];
                butterfly_fwd(U, V, tw[m + i], q, qinv, twoq);
                dsm[j] = U;
                dsm[j + (1u << sh)] = V;
            }
            __syncthreads();
        }
        for (u32 i = threadIdx.x; i < N; i += blockDim.x) s.out[i] = bred_add(dsm[i], q, s.c.u0);
    } else {
        const u64 ninv = a.T.ninv[s.tl];
        u32 sh = 0;
        for (u32 h = N >> 1; h >= 1; h >>= 1, ++sh) {
            for (u32 k = threadIdx.x; k < (N >> 1); k += blockDim.x) {
                const u32 i = k >> sh, jj = k & ((1u << sh) - 1);
                const u32 j = (i << (sh + 1)) + jj;
                u64 U = dsm[j], V = dsm[j + (1u << sh)];
                butterfly_inv(U, V, tw[h + i], q, qinv, twoq);
                dsm[j] = U;
                dsm[j + (1u << sh)] = V;
            }
            __syncthreads();
        }
        for (u32 i = threadIdx.x; i < N; i += blockDim.x) s.out[i] = mred(dsm[i], ninv, q, qinv);
    }
}

template <int L>
void launch_strided(bool fwd, bool literal, const NttArgs& a, dim3 grid, cudaStream_t st) {
    if (fwd) {
        if (literal)
            ntt_fwd_strided<L, true><<<grid, 256, 0, st>>>(a);
        else
            ntt_fwd_strided<L, false><<<grid, 256, 0, st>>>(a);
    } else {
        if (literal)
            ntt_inv_strided<L, true><<<grid, 256, 0, st>>>(a);
        else
            ntt_inv_strided<L, false><<<grid, 256, 0, st>>>(a);
    }
}

// CUtensorMap of a buffer of 64-bit words seen as `rows` rows of `cols` words (row pitch = cols words), boxes of box_cols x
// box_rows; 0 = ok (needs a driver with cuTensorMapEncodeTiled)
static int encode_words_2d(void* map, const u64* ptr, size_t cols, size_t rows, unsigned box_cols, unsigned box_rows, bool swizzle128) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = [] {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qr;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            fn = nullptr;
        }
        return (EncodeFn)fn;
    }();
    if (!encode || rows == 0 || rows > 0x7fffffffull || ((uintptr_t)ptr & 15) != 0) return 1;
    const cuuint64_t gdim[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)cols * 8};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estride[2] = {1, 1};
    CUtensorMap m;
    if (encode(&m, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, (void*)ptr, gdim, gstride, box, estride, CU_TENSOR_MAP_INTERLEAVE_NONE,
               swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return 1;
    static_assert(sizeof(CUtensorMap) == 128, "CUtensorMap is 128 bytes");
    memcpy(map, &m, sizeof(m));
    return 0;
}

// forward strided phase by TMA (ntt_fwd_strided_tma): false = not applicable, the caller launches ntt_fwd_strided.  (The
// same ring for the inverse strided phase measured slower, 620 against 610 us per 1088 limb-NTTs: with its N^-1 product that
// phase is bound by the FP64 pipe, 70 % busy, and loses more with 16 instead of 32 warps per SM than the asynchronous
// copies give back.)
template <int L>
static bool launch_strided_tma(bool literal, const NttArgs& a, int batch, int nlimbs, cudaStream_t st) {
    const u32 N = a.T.N;
    const size_t in_ls = a.bcast.enabled ? 0 : (a.in_ls ? a.in_ls : N), out_ls = a.out_ls ? a.out_ls : N;
    if ((a.in_bstride | a.out_bstride | in_ls | out_ls) & 255) return false;
    if (batch > 1 && (a.in_bstride == 0 || a.out_bstride == 0)) return false;
    // The ring pays where the phase is HBM-bound: limbs on the 8-instruction FP64 butterflies, and enough work for two
    // waves of CTAs that still pipeline four entries each.  Integer-butterfly limbs (60-bit rings: C1 3.43 -> 3.36 M NTT/s
    // with the ring) and small batches (C5, 8 parties: 13.6 -> 12.7 k rounds/s) are bound by issue slots and by grid size and
    // keep the per-thread-load kernel with its 32 warps per SM.
    int fp = 0;
    for (int j = 0; j < nlimbs; ++j) {
        const int tl = a.map(j);
        fp += (tl >= 0 && tl < 64) ? (int)((a.T.d64_mask >> tl) & 1) : 0;
    }
    if (a.no_d64 || 2 * fp < nlimbs) return false;
    if ((long)(N / 4096) * nlimbs * ((batch + 3) / 4) < 2L * 148 * 2) return false;
    const size_t rows_in = ((size_t)(batch - 1) * a.in_bstride + (size_t)(nlimbs - 1) * in_ls + N) >> 8;
    const size_t rows_out = ((size_t)(batch - 1) * a.out_bstride + (size_t)(nlimbs - 1) * out_ls + N) >> 8;
    constexpr unsigned W = 256u >> (L - 4), R = 1u << L;
    CUtensorMap in_map, out_map;
    if (encode_words_2d(&in_map, a.in, 256, rows_in, W, R, false) != 0) return false;
    if (encode_words_2d(&out_map, a.out, 256, rows_out, W, R, false) != 0) return false;
    int bpc = batch < 8 ? batch : 8;
    const long tiles = N / 4096;
    while (bpc > 2 && tiles * nlimbs * ((batch + bpc - 1) / bpc) < 2L * 148 * 2) bpc = (bpc + 1) / 2;
    if (a.skip_alpha > 0)
        while (a.skip_div % bpc) --bpc;
    const int groups = (batch + bpc - 1) / bpc;
    if (groups > 65535) return false;
    const dim3 grid((unsigned)tiles, (unsigned)groups, (unsigned)nlimbs);
    const size_t smem = STMA_SMEM_BYTES(L);
    if (literal) {
        lg_ensure_dyn_smem<ntt_fwd_strided_tma<L, true>>(smem);
        ntt_fwd_strided_tma<L, true><<<grid, 256, smem, st>>>(a, in_map, out_map, batch, bpc);
    } else {
        lg_ensure_dyn_smem<ntt_fwd_strided_tma<L, false>>(smem);
        ntt_fwd_strided_tma<L, false><<<grid, 256, smem, st>>>(a, in_map, out_map, batch, bpc);
    }
    return true;
}

void launch_strided_any(int L, bool fwd, bool literal, const NttArgs& a0, dim3 grid, cudaStream_t st) {
    NttArgs a = a0;
    if (fwd && grid.x >= 2 && !lg_switches().no_strided_tma.load(std::memory_order_relaxed)) {  // grid = (batch, tiles, limbs)
        bool done = false;
        switch (L) {
            case 4: done = launch_strided_tma<4>(literal, a, (int)grid.x, (int)grid.z, st); break;
            case 5: done = launch_strided_tma<5>(literal, a, (int)grid.x, (int)grid.z, st); break;
            case 6: done = launch_strided_tma<6>(literal, a, (int)grid.x, (int)grid.z, st); break;
            case 7: done = launch_strided_tma<7>(literal, a, (int)grid.x, (int)grid.z, st); break;
            default: done = launch_strided_tma<8>(literal, a, (int)grid.x, (int)grid.z, st); break;
        }
        if (done) return;
    }
    a.tfast = (lg_switches().tile_fastest.load(std::memory_order_relaxed) && grid.x <= 65535u) ? 1 : 0;
    if (a.tfast) grid = dim3(grid.y, grid.x, grid.z);
    switch (L) {
        case 4: launch_strided<4>(fwd, literal, a, grid, st); break;
        case 5: launch_strided<5>(fwd, literal, a, grid, st); break;
        case 6: launch_strided<6>(fwd, literal, a, grid, st); break;
        case 7: launch_strided<7>(fwd, literal, a, grid, st); break;
        default: launch_strided<8>(fwd, literal, a, grid, st); break;
    }
}

// entries per CTA of the pipelined contiguous phase: as many as keep about two waves of CTAs in flight
template <bool FWD, bool LITERAL, bool TAIL>
void launch_contig_pipe_t(const NttArgs& a, int nlimbs, int batch, cudaStream_t st) {
    const int tiles = (int)(a.T.N / CONTIG_TILE);
    int bpc = batch < 8 ? batch : 8;
    while (bpc > 1 && (long)tiles * nlimbs * ((batch + bpc - 1) / bpc) < 2L * 148 * 4) bpc = (bpc + 1) / 2;
    if (a.skip_alpha > 0)
        while (a.skip_div % bpc) --bpc;
    const size_t smem = PIPE_SMEM_WORDS * sizeof(u64);
    lg_ensure_dyn_smem<ntt_contig_pipe<FWD, LITERAL, TAIL>>(smem);
    ntt_contig_pipe<FWD, LITERAL, TAIL><<<dim3((batch + bpc - 1) / bpc, tiles, nlimbs), CONTIG_THREADS, smem, st>>>(a, batch, bpc);
}
void launch_contig_pipe(bool fwd, bool literal, const NttArgs& a, int nlimbs, int batch, cudaStream_t st) {
    if (fwd && a.tail.enabled) {
        if (literal)
            launch_contig_pipe_t<true, true, true>(a, nlimbs, batch, st);
        else
            launch_contig_pipe_t<true, false, true>(a, nlimbs, batch, st);
    } else if (fwd) {
        if (literal)
            launch_contig_pipe_t<true, true, false>(a, nlimbs, batch, st);
        else
            launch_contig_pipe_t<true, false, false>(a, nlimbs, batch, st);
    } else {
        if (literal)
            launch_contig_pipe_t<false, true, false>(a, nlimbs, batch, st);
        else
            launch_contig_pipe_t<false, false, false>(a, nlimbs, batch, st);
    }
}

bool literal_ntt() { return lg_switches().literal_ntt.load(std::memory_order_relaxed) != 0; }

}  // namespace

// auxiliary streams, one set per (host thread, device): forked from and joined to the caller's stream with events
LgAux* lg_aux_streams() {
    thread_local LgAux* per_dev[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return nullptr;
    if (!per_dev[dev]) {
        LgAux* a = new LgAux;
        bool ok = cudaEventCreateWithFlags(&a->fork, cudaEventDisableTiming) == cudaSuccess;
        for (int k = 0; k < LG_AUX_STREAMS && ok; ++k)
            ok = cudaStreamCreateWithFlags(&a->s[k], cudaStreamNonBlocking) == cudaSuccess &&
                 cudaEventCreateWithFlags(&a->join[k], cudaEventDisableTiming) == cudaSuccess;
        if (!ok) {
            cudaGetLastError();
            delete a;
            return nullptr;
        }
        per_dev[dev] = a;
    }
    return per_dev[dev];
}
void lg_aux_fork(LgAux* a, cudaStream_t st, int n) {
    cudaEventRecord(a->fork, st);
    for (int k = 0; k < n; ++k) cudaStreamWaitEvent(a->s[k], a->fork, 0);
}
void lg_aux_join(LgAux* a, cudaStream_t st, int n) {
    for (int k = 0; k < n; ++k) {
        cudaEventRecord(a->join[k], a->s[k]);
        cudaStreamWaitEvent(st, a->join[k], 0);
    }
}

namespace {
typedef LgAux NttAux;
NttAux* ntt_aux() { return lg_aux_streams(); }
}  // namespace

int lg_launch_ntt(const NttArgs& args, int nlimbs, int batch, bool inverse, cudaStream_t st) {
    if (nlimbs <= 0 || batch <= 0) return 0;
    const u32 logN = args.T.logN, N = args.T.N;
    if (logN < 1 || logN > 16) return 1;
    if ((args.tail.enabled || args.bcast.enabled) && (inverse || logN <= 11)) return 1;
    if (logN <= 11) {
        const u32 threads = (N >> 1) < 32 ? 32 : ((N >> 1) > 512 ? 512 : (N >> 1));
        dim3 grid(batch, 1, nlimbs);
        if (inverse)
            ntt_small<false><<<grid, threads, N * sizeof(u64), st>>>(args);
        else
            ntt_small<true><<<grid, threads, N * sizeof(u64), st>>>(args);
        lg_g_launches += 1;
        return 0;
    }
    const int L = (int)logN - 8;
    const bool literal = literal_ntt();
    const dim3 sgrid(batch, N / 4096, nlimbs);
    // Groups of batch entries whose first-phase output fits the L2 budget: the second phase then reads it from L2
    // instead of HBM (the transform is bandwidth-bound once the butterflies run on the FP64 pipe).
    int cb = batch;
    const size_t l2_bytes = (size_t)lg_switches().ntt_l2_bytes.load(std::memory_order_relaxed);
    if (l2_bytes && args.skip_alpha == 0) {
        const size_t per = (size_t)nlimbs * N * sizeof(u64);
        cb = (int)(l2_bytes / per);
        if (cb < 1) cb = 1;
        if (cb > batch) cb = batch;
    }
    (void)sgrid;
    // "ntt_l2_streams" = 2: groups of LIMBS over the whole batch instead (plain transforms only) -- the contiguous phase
    // keeps its eight batch entries per CTA and a group touches the twiddles of its own limbs alone
    if (l2_bytes && lg_switches().ntt_l2_streams.load(std::memory_order_relaxed) == 2 && !args.tail.enabled && !args.bcast.enabled &&
        !args.flags && args.skip_alpha == 0 && args.skip0 >= args.skip1) {
        int cl = (int)(l2_bytes / ((size_t)batch * N * sizeof(u64)));
        if (cl < 1) cl = 1;
        if (cl < nlimbs) {
            NttAux* aux = ntt_aux();
            if (aux) {
                cudaEventRecord(aux->fork, st);
                for (int k = 0; k < 2; ++k) cudaStreamWaitEvent(aux->s[k], aux->fork, 0);
            }
            int gi = 0;
            for (int j0 = 0; j0 < nlimbs; j0 += cl, ++gi) {
                cudaStream_t gs = aux ? aux->s[gi & 1] : st;
                const int nj = (nlimbs - j0) < cl ? (nlimbs - j0) : cl;
                NttArgs first = args;
                first.no_d64 = lg_switches().no_d64_ntt.load(std::memory_order_relaxed) ? 1 : 0;
                first.in = args.in + (size_t)j0 * (args.in_ls ? args.in_ls : N);
                first.out = args.out + (size_t)j0 * (args.out_ls ? args.out_ls : N);
                const LimbMap m = args.map;
                first.map.n0 = m.n0 > j0 ? m.n0 - j0 : 0;
                first.map.l0 = m.l0 + j0 * m.st;
                first.map.l1 = j0 > m.n0 ? m.l1 + (j0 - m.n0) * m.st : m.l1;
                NttArgs second = first;
                second.in = first.out;
                second.in_bstride = args.out_bstride;
                second.in_ls = args.out_ls;
                second.rev = 0;
                second.pf = 0;
                const dim3 grid(batch, N / 4096, nj);
                if (!inverse) {
                    launch_strided_any(L, true, literal, first, grid, gs);
                    launch_contig_pipe(true, literal, second, nj, batch, gs);
                } else {
                    launch_contig_pipe(false, literal, first, nj, batch, gs);
                    launch_strided_any(L, false, literal, second, grid, gs);
                }
                lg_g_launches += 2;
            }
            if (aux) {
                for (int k = 0; k < 2; ++k) {
                    cudaEventRecord(aux->join[k], aux->s[k]);
                    cudaStreamWaitEvent(st, aux->join[k], 0);
                }
            }
            return 0;
        }
    }
    // With L2-sized groups the launch pairs of consecutive groups go to two auxiliary streams in turn, so that the tail of
    // one group's kernels overlaps the next group's (the groups are independent); `st` forks into them and joins them.
    const bool fork = cb < batch && lg_switches().ntt_l2_streams.load(std::memory_order_relaxed) != 0;
    NttAux* aux = fork ? ntt_aux() : nullptr;
    if (aux) {
        cudaEventRecord(aux->fork, st);
        for (int k = 0; k < 2; ++k) cudaStreamWaitEvent(aux->s[k], aux->fork, 0);
    }
    int gi = 0;
    for (int g0 = 0; g0 < batch; g0 += cb, ++gi) {
        cudaStream_t gs = aux ? aux->s[gi & 1] : st;
        const int nb = (batch - g0) < cb ? (batch - g0) : cb;
        NttArgs first = args;
        first.no_d64 = lg_switches().no_d64_ntt.load(std::memory_order_relaxed) ? 1 : 0;
        first.in = args.in + (size_t)g0 * args.in_bstride;
        first.out = args.out + (size_t)g0 * args.out_bstride;
        first.batch0 = g0;
        first.flags = args.flags;  // per limb, whatever the batch group
        NttArgs second = first;  // the second phase runs in place on the output
        second.in = first.out;
        second.in_bstride = args.out_bstride;
        second.in_ls = args.out_ls;
        second.bcast.enabled = 0;
        second.rev = lg_switches().reverse_walk.load(std::memory_order_relaxed) ? 1 : 0;
        second.pf = lg_switches().tail_pf.load(std::memory_order_relaxed);
        const dim3 grid(nb, N / 4096, nlimbs);
        if (!inverse) {
            launch_strided_any(L, true, literal, first, grid, gs);
            launch_contig_pipe(true, literal, second, nlimbs, nb, gs);
        } else {
            launch_contig_pipe(false, literal, first, nlimbs, nb, gs);
            launch_strided_any(L, false, literal, second, grid, gs);
        }
        lg_g_launches += 2;
    }
    if (aux) {
        for (int k = 0; k < 2; ++k) {
            cudaEventRecord(aux->join[k], aux->s[k]);
            cudaStreamWaitEvent(st, aux->join[k], 0);
        }
    }
    return 0;
}

int lg_launch_ntt_fwd_strided(const NttArgs& args, int nlimbs, int batch, cudaStream_t st) {
    if (nlimbs <= 0 || batch <= 0) return 0;
    const u32 logN = args.T.logN, N = args.T.N;
    if (logN < 12 || logN > 16) return 1;
    NttArgs first = args;
    first.no_d64 = lg_switches().no_d64_ntt.load(std::memory_order_relaxed) ? 1 : 0;
    launch_strided_any((int)logN - 8, true, literal_ntt(), first, dim3(batch, N / 4096, nlimbs), st);
    lg_g_launches += 1;
    return 0;
}

int lg_launch_range_flags(const NttArgs& args, int nlimbs, int batch, u32* flags, cudaStream_t st) {
    if (nlimbs <= 0 || batch <= 0) return 0;
    cudaMemsetAsync(flags, 0, (size_t)nlimbs * sizeof(u32), st);
    range_flags_kernel<<<dim3(batch, (args.T.N + 4095) / 4096, nlimbs), 256, 0, st>>>(args, flags);
    lg_g_launches += 1;
    return 0;
}

int lg_encode_key_tensor_map(void* map, const u64* keyf, size_t words) {
    if (words % 16 != 0) return 1;
    return encode_words_2d(map, keyf, 16, words >> 4, 16, 32, true);  // a warp pair's 32 rows of a 2048-word tile
}

int lg_launch_ks_fused(const KsFusedArgs& a, int nlimbs, int batch, cudaStream_t st) {
    if (nlimbs <= 0 || batch <= 0) return 0;
    if (a.T.logN < 12 || a.T.logN > 16 || a.beta < 1 || nlimbs > LG_MAX_LIMBS) return 1;
    const u32 tiles = a.T.N / CONTIG_TILE;
    const size_t smem = KS_SMEM_WORDS * sizeof(u64);
    KsFusedArgs k = a;
    k.acc64 = lg_switches().ks_acc64.load(std::memory_order_relaxed) ? 1 : 0;
    k.no_d64 = lg_switches().no_d64_ntt.load(std::memory_order_relaxed) ? 1 : 0;
    k.pf = lg_switches().ks_key_pf.load(std::memory_order_relaxed);
    k.rev = lg_switches().reverse_walk.load(std::memory_order_relaxed) ? 1 : 0;  // the strided phase wrote the high limbs last
    k.use_zl = 0;
    const bool literal = literal_ntt();
    // FP64-class limbs whose key words are all canonical go to the TMA kernel, the others stay on ks_fused_kernel
    int nfp = 0, nint = 0;
    unsigned char zfp[LG_MAX_LIMBS], zint[LG_MAX_LIMBS];
    // (a single entry would leave half of every CTA idle: it stays on ks_fused_kernel)
    const bool tma = a.h_keymap && a.h_fp_ok && a.evk_f && !literal && !k.acc64 && !k.no_d64 && a.beta <= 32 && batch >= 2 && batch < (1 << 20) &&
                     (a.evk_ds % 16 == 0) && (a.evk_hs % 16 == 0) && !lg_switches().no_ks_tma.load(std::memory_order_relaxed);
    for (int j = 0; j < nlimbs; ++j) {
        if (tma && a.h_fp_ok[a.map(j)])
            zfp[nfp++] = (unsigned char)j;
        else
            zint[nint++] = (unsigned char)j;
    }
    // The two launches write disjoint limbs: when both exist the integer one goes to an auxiliary stream, forked from `st`
    // before the TMA launch and joined after, so that its CTAs fill the SMs as the TMA kernel drains (and the other way round).
    LgAux* aux = (nfp > 0 && nint > 0 && !lg_switches().no_aux_streams.load(std::memory_order_relaxed)) ? lg_aux_streams() : nullptr;
    cudaStream_t ks = st;
    if (aux) {
        lg_aux_fork(aux, st, 1);
        ks = aux->s[0];
    }
    if (nfp > 0) {
        KsFusedArgs f = k;
        f.rev = 0;
        f.use_zl = 1 | (batch << 8);
        memcpy(f.zl, zfp, sizeof(f.zl));
        CUtensorMap kmap;
        memcpy(&kmap, a.h_keymap, sizeof(kmap));
        lg_ensure_dyn_smem<ks_fused_tma_kernel>(KSF_SMEM_BYTES);
        ks_fused_tma_kernel<<<dim3((batch + KSF_GROUPS - 1) / KSF_GROUPS, tiles, nfp), KSF_THREADS, KSF_SMEM_BYTES, st>>>(f, kmap);
        lg_g_launches += 1;
        if (nint == 0) return 0;
        k.use_zl = 1;
        memcpy(k.zl, zint, sizeof(k.zl));
    }
    const dim3 grid(batch, tiles, nint);
    if (literal) {
        lg_ensure_dyn_smem<ks_fused_kernel<true>>(smem);
        ks_fused_kernel<true><<<grid, CONTIG_THREADS, smem, ks>>>(k);
    } else {
        lg_ensure_dyn_smem<ks_fused_kernel<false>>(smem);
        ks_fused_kernel<false><<<grid, CONTIG_THREADS, smem, ks>>>(k);
    }
    lg_g_launches += 1;
    if (aux) lg_aux_join(aux, st, 1);
    return 0;
}
