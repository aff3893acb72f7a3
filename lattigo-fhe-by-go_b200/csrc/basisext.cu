// basisext.cu -- kernel family K3b: exact RNS basis extension.
//
// Replaces modUpExact (ring/ring_basis_extension.go:352-393) and the identical
// inner loops of Decomposer.Decompose / DecomposeAndSplit (:476-713):
//   y_i = MRed(a_i, qibMont_i)            per source limb
//   v   = uint64( sum_i float64(y_i)/float64(q_i) )   sequential IEEE-RN sum
//   out_j = BRedAdd( sum_i MRed(y_i, qispjMont[i][j]) + qpjInv[j][v] )
// with a BRedAdd of the running sum whenever i&7 == 6.  The float64 part uses
// __ull2double_rn / __ddiv_rn / __dadd_rn in the reference's order (no FMA
// contraction, no reciprocal), and truncation toward zero for uint64(vi).
// One thread per coefficient; every load/store is coalesced across the warp.
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace {

template <int MAXSRC>
__global__ void __launch_bounds__(256) modup_kernel(const ModUpArgs a) {
    const u32 x = blockIdx.x * blockDim.x + threadIdx.x;
    const int bt = blockIdx.y;
    if (x >= a.N) return;
    const ModUpTables& M = a.M;
    u64 y[MAXSRC];
    double vi = 0.0;
    const u64* in = a.in + bt * a.in_bs + x;
#pragma unroll
    for (int i = 0; i < MAXSRC; ++i) {
        if (i < a.nsrc) {
            const u64 val = in[(size_t)i * a.N];
            if (a.copy_out) a.copy_out[bt * a.copy_bs + (size_t)i * a.N + x] = val;
            const u64 qi = __ldg(M.srcQ + i);
            y[i] = mred(val, __ldg(M.qib + i), qi, __ldg(M.srcQinv + i));
            vi = __dadd_rn(vi, __ddiv_rn(__ull2double_rn(y[i]), __ull2double_rn(qi)));
        }
    }
    const u64 v = __double2ull_rz(vi);
#pragma unroll 1
    for (int k = 0; k < a.nruns; ++k) {
        u64* out = a.out[k] + bt * a.out_bs[k] + x;
#pragma unroll 1
        for (int t = 0; t < a.ndst[k]; ++t) {
            const int tg = a.tgt0[k] + t * (a.tstep ? a.tstep : 1);
            const u64 pj = __ldg(M.dstQ + tg), pinv = __ldg(M.dstQinv + tg);
            // The reference sums canonical MRed(y_i, qispjMont[i][j]) terms (BRedAdd when i&7==6) and
            // canonicalises with a final BRedAdd (:379-389), i.e. it returns
            // (sum_i y_i*C_ij + qpjInv[j][v]) mod p_j in [0,p_j).  The same residue is obtained with
            // one Montgomery reduction per 8 sources of the 128-bit sum of y_i * qispjMont[i][j]
            // (< 8 * 2^61 * p_j, so REDC lands in [0,2p_j)) followed by conditional subtractions.
            u64 total = 0;
#pragma unroll
            for (int c0 = 0; c0 < MAXSRC; c0 += 8) {
                if (c0 < a.nsrc) {
                    unsigned __int128 S = 0;
#pragma unroll
                    for (int i = c0; i < c0 + 8 && i < MAXSRC; ++i)
                        if (i < a.nsrc) S += (unsigned __int128)y[i] * __ldg(M.qispj + (size_t)i * M.dst_total + tg);
                    const u64 shi = (u64)(S >> 64), slo = (u64)S;
                    const u64 r = shi - mul_hi(mul_lo(slo, pinv), pj) + pj;  // in [0, 2p_j]
                    total += cred(cred(r, pj), pj);
                }
            }
            total += __ldg(M.qpjinv + (size_t)tg * (M.src_total + 1) + v);
            if (MAXSRC <= 8)
                out[(size_t)t * a.N] = cred(total, pj);  // total < 2 p_j
            else
                out[(size_t)t * a.N] = bred_add(total, pj, __ldg(M.dstU0 + tg));
        }
    }
}

// Fast path for 1..4 source limbs with every modulus below 2^61 (all decompositions and ModDowns of the
// named parameter sets): two coefficients per thread (128-bit loads/stores), the per-target constants staged
// in shared memory, and the 128-bit sum of y_i * qispjMont[i][j] kept in column form
//   A0 = sum y0*c0 (64 bits + carry count), A1 = sum (y0*c1 + y1*c0) (2*NSRC terms below 2^61 fit 64 bits),
//   A2 = sum y1*c1
// i.e. four multiply-adds and one carry per term; qpjInv[j][v] joins the high word before the single
// Montgomery reduction.  Same residue as the reference's sum of MRed terms, canonicalised the same way.
template <int NSRC>
__global__ void __launch_bounds__(128) modup_fast_kernel(const ModUpArgs a) {
    constexpr int ROW = 2 * NSRC + 3;  // p, pinv, c[NSRC], qpjinv[NSRC+1]
    __shared__ u64 tab[LG_MAX_LIMBS * ROW];
    const ModUpTables& M = a.M;
    int ntg = 0;
    for (int k = 0; k < a.nruns; ++k) ntg += a.ndst[k];
    for (int e = threadIdx.x; e < ntg * ROW; e += blockDim.x) {
        int idx = e / ROW, f = e - idx * ROW, tg = 0;
        for (int k = 0, o = idx; k < a.nruns; ++k) {
            if (o < a.ndst[k]) {
                tg = a.tgt0[k] + o * (a.tstep ? a.tstep : 1);
                break;
            }
            o -= a.ndst[k];
        }
        u64 val;
        if (f == 0)
            val = M.dstQ[tg];
        else if (f == 1)
            val = M.dstQinv[tg];
        else if (f < 2 + NSRC)
            val = M.qispj[(size_t)(f - 2) * M.dst_total + tg];
        else
            val = M.qpjinv[(size_t)tg * (M.src_total + 1) + (f - 2 - NSRC)];
        tab[e] = val;
    }
    __syncthreads();
    const u32 x = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int bt = blockIdx.y;
    if (x >= a.N) return;
    u32 y0[NSRC][2], y1[NSRC][2];
    u32 v[2];
    {
        double vi0 = 0.0, vi1 = 0.0;
        const u64* in = a.in + bt * a.in_bs + x;
#pragma unroll
        for (int i = 0; i < NSRC; ++i) {
            const u64* srcp = a.src[0] ? a.src[i] + bt * a.in_bs + x : in + (size_t)i * a.N;
            const ulonglong2 val = *reinterpret_cast<const ulonglong2*>(srcp);
            if (a.copy_out) *reinterpret_cast<ulonglong2*>(a.copy_out + bt * a.copy_bs + (size_t)i * a.N + x) = val;
            const u64 qi = __ldg(M.srcQ + i), qib = __ldg(M.qib + i), qinv = __ldg(M.srcQinv + i);
            const u64 ya = mred(val.x, qib, qi, qinv), yb = mred(val.y, qib, qi, qinv);
            const double qd = __ull2double_rn(qi);
            vi0 = __dadd_rn(vi0, __ddiv_rn(__ull2double_rn(ya), qd));
            vi1 = __dadd_rn(vi1, __ddiv_rn(__ull2double_rn(yb), qd));
            y0[i][0] = (u32)ya;
            y1[i][0] = (u32)(ya >> 32);
            y0[i][1] = (u32)yb;
            y1[i][1] = (u32)(yb >> 32);
        }
        v[0] = (u32)__double2ull_rz(vi0);
        v[1] = (u32)__double2ull_rz(vi1);
    }
    int idx = 0;
#pragma unroll 1
    for (int k = 0; k < a.nruns; ++k) {
        u64* out = a.out[k] + bt * a.out_bs[k] + x;
#pragma unroll 1
        for (int t = 0; t < a.ndst[k]; ++t, ++idx) {
            const u64* row = tab + idx * ROW;
            const u64 pj = row[0], pinv = row[1];
            u64 res[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                u64 A0 = 0, A1 = 0, A2 = 0;
                u32 cnt = 0;
#pragma unroll
                for (int i = 0; i < NSRC; ++i) {
                    const u64 c = row[2 + i];
                    const u32 c0 = (u32)c, c1 = (u32)(c >> 32);
                    const u64 t0 = mul_wide(y0[i][e], c0);
                    asm("add.cc.u64 %0, %0, %2;\n\taddc.u32 %1, %1, 0;" : "+l"(A0), "+r"(cnt) : "l"(t0));
                    A1 = mad_wide(y0[i][e], c1, A1);
                    A1 = mad_wide(y1[i][e], c0, A1);
                    A2 = mad_wide(y1[i][e], c1, A2);
                }
                // S = A0 + (A1 << 32) + ((A2 + cnt) << 64), plus qpjInv[j][v] on the high word
                u64 slo, shi;
                asm("add.cc.u64 %0, %2, %3;\n\taddc.u64 %1, %4, %5;"
                    : "=l"(slo), "=l"(shi)
                    : "l"(A0), "l"(A1 << 32), "l"(A2 + cnt), "l"(A1 >> 32));
                shi += row[2 + NSRC + v[e]];
                const u64 r = shi - mul_hi(mul_lo(slo, pinv), pj) + pj;  // REDC: in (0, 3p)
                res[e] = cred(cred(r, pj), pj);
            }
            *reinterpret_cast<ulonglong2*>(out + (size_t)t * a.N) = make_ulonglong2(res[0], res[1]);
        }
    }
}

// The same column form for 5..MAXSRC source limbs (every modulus below 2^61): BFV's tensor product extends 12 limbs
// of Q to QMul and back (bfv/evaluator.go:300-372).  Sources are taken four at a time -- 8 cross terms below 2^61
// fit the middle column -- and every group of four is folded into a 128-bit running sum (nsrc * 2^122 < 2^128);
// one Montgomery reduction and a BRedAdd (the high word can exceed 2p) per target.  Constants sit in shared memory
// (row = p, pinv, bredParams[0], c[nsrc], qpjinv[nsrc+1]); two coefficients per thread, 128-bit access.
template <int MAXSRC>
__global__ void __launch_bounds__(128) modup_wide_kernel(const ModUpArgs a) {
    extern __shared__ u64 wtab[];
    const ModUpTables& M = a.M;
    const int nsrc = a.nsrc, ROW = 2 * nsrc + 4;
    int ntg = 0;
    for (int k = 0; k < a.nruns; ++k) ntg += a.ndst[k];
    for (int e = threadIdx.x; e < ntg * ROW; e += blockDim.x) {
        int idx = e / ROW, f = e - idx * ROW, tg = 0;
        for (int k = 0, o = idx; k < a.nruns; ++k) {
            if (o < a.ndst[k]) {
                tg = a.tgt0[k] + o * (a.tstep ? a.tstep : 1);
                break;
            }
            o -= a.ndst[k];
        }
        u64 val;
        if (f == 0)
            val = M.dstQ[tg];
        else if (f == 1)
            val = M.dstQinv[tg];
        else if (f == 2)
            val = M.dstU0[tg];
        else if (f < 3 + nsrc)
            val = M.qispj[(size_t)(f - 3) * M.dst_total + tg];
        else
            val = M.qpjinv[(size_t)tg * (M.src_total + 1) + (f - 3 - nsrc)];
        wtab[e] = val;
    }
    __syncthreads();
    const u32 x = 2 * (blockIdx.x * blockDim.x + threadIdx.x);
    const int bt = blockIdx.y;
    if (x >= a.N) return;
    u32 y0[MAXSRC][2], y1[MAXSRC][2];
    u32 v[2];
    {
        double vi0 = 0.0, vi1 = 0.0;
        const u64* in = a.in + bt * a.in_bs + x;
#pragma unroll
        for (int i = 0; i < MAXSRC; ++i) {
            y0[i][0] = y0[i][1] = y1[i][0] = y1[i][1] = 0;
            if (i < nsrc) {
                const ulonglong2 val = *reinterpret_cast<const ulonglong2*>(in + (size_t)i * a.N);
                if (a.copy_out) *reinterpret_cast<ulonglong2*>(a.copy_out + bt * a.copy_bs + (size_t)i * a.N + x) = val;
                const u64 qi = __ldg(M.srcQ + i), qib = __ldg(M.qib + i), qinv = __ldg(M.srcQinv + i);
                const u64 ya = mred(val.x, qib, qi, qinv), yb = mred(val.y, qib, qi, qinv);
                const double qd = __ull2double_rn(qi);
                vi0 = __dadd_rn(vi0, __ddiv_rn(__ull2double_rn(ya), qd));  // :363-375, sequential as in the reference
                vi1 = __dadd_rn(vi1, __ddiv_rn(__ull2double_rn(yb), qd));
                y0[i][0] = (u32)ya;
                y1[i][0] = (u32)(ya >> 32);
                y0[i][1] = (u32)yb;
                y1[i][1] = (u32)(yb >> 32);
            }
        }
        v[0] = (u32)__double2ull_rz(vi0);
        v[1] = (u32)__double2ull_rz(vi1);
    }
    int idx = 0;
#pragma unroll 1
    for (int k = 0; k < a.nruns; ++k) {
        u64* out = a.out[k] + bt * a.out_bs[k] + x;
#pragma unroll 1
        for (int t = 0; t < a.ndst[k]; ++t, ++idx) {
            const u64* row = wtab + idx * ROW;
            const u64 pj = row[0], pinv = row[1], u0 = row[2];
            u64 res[2];
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                u64 slo = 0, shi = 0;
#pragma unroll
                for (int g = 0; g < MAXSRC; g += 4) {
                    if (g < nsrc) {  // sources beyond nsrc hold zeros; their constants are read as zero below
                        u64 A0 = 0, A1 = 0, A2 = 0;
                        u32 cnt = 0;
#pragma unroll
                        for (int i = g; i < g + 4 && i < MAXSRC; ++i) {
                            const u64 c = i < nsrc ? row[3 + i] : 0ull;
                            const u32 c0 = (u32)c, c1 = (u32)(c >> 32);
                            const u64 t0 = mul_wide(y0[i][e], c0);
                            asm("add.cc.u64 %0, %0, %2;\n\taddc.u32 %1, %1, 0;" : "+l"(A0), "+r"(cnt) : "l"(t0));
                            A1 = mad_wide(y0[i][e], c1, A1);
                            A1 = mad_wide(y1[i][e], c0, A1);
                            A2 = mad_wide(y1[i][e], c1, A2);
                        }
                        // (shi:slo) += A0 + (A1 << 32) + ((A2 + cnt) << 64)
                        asm("add.cc.u64 %0, %0, %2;\n\taddc.u64 %1, %1, %3;\n\t"
                            "add.cc.u64 %0, %0, %4;\n\taddc.u64 %1, %1, %5;"
                            : "+l"(slo), "+l"(shi)
                            : "l"(A0), "l"(A2 + cnt), "l"(A1 << 32), "l"(A1 >> 32));
                    }
                }
                shi += row[3 + nsrc + v[e]];
                const u64 r = shi - mul_hi(mul_lo(slo, pinv), pj) + pj;  // REDC: congruent, below 2^64
                res[e] = bred_add(r, pj, u0);
            }
            *reinterpret_cast<ulonglong2*>(out + (size_t)t * a.N) = make_ulonglong2(res[0], res[1]);
        }
    }
}

// FP64-quotient path for 1..4 source limbs whose moduli sum to less than 2^48 (every digit of the CKKS scale primes),
// targets below 2^61.  The value to produce is R = (sum_i y_i*C_ij + qpjInv[j][v]) mod p_j with C_ij = Q/q_i mod p_j in
// plain form (the canonical residue the reference's chain of MRed terms and its final BRedAdd return, :379-389).
//   * low words:  S' = sum_i y_i*C_ij + K_v taken mod 2^64 -- one IMAD.WIDE and two IMADs per term;
//   * quotient:   every y_i < 2^48 is exact in binary64; with C_ij, K_v, 1/p_j rounded DOWN and every operation
//                 rounded down, t = RD(s * (1/p_j) + 2^52) holds an integer qh in its mantissa with
//                 S'/p_j - 1 - (sum_i q_i + 1) * 2^-49 < qh <= S'/p_j, so S' - qh*p_j lies in [0, 2*p_j)
//                 (nine downward roundings of relative size 2^-52 at most: 2^-49 on a value below sum_i q_i + 1);
//   * the product qh*p_j is taken mod 2^64 as bits(t) * (2^64 - p_j), the exponent bits of t being folded
//     into the K_v table entry, and one conditional subtraction canonicalises.
// 16 integer multiply-adds, 6 FP64 operations and one CRed per target coefficient instead of the 128-bit
// column sums and the Montgomery reduction of modup_fast_kernel.
// LAZY: the result is left in [0, 2p) (no conditional subtraction) -- for the key-switch digits, whose only reader
// is the forward NTT (exact for any input below its headroom).
// CPT coefficients per thread (256-bit access for CPT = 4): the per-target constants (interleaved so that one 128-bit
// shared-memory load brings a (C, Cd) or (K', Kd) pair) and the loop overhead are paid once per CPT coefficients --
// throughput follows the instruction count (profiles/r02_fp64_butterfly.txt): 29 instead of 41 per target coefficient.
template <int NSRC, bool LAZY, int CPT>
__global__ void __launch_bounds__(128) modup_fp_kernel(const ModUpArgs a) {
    // row: {np, p}, {1/p, -}, {C_i, Cd_i} x NSRC, {K'_v, Kd_v} x (NSRC+1)
    constexpr int ROW = 4 + 2 * NSRC + 2 * (NSRC + 1);
    constexpr int O_C = 4, O_K = 4 + 2 * NSRC;
    __shared__ __align__(16) u64 tab[LG_MAX_LIMBS * ROW];
    const ModUpTables& M = a.M;
    int ntg = 0;
    for (int k = 0; k < a.nruns; ++k) ntg += a.ndst[k];
    for (int idx = threadIdx.x; idx < ntg; idx += blockDim.x) {
        int tg = 0;
        for (int k = 0, o = idx; k < a.nruns; ++k) {
            if (o < a.ndst[k]) {
                tg = a.tgt0[k] + o * (a.tstep ? a.tstep : 1);
                break;
            }
            o -= a.ndst[k];
        }
        u64* row = tab + idx * ROW;
        const u64 p = M.dstQ[tg], pinv = M.dstQinv[tg];
        row[0] = 0 - p;
        row[1] = p;
        row[2] = (u64)__double_as_longlong(__ddiv_rd(1.0, __ull2double_ru(p)));
        row[3] = 0;
#pragma unroll
        for (int i = 0; i < NSRC; ++i) {
            const u64 c = mred(M.qispj[(size_t)i * M.dst_total + tg], 1, p, pinv);  // out of Montgomery form
            row[O_C + 2 * i] = c;
            row[O_C + 2 * i + 1] = (u64)__double_as_longlong(__ull2double_rd(c));
        }
#pragma unroll
        for (int v = 0; v <= NSRC; ++v) {
            const u64 kv = M.qpjinv[(size_t)tg * (M.src_total + 1) + v];
            row[O_K + 2 * v] = kv + 0x4330000000000000ull * p;  // + bits(2^52) * p: cancels the exponent field of t
            row[O_K + 2 * v + 1] = (u64)__double_as_longlong(__ull2double_rd(kv));
        }
    }
    __syncthreads();
    const u32 x = CPT * (blockIdx.x * blockDim.x + threadIdx.x);
    const int bt = blockIdx.y;
    if (x >= a.N) return;
    u32 y0[NSRC][CPT], y1[NSRC][CPT];
    double yd[NSRC][CPT];
    u32 v[CPT];
    {
        double vi[CPT];
#pragma unroll
        for (int e = 0; e < CPT; ++e) vi[e] = 0.0;
        const u64* in = a.in + bt * a.in_bs + x;
#pragma unroll
        for (int i = 0; i < NSRC; ++i) {
            const u64* srcp = a.src[0] ? a.src[i] + bt * a.in_bs + x : in + (size_t)i * a.N;
            u64 val[CPT];
#pragma unroll
            for (int h = 0; h < CPT / 2; ++h) {
                const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(srcp + 2 * h);
                val[2 * h] = t.x;
                val[2 * h + 1] = t.y;
            }
            if (a.copy_out) {
#pragma unroll
                for (int h = 0; h < CPT / 2; ++h)
                    *reinterpret_cast<ulonglong2*>(a.copy_out + bt * a.copy_bs + (size_t)i * a.N + x + 2 * h) =
                        make_ulonglong2(val[2 * h], val[2 * h + 1]);
            }
            const u64 qi = __ldg(M.srcQ + i), qib = __ldg(M.qib + i), qinv = __ldg(M.srcQinv + i);
            const double qd = __ull2double_rn(qi);
#pragma unroll
            for (int e = 0; e < CPT; ++e) {
                const u64 ye = mred(val[e], qib, qi, qinv);
                yd[i][e] = __ull2double_rn(ye);  // exact: below 2^48
                vi[e] = __dadd_rn(vi[e], __ddiv_rn(yd[i][e], qd));  // :363-375, sequential as in the reference
                y0[i][e] = (u32)ye;
                y1[i][e] = (u32)(ye >> 32);
            }
        }
#pragma unroll
        for (int e = 0; e < CPT; ++e) v[e] = (u32)__double2ull_rz(vi[e]);
    }
    int idx = 0;
#pragma unroll 1
    for (int k = 0; k < a.nruns; ++k) {
        u64* out = a.out[k] + bt * a.out_bs[k] + x;
#pragma unroll 1
        for (int t = 0; t < a.ndst[k]; ++t, ++idx) {
            const u64* row = tab + idx * ROW;
            const ulonglong2 h0 = *reinterpret_cast<const ulonglong2*>(row);      // {np, p}
            const u64 np = h0.x, pj = h0.y;
            const double pinvd = __longlong_as_double((long long)row[2]);
            u64 c[NSRC];
            double cd[NSRC];
#pragma unroll
            for (int i = 0; i < NSRC; ++i) {
                const ulonglong2 pr = *reinterpret_cast<const ulonglong2*>(row + O_C + 2 * i);
                c[i] = pr.x;
                cd[i] = __longlong_as_double((long long)pr.y);
            }
            u64 res[CPT];
#pragma unroll
            for (int e = 0; e < CPT; ++e) {
                const ulonglong2 kk = *reinterpret_cast<const ulonglong2*>(row + O_K + 2 * v[e]);  // {K'_v, Kd_v}
                double s = __dmul_rd(yd[0][e], cd[0]);
#pragma unroll
                for (int i = 1; i < NSRC; ++i) s = __fma_rd(yd[i][e], cd[i], s);
                s = __dadd_rd(s, __longlong_as_double((long long)kk.y));
                const u64 tb = (u64)__double_as_longlong(__fma_rd(s, pinvd, 4503599627370496.0));
                u64 acc = kk.x;
                u32 h = 0;
#pragma unroll
                for (int i = 0; i < NSRC; ++i) {
                    acc = mad_wide(y0[i][e], (u32)c[i], acc);
                    h = mad_lo32(y0[i][e], (u32)(c[i] >> 32), h);
                    h = mad_lo32(y1[i][e], (u32)c[i], h);
                }
                acc = mad_wide((u32)tb, (u32)np, acc);
                h = mad_lo32((u32)tb, (u32)(np >> 32), h);
                h = mad_lo32((u32)(tb >> 32), (u32)np, h);
                res[e] = LAZY ? acc + ((u64)h << 32) : cred(acc + ((u64)h << 32), pj);
            }
#pragma unroll
            for (int h = 0; h < CPT / 2; ++h)
                *reinterpret_cast<ulonglong2*>(out + (size_t)t * a.N + 2 * h) = make_ulonglong2(res[2 * h], res[2 * h + 1]);
        }
    }
}

// Two-step variant for wider sources (e.g. the 55-bit special primes of a ModDown, or a digit holding the first
// prime): sum_i q_i < 2^(52+SH).  The first quotient is taken at a granularity of 2^SH -- t = RD(s * (1/p) + 2^(52+SH))
// holds qh1 / 2^SH in its mantissa, s <= S being the all-rounded-down binary64 image of S = sum_i y_i*C_ij (the y_i
// themselves rounded down) -- so that r1 = S + K_v - qh1*p lies in [0, (2^SH + 8 * 2^-52 * sum_i q_i + 2) * p) (eight downward roundings), which
// the host checks to be below 2^64 for every target before choosing this kernel.  A second quotient on r1 (converted
// rounding down) leaves [0, 2p) and one conditional subtraction.  Everything integer is mod 2^64 as in modup_fp_kernel.
// CPT coefficients per thread (256-bit access for CPT = 4) and the per-target constants as 128-bit shared-memory pairs, as in
// modup_fp_kernel: the constant loads and the loop overhead are paid once per CPT coefficients.
template <int NSRC, bool LAZY, int CPT>
__global__ void __launch_bounds__(128) modup_fp2_kernel(const ModUpArgs a) {
    // row: {np << SH, p}, {1/p, bits(2^52) * p}, {np, -}, {C_i, Cd_i} x NSRC, K'[NSRC+1] (+ pad to an even count)
    constexpr int O_C = 6, O_K = 6 + 2 * NSRC;
    constexpr int ROW = (O_K + NSRC + 1 + 1) & ~1;
    __shared__ __align__(16) u64 tab[LG_MAX_LIMBS * ROW];
    const ModUpTables& M = a.M;
    const int sh = a.fp_shift;
    const u64 magic_bits = (u64)(0x433 + sh) << 52;  // 2^(52+sh)
    int ntg = 0;
    for (int k = 0; k < a.nruns; ++k) ntg += a.ndst[k];
    for (int idx = threadIdx.x; idx < ntg; idx += blockDim.x) {
        int tg = 0;
        for (int k = 0, o = idx; k < a.nruns; ++k) {
            if (o < a.ndst[k]) {
                tg = a.tgt0[k] + o * (a.tstep ? a.tstep : 1);
                break;
            }
            o -= a.ndst[k];
        }
        u64* row = tab + idx * ROW;
        const u64 p = M.dstQ[tg], pinv = M.dstQinv[tg];
        row[0] = (0 - p) << sh;
        row[1] = p;
        row[2] = (u64)__double_as_longlong(__ddiv_rd(1.0, __ull2double_ru(p)));
        row[3] = 0x4330000000000000ull * p;
        row[4] = 0 - p;
        row[5] = 0;
#pragma unroll
        for (int i = 0; i < NSRC; ++i) {
            const u64 c = mred(M.qispj[(size_t)i * M.dst_total + tg], 1, p, pinv);  // out of Montgomery form
            row[O_C + 2 * i] = c;
            row[O_C + 2 * i + 1] = (u64)__double_as_longlong(__ull2double_rd(c));
        }
#pragma unroll
        for (int v = 0; v <= NSRC; ++v)
            row[O_K + v] = M.qpjinv[(size_t)tg * (M.src_total + 1) + v] + magic_bits * (p << sh);
    }
    __syncthreads();
    const double magic = __longlong_as_double((long long)magic_bits);
    const u32 x = CPT * (blockIdx.x * blockDim.x + threadIdx.x);
    const int bt = blockIdx.y;
    if (x >= a.N) return;
    u32 y0[NSRC][CPT], y1[NSRC][CPT];
    double yd[NSRC][CPT];
    u32 v[CPT];
    {
        double vi[CPT];
#pragma unroll
        for (int e = 0; e < CPT; ++e) vi[e] = 0.0;
        const u64* in = a.in + bt * a.in_bs + x;
#pragma unroll
        for (int i = 0; i < NSRC; ++i) {
            const u64* srcp = a.src[0] ? a.src[i] + bt * a.in_bs + x : in + (size_t)i * a.N;
            u64 val[CPT];
#pragma unroll
            for (int h = 0; h < CPT / 2; ++h) {
                const ulonglong2 t = *reinterpret_cast<const ulonglong2*>(srcp + 2 * h);
                val[2 * h] = t.x;
                val[2 * h + 1] = t.y;
            }
            if (a.copy_out) {
#pragma unroll
                for (int h = 0; h < CPT / 2; ++h)
                    *reinterpret_cast<ulonglong2*>(a.copy_out + bt * a.copy_bs + (size_t)i * a.N + x + 2 * h) =
                        make_ulonglong2(val[2 * h], val[2 * h + 1]);
            }
            const u64 qi = __ldg(M.srcQ + i), qib = __ldg(M.qib + i), qinv = __ldg(M.srcQinv + i);
            const double qd = __ull2double_rn(qi);
#pragma unroll
            for (int e = 0; e < CPT; ++e) {
                const u64 ye = mred(val[e], qib, qi, qinv);
                vi[e] = __dadd_rn(vi[e], __ddiv_rn(__ull2double_rn(ye), qd));  // :363-375, sequential as in the reference
                yd[i][e] = __ull2double_rd(ye);
                y0[i][e] = (u32)ye;
                y1[i][e] = (u32)(ye >> 32);
            }
        }
#pragma unroll
        for (int e = 0; e < CPT; ++e) v[e] = (u32)__double2ull_rz(vi[e]);
    }
    int idx = 0;
#pragma unroll 1
    for (int k = 0; k < a.nruns; ++k) {
        u64* out = a.out[k] + bt * a.out_bs[k] + x;
#pragma unroll 1
        for (int t = 0; t < a.ndst[k]; ++t, ++idx) {
            const u64* row = tab + idx * ROW;
            const ulonglong2 h0 = *reinterpret_cast<const ulonglong2*>(row);      // {np << SH, p}
            const ulonglong2 h1 = *reinterpret_cast<const ulonglong2*>(row + 2);  // {1/p, bits(2^52) * p}
            const u64 npsh = h0.x, pj = h0.y, c52 = h1.y, np = row[4];
            const double pinvd = __longlong_as_double((long long)h1.x);
            u64 c[NSRC];
            double cd[NSRC];
#pragma unroll
            for (int i = 0; i < NSRC; ++i) {
                const ulonglong2 pr = *reinterpret_cast<const ulonglong2*>(row + O_C + 2 * i);
                c[i] = pr.x;
                cd[i] = __longlong_as_double((long long)pr.y);
            }
            u64 res[CPT];
#pragma unroll
            for (int e = 0; e < CPT; ++e) {
                double s = __dmul_rd(yd[0][e], cd[0]);
#pragma unroll
                for (int i = 1; i < NSRC; ++i) s = __fma_rd(yd[i][e], cd[i], s);
                const u64 tb = (u64)__double_as_longlong(__fma_rd(s, pinvd, magic));
                u64 acc = row[O_K + v[e]];
                u32 h = 0;
#pragma unroll
                for (int i = 0; i < NSRC; ++i) {
                    acc = mad_wide(y0[i][e], (u32)c[i], acc);
                    h = mad_lo32(y0[i][e], (u32)(c[i] >> 32), h);
                    h = mad_lo32(y1[i][e], (u32)c[i], h);
                }
                acc = mad_wide((u32)tb, (u32)npsh, acc);
                h = mad_lo32((u32)tb, (u32)(npsh >> 32), h);
                h = mad_lo32((u32)(tb >> 32), (u32)npsh, h);
                const u64 r1 = acc + ((u64)h << 32);
                const u64 t2 = (u64)__double_as_longlong(__fma_rd(__ull2double_rd(r1), pinvd, 4503599627370496.0));
                u64 r2 = mad_wide((u32)t2, (u32)np, r1 + c52);
                u32 h2 = mad_lo32((u32)t2, (u32)(np >> 32), (u32)(t2 >> 32) * (u32)np);
                r2 += (u64)h2 << 32;
                res[e] = LAZY ? r2 : cred(r2, pj);
            }
#pragma unroll
            for (int h = 0; h < CPT / 2; ++h)
                *reinterpret_cast<ulonglong2*>(out + (size_t)t * a.N + 2 * h) = make_ulonglong2(res[2 * h], res[2 * h + 1]);
        }
    }
}

__global__ void __launch_bounds__(256) fanout_kernel(const FanoutArgs a) {
    const u32 x = blockIdx.x * blockDim.x + threadIdx.x;
    const int bt = blockIdx.y;
    if (x >= a.N) return;
    u64 v = a.in[bt * a.in_bs + x];
    if (a.mode == 1) v = cred(v + a.phalf, a.plast);  // ring_scaling.go:83-88
    int idx = 0;
    for (int k = 0; k < a.nruns; ++k) {
        u64* out = a.out[k] + bt * a.out_bs[k] + x;
        for (int t = 0; t < a.ndst[k]; ++t, ++idx)
            out[(size_t)t * a.N] = (a.mode == 1) ? v + a.add[idx] : v;  // :99-103 (unreduced add)
    }
}

}  // namespace

int lg_launch_modup(const ModUpArgs& a, int batch, cudaStream_t st) {
    if (batch <= 0) return 0;
    const bool no_fp = lg_switches().no_fp_modup.load(std::memory_order_relaxed) != 0;
    const bool no_lazy = lg_switches().no_lazy_modup.load(std::memory_order_relaxed) != 0;
    const bool lazy = a.lazy_out && !no_lazy;
    if (a.fast == 2 && !no_fp && a.nsrc >= 1 && a.nsrc <= 4 && a.N >= 2) {
        // four coefficients per thread when that still fills the GPU (two otherwise, and for one or two sources: that
        // variant is bound by its HBM writes, 4.8 TB/s, and loses with the lower occupancy -- 166 against 120 us per digit)
        const bool four = !lg_switches().modup_cpt2.load(std::memory_order_relaxed) && a.N >= 4 && a.nsrc >= 3 &&
                          (size_t)batch * (a.N / 4 / 128) >= 2 * 148;
        const int cpt = four ? 4 : 2;
        dim3 fgrid((a.N / cpt + 127) / 128, batch);
#define LG_FP(NS)                                                                                  \
    if (four) {                                                                                    \
        if (lazy) modup_fp_kernel<NS, true, 4><<<fgrid, 128, 0, st>>>(a);                          \
        else modup_fp_kernel<NS, false, 4><<<fgrid, 128, 0, st>>>(a);                              \
    } else {                                                                                       \
        if (lazy) modup_fp_kernel<NS, true, 2><<<fgrid, 128, 0, st>>>(a);                          \
        else modup_fp_kernel<NS, false, 2><<<fgrid, 128, 0, st>>>(a);                              \
    }
        switch (a.nsrc) {
            case 1: LG_FP(1) break;
            case 2: LG_FP(2) break;
            case 3: LG_FP(3) break;
            default: LG_FP(4) break;
        }
#undef LG_FP
        lg_g_launches += 1;
        return 0;
    }
    if (a.fast == 3 && !no_fp && a.nsrc >= 1 && a.nsrc <= 4 && a.N >= 2) {
        const bool four = !lg_switches().modup_cpt2.load(std::memory_order_relaxed) && a.N >= 4 &&
                          (size_t)batch * (a.N / 4 / 128) >= 2 * 148;
        const int cpt = four ? 4 : 2;
        dim3 fgrid((a.N / cpt + 127) / 128, batch);
#define LG_FP2(NS)                                                                                 \
    if (four) {                                                                                    \
        if (lazy) modup_fp2_kernel<NS, true, 4><<<fgrid, 128, 0, st>>>(a);                         \
        else modup_fp2_kernel<NS, false, 4><<<fgrid, 128, 0, st>>>(a);                             \
    } else {                                                                                       \
        if (lazy) modup_fp2_kernel<NS, true, 2><<<fgrid, 128, 0, st>>>(a);                         \
        else modup_fp2_kernel<NS, false, 2><<<fgrid, 128, 0, st>>>(a);                             \
    }
        switch (a.nsrc) {
            case 1: LG_FP2(1) break;
            case 2: LG_FP2(2) break;
            case 3: LG_FP2(3) break;
            default: LG_FP2(4) break;
        }
#undef LG_FP2
        lg_g_launches += 1;
        return 0;
    }
    if (a.fast && a.nsrc >= 1 && a.nsrc <= 4 && a.N >= 2) {
        dim3 fgrid((a.N / 2 + 127) / 128, batch);
        switch (a.nsrc) {
            case 1: modup_fast_kernel<1><<<fgrid, 128, 0, st>>>(a); break;
            case 2: modup_fast_kernel<2><<<fgrid, 128, 0, st>>>(a); break;
            case 3: modup_fast_kernel<3><<<fgrid, 128, 0, st>>>(a); break;
            default: modup_fast_kernel<4><<<fgrid, 128, 0, st>>>(a); break;
        }
        lg_g_launches += 1;
        return 0;
    }
    const bool no_wide = lg_switches().no_wide_modup.load(std::memory_order_relaxed) != 0;
    if (a.fast && !no_wide && a.nsrc > 4 && a.nsrc <= 16 && a.N >= 2) {
        dim3 fgrid((a.N / 2 + 127) / 128, batch);
        int ntg = 0;
        for (int k = 0; k < a.nruns; ++k) ntg += a.ndst[k];
        const size_t smem = (size_t)ntg * (2 * a.nsrc + 4) * sizeof(u64);  // at most 64 * 36 words = 18 KiB
        if (a.nsrc <= 8)
            modup_wide_kernel<8><<<fgrid, 128, smem, st>>>(a);
        else if (a.nsrc <= 12)
            modup_wide_kernel<12><<<fgrid, 128, smem, st>>>(a);
        else
            modup_wide_kernel<16><<<fgrid, 128, smem, st>>>(a);
        lg_g_launches += 1;
        return 0;
    }
    dim3 grid((a.N + 255) / 256, batch);
    if (a.nsrc <= 2)
        modup_kernel<2><<<grid, 256, 0, st>>>(a);
    else if (a.nsrc <= 4)
        modup_kernel<4><<<grid, 256, 0, st>>>(a);
    else if (a.nsrc <= 8)
        modup_kernel<8><<<grid, 256, 0, st>>>(a);
    else if (a.nsrc <= 16)
        modup_kernel<16><<<grid, 256, 0, st>>>(a);
    else if (a.nsrc <= LG_MAX_LIMBS)
        modup_kernel<LG_MAX_LIMBS><<<grid, 256, 0, st>>>(a);
    else
        return 1;
    lg_g_launches += 1;
    return 0;
}

int lg_launch_fanout(const FanoutArgs& a, int batch, cudaStream_t st) {
    if (batch <= 0) return 0;
    dim3 grid((a.N + 255) / 256, batch);
    fanout_kernel<<<grid, 256, 0, st>>>(a);
    lg_g_launches += 1;
    return 0;
}
