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
            const int tg = a.tgt0[k] + t;
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
                tg = a.tgt0[k] + o;
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
            const ulonglong2 val = *reinterpret_cast<const ulonglong2*>(in + (size_t)i * a.N);
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
