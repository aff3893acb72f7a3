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
