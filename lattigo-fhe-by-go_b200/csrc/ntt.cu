// ntt.cu -- kernel family K1: batched per-limb negacyclic NTT / InvNTT.
//
// Replaces ring/ntt.go:53-139 of the reference (Cooley-Tukey forward with lazy butterflies and a
// final BRedAdd; Gentleman-Sande inverse with a final MRed by N^-1).
//
// Bit-exactness.  InvNTT keeps every radix-2 butterfly literal (InvButterfly can wrap 64 bits on
// out-of-range input, so its result is formula-specific).  The forward transform of the reference never
// wraps (see modarith.cuh), so its output is the canonical transform of (x mod q) for EVERY 64-bit
// input; the forward kernels therefore use cheaper exact Shoup/Harvey butterflies (5-6 wide multiplies
// instead of 9) and still match the reference bit for bit, including on the unreduced words its own
// benchmarks feed (tests: "words" cases).  LATTIGPU_LITERAL_NTT=1 selects the literal forward
// butterflies instead (A/B and cross-check).
//
// Schedule (N = 2^logN, one limb = N words, grid = batch x tiles x limbs -- batch fastest, so the CTAs that
// share a limb's twiddles and key tile run together and hit L2):
//   logN <= 11 : one CTA per limb, radix-2 stages in shared memory.
//   logN >= 12 : two phases of register-resident radix-16 blocks
//     "strided" phase : the top L = logN-8 stages; a CTA owns all 2^L rows of
//                       W = 4096/2^L adjacent columns (coalesced 8*W-byte rows),
//     "contig"  phase : the low 8 stages on 16 contiguous 256-word segments.
//   Each thread keeps 16 coefficients in registers for 4 stages, then the CTA
//   re-distributes them through (padded, conflict-free) shared memory.
//   Twiddle index for the butterfly on (j, j+2^s): (N >> (s+1)) + (j >> (s+1))
//   in both directions (ring/ntt.go:74 and :120).
#include <stdlib.h>

#include "common.cuh"
#include "kernels.h"

namespace {

// forward butterfly flavours
enum { BF_LITERAL = 0, BF_4Q = 1, BF_FREE = 2 };

// ---- register blocks --------------------------------------------------------
// x[r] holds the coefficient at global index j0 + r*2^s (bits [s,s+4) of j0 are
// zero); twbase = (N + j0) >> s.  Stage u pairs r and r + 2^u (stride 2^(s+u)).

struct FwdConst {
    u64 q, qinv, twoq, fourq;
    const u64* tw;   // literal: nttPsi (Montgomery).  fast: psi_w
    const u64* tws;  // fast: psi_ws
};

template <bool VEC, int NG>
LG_DEV void load_tw(u64 (&w)[8], const u64* __restrict__ t, u32 base) {
    if (VEC && NG >= 2) {
#pragma unroll
        for (int g = 0; g < NG; g += 2) {
            const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(t + base + g));
            w[g] = v.x;
            w[g + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int g = 0; g < NG; ++g) w[g] = __ldg(t + base + g);
    }
}

template <int U, bool VEC, int MODE>
LG_DEV void fwd_stage(u64 (&x)[16], const FwdConst& c, u32 twbase) {
    constexpr int NG = 16 >> (U + 1);
    const u32 base = twbase >> (U + 1);
    u64 w[8], ws[8];
    load_tw<VEC, NG>(w, c.tw, base);
    if (MODE != BF_LITERAL) load_tw<VEC, NG>(ws, c.tws, base);
#pragma unroll
    for (int g = 0; g < NG; ++g) {
#pragma unroll
        for (int k = 0; k < (1 << U); ++k) {
            const int r = (g << (U + 1)) + k;
            if (MODE == BF_LITERAL)
                butterfly_fwd(x[r], x[r + (1 << U)], w[g], c.q, c.qinv, c.twoq);
            else if (MODE == BF_4Q)
                butterfly_fwd_4q(x[r], x[r + (1 << U)], w[g], ws[g], c.q, c.twoq);
            else
                butterfly_fwd_free(x[r], x[r + (1 << U)], w[g], ws[g], c.q, c.fourq);
        }
    }
}

// stages UHI, UHI-1, ..., ULO
template <int UHI, int ULO, bool VEC, int MODE>
LG_DEV void fwd_stages(u64 (&x)[16], const FwdConst& c, u32 twbase) {
    if (UHI >= 3 && ULO <= 3) fwd_stage<3, VEC, MODE>(x, c, twbase);
    if (UHI >= 2 && ULO <= 2) fwd_stage<2, VEC, MODE>(x, c, twbase);
    if (UHI >= 1 && ULO <= 1) fwd_stage<1, VEC, MODE>(x, c, twbase);
    if (UHI >= 0 && ULO <= 0) fwd_stage<0, VEC, MODE>(x, c, twbase);
}

template <int ULO, int UHI, bool VEC>
LG_DEV void inv_stages(u64 (&x)[16], const u64* __restrict__ tw, u32 twbase, u64 q, u64 qinv, u64 twoq) {
#pragma unroll
    for (int u = ULO; u <= UHI; ++u) {
        const u32 base = twbase >> (u + 1);
        const int ngroups = 16 >> (u + 1);
        u64 w[8];
        if (VEC && ngroups >= 2) {
#pragma unroll
            for (int g = 0; g < ngroups; g += 2) {
                const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(tw + base + g));
                w[g] = v.x;
                w[g + 1] = v.y;
            }
        } else {
#pragma unroll
            for (int g = 0; g < ngroups; ++g) w[g] = __ldg(tw + base + g);
        }
#pragma unroll
        for (int g = 0; g < ngroups; ++g) {
#pragma unroll
            for (int k = 0; k < (1 << u); ++k) {
                const int r = (g << (u + 1)) + k;
                butterfly_inv(x[r], x[r + (1 << u)], w[g], q, qinv, twoq);
            }
        }
    }
}

LG_DEV u32 pad16(u32 e) { return e + (e >> 4); }

struct LimbSetup {
    LimbConst c;
    const u64* in;
    u64* out;
    const u64* tw;
    u64 ninv;
    int tl;
    bool skip;
};

template <bool FWD>
LG_DEV LimbSetup setup_limb(const NttArgs& a) {
    LimbSetup s;
    const int j = blockIdx.z, b = blockIdx.x;
    s.skip = (j >= a.skip0 && j < a.skip1);
    s.tl = a.map(j);
    s.c = load_limb_const(a.T, s.tl);
    s.tw = (FWD ? a.T.psi : a.T.psi_inv) + (size_t)s.tl * a.T.N;
    s.ninv = FWD ? 0 : a.T.ninv[s.tl];
    s.in = a.in + (size_t)b * a.in_bstride + (size_t)j * a.T.N;
    s.out = a.out + (size_t)b * a.out_bstride + (size_t)j * a.T.N;
    return s;
}

template <int MODE>
LG_DEV FwdConst fwd_const(const NttArgs& a, const LimbSetup& s) {
    FwdConst c;
    c.q = s.c.q;
    c.qinv = s.c.qinv;
    c.twoq = 2 * s.c.q;
    c.fourq = 4 * s.c.q;
    if (MODE == BF_LITERAL) {
        c.tw = s.tw;
        c.tws = nullptr;
    } else {
        c.tw = a.T.psi_w + (size_t)s.tl * a.T.N;
        c.tws = a.T.psi_ws + (size_t)s.tl * a.T.N;
    }
    return c;
}

// which butterfly a limb may use: BF_FREE needs q < 2^56 (16 stages x 4q < 2^62), BF_4Q needs 4q < 2^64
LG_DEV int fast_mode(u64 q) { return q < (1ull << 56) ? BF_FREE : BF_4Q; }

// ---- forward, strided phase: stages 1..L -----------------------------------
template <int L, int MODE>
LG_DEV void fwd_strided_body(const NttArgs& a, const LimbSetup& s, u64* sm) {
    constexpr int G = 1 << (L - 4);  // threads per column
    constexpr int W = 256 / G;       // columns per CTA
    constexpr int N2 = L - 4;        // stages of the second register block
    const FwdConst c = fwd_const<MODE>(a, s);
    const u32 LB = a.T.logN - L;
    const int t = threadIdx.x, col = t % W, g = t / W;
    const size_t colg = (size_t)blockIdx.y * W + col;
    u64 x[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) x[r] = s.in[((size_t)(g + r * G) << LB) + colg];
    if (MODE == BF_FREE) {  // growth headroom: everything below 2^63 (canonical inputs never take this branch)
#pragma unroll
        for (int r = 0; r < 16; ++r)
            if (x[r] >> 63) x[r] = bred_add(x[r], c.q, s.c.u0);
    }
    if (MODE == BF_4Q) {  // Harvey's invariant: values in [0,4q)
#pragma unroll
        for (int r = 0; r < 16; ++r)
            if (x[r] >= c.fourq) x[r] = bred_add(x[r], c.q, s.c.u0);
    }
    fwd_stages<3, 0, false, MODE>(x, c, 16u);
    if (N2 > 0) {
#pragma unroll
        for (int r = 0; r < 16; ++r) sm[(g + r * G) * W + col] = x[r];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] = sm[(16 * g + r) * W + col];
        fwd_stages<(N2 > 0 ? N2 - 1 : 0), 0, false, MODE>(x, c, (1u << L) + 16u * g);
#pragma unroll
        for (int r = 0; r < 16; ++r) s.out[((size_t)(16 * g + r) << LB) + colg] = x[r];
    } else {
#pragma unroll
        for (int r = 0; r < 16; ++r) s.out[((size_t)(g + r * G) << LB) + colg] = x[r];
    }
}

template <int L, bool LITERAL>
__global__ void __launch_bounds__(256) ntt_fwd_strided(const NttArgs a) {
    __shared__ u64 sm[(L - 4) > 0 ? 4096 : 1];
    const LimbSetup s = setup_limb<true>(a);
    if (s.skip) return;
    if (LITERAL)
        fwd_strided_body<L, BF_LITERAL>(a, s, sm);
    else if (fast_mode(s.c.q) == BF_FREE)
        fwd_strided_body<L, BF_FREE>(a, s, sm);
    else
        fwd_strided_body<L, BF_4Q>(a, s, sm);
}

// ---- forward, contiguous phase: last 8 stages + BRedAdd ---------------------
template <int MODE>
LG_DEV void fwd_contig_body(const NttArgs& a, const LimbSetup& s, u64* sm) {
    const FwdConst c = fwd_const<MODE>(a, s);
    const u32 N = a.T.N;
    const u32 t = threadIdx.x, seg = t >> 4, cc = t & 15;
    const u32 base = blockIdx.y * 4096u;
    const u32 j0 = base + seg * 256u + cc;
    u64 x[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) x[r] = s.in[j0 + 16 * r];
    fwd_stages<3, 0, false, MODE>(x, c, (N + j0) >> 4);
#pragma unroll
    for (int r = 0; r < 16; ++r) sm[pad16(seg * 256u + cc + 16 * r)] = x[r];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; ++r) x[r] = sm[pad16(seg * 256u + 16 * cc + r)];
    const u32 j1 = base + seg * 256u + 16 * cc;
    fwd_stages<3, 0, true, MODE>(x, c, N + j1);
    // ring/ntt.go:83-85
#pragma unroll
    for (int r = 0; r < 16; ++r) sm[pad16(seg * 256u + 16 * cc + r)] = bred_add(x[r], c.q, s.c.u0);
    __syncthreads();
}

// MAC = true fuses the key-switch multiply-accumulate into the epilogue (NttMac).
template <bool MAC, bool LITERAL>
__global__ void __launch_bounds__(256) ntt_fwd_contig(const NttArgs a) {
    __shared__ u64 sm[4096 + 256];
    const LimbSetup s = setup_limb<true>(a);
    if (!MAC && s.skip) return;
    const u32 N = a.T.N;
    const u64 q = s.c.q, qinv = s.c.qinv;
    const u32 t = threadIdx.x;
    const u32 base = blockIdx.y * 4096u;
    if (!(MAC && s.skip)) {
        if (LITERAL)
            fwd_contig_body<BF_LITERAL>(a, s, sm);
        else if (fast_mode(q) == BF_FREE)
            fwd_contig_body<BF_FREE>(a, s, sm);
        else
            fwd_contig_body<BF_4Q>(a, s, sm);
    }
    if (!MAC) {
#pragma unroll
        for (int k = 0; k < 16; ++k) s.out[base + t + 256u * k] = sm[pad16(t + 256u * k)];
    } else {
        const int j = blockIdx.z, b = blockIdx.x;
        const size_t ko = (size_t)s.tl * N + base, ao = (size_t)b * a.mac.acc_bs + (size_t)j * N + base;
        const ulonglong2* e0 = reinterpret_cast<const ulonglong2*>(a.mac.evk0 + ko);
        const ulonglong2* e1 = reinterpret_cast<const ulonglong2*>(a.mac.evk1 + ko);
        ulonglong2* p0 = reinterpret_cast<ulonglong2*>(a.mac.acc0 + ao);
        ulonglong2* p1 = reinterpret_cast<ulonglong2*>(a.mac.acc1 + ao);
        const ulonglong2* cx = reinterpret_cast<const ulonglong2*>(a.mac.cx + (size_t)b * a.mac.cx_bs + (size_t)j * N + base);
#pragma unroll 4
        for (int k = 0; k < 8; ++k) {
            const u32 v = t + 256u * k;  // pair index: elements 2v, 2v+1
            ulonglong2 d;
            if (s.skip) {
                d = cx[v];
            } else {
                d.x = sm[pad16(2 * v)];
                d.y = sm[pad16(2 * v + 1)];
            }
            const ulonglong2 k0 = __ldg(e0 + v), k1 = __ldg(e1 + v);
            ulonglong2 r0, r1;
            r0.x = mred(k0.x, d.x, q, qinv);
            r0.y = mred(k0.y, d.y, q, qinv);
            r1.x = mred(k1.x, d.x, q, qinv);
            r1.y = mred(k1.y, d.y, q, qinv);
            if (!a.mac.first) {
                const ulonglong2 o0 = p0[v], o1 = p1[v];
                r0.x += o0.x;
                r0.y += o0.y;
                r1.x += o1.x;
                r1.y += o1.y;
            }
            if (a.mac.reduce) {
                r0.x = bred_add(r0.x, q, s.c.u0);
                r0.y = bred_add(r0.y, q, s.c.u0);
                r1.x = bred_add(r1.x, q, s.c.u0);
                r1.y = bred_add(r1.y, q, s.c.u0);
            }
            p0[v] = r0;
            p1[v] = r1;
        }
    }
}

// ---- inverse, contiguous phase: first 8 stages -------------------------------
__global__ void __launch_bounds__(256) ntt_inv_contig(const NttArgs a) {
    __shared__ u64 sm[4096 + 256];
    const LimbSetup s = setup_limb<false>(a);
    if (s.skip) return;
    const u32 N = a.T.N;
    const u64 q = s.c.q, qinv = s.c.qinv, twoq = 2 * s.c.q;
    const u32 t = threadIdx.x, seg = t >> 4, c = t & 15;
    const u32 base = blockIdx.y * 4096u;
#pragma unroll
    for (int k = 0; k < 16; ++k) sm[pad16(t + 256u * k)] = s.in[base + t + 256u * k];
    __syncthreads();
    u64 x[16];
#pragma unroll
    for (int r = 0; r < 16; ++r) x[r] = sm[pad16(seg * 256u + 16 * c + r)];
    const u32 j1 = base + seg * 256u + 16 * c;
    inv_stages<0, 3, true>(x, s.tw, N + j1, q, qinv, twoq);
#pragma unroll
    for (int r = 0; r < 16; ++r) sm[pad16(seg * 256u + 16 * c + r)] = x[r];
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 16; ++r) x[r] = sm[pad16(seg * 256u + c + 16 * r)];
    const u32 j0 = base + seg * 256u + c;
    inv_stages<0, 3, false>(x, s.tw, (N + j0) >> 4, q, qinv, twoq);
#pragma unroll
    for (int r = 0; r < 16; ++r) s.out[j0 + 16 * r] = x[r];
}

// ---- inverse, strided phase: last L stages + MRed by N^-1 --------------------
template <int L>
__global__ void __launch_bounds__(256) ntt_inv_strided(const NttArgs a) {
    constexpr int G = 1 << (L - 4);
    constexpr int W = 256 / G;
    constexpr int N2 = L - 4;
    __shared__ u64 sm[N2 > 0 ? 4096 : 1];
    const LimbSetup s = setup_limb<false>(a);
    if (s.skip) return;
    const u32 LB = a.T.logN - L;
    const u64 q = s.c.q, qinv = s.c.qinv, twoq = 2 * s.c.q;
    const int t = threadIdx.x, col = t % W, g = t / W;
    const size_t colg = (size_t)blockIdx.y * W + col;
    u64 x[16];
    if (N2 > 0) {
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] = s.in[((size_t)(16 * g + r) << LB) + colg];
        inv_stages<0, (N2 > 0 ? N2 - 1 : 0), false>(x, s.tw, (1u << L) + 16u * g, q, qinv, twoq);
#pragma unroll
        for (int r = 0; r < 16; ++r) sm[(16 * g + r) * W + col] = x[r];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] = sm[(g + r * G) * W + col];
    } else {
#pragma unroll
        for (int r = 0; r < 16; ++r) x[r] = s.in[((size_t)(g + r * G) << LB) + colg];
    }
    inv_stages<0, 3, false>(x, s.tw, 16u, q, qinv, twoq);
    // ring/ntt.go:136-138
#pragma unroll
    for (int r = 0; r < 16; ++r) s.out[((size_t)(g + r * G) << LB) + colg] = mred(x[r], s.ninv, q, qinv);
}

// ---- small rings (logN <= 11): one CTA per limb, radix-2 in shared memory ----
template <bool FWD>
__global__ void ntt_small(const NttArgs a) {
    extern __shared__ u64 dsm[];
    const LimbSetup s = setup_limb<FWD>(a);
    if (s.skip) return;
    const u32 N = a.T.N;
    const u64 q = s.c.q, qinv = s.c.qinv, twoq = 2 * s.c.q;
    for (u32 i = threadIdx.x; i < N; i += blockDim.x) dsm[i] = s.in[i];
    __syncthreads();
    if (FWD) {
        u32 sh = a.T.logN - 1;  // log2(t)
        for (u32 m = 1; m < N; m <<= 1, --sh) {
            for (u32 k = threadIdx.x; k < (N >> 1); k += blockDim.x) {
                const u32 i = k >> sh, jj = k & ((1u << sh) - 1);
                const u32 j = (i << (sh + 1)) + jj;
                u64 U = dsm[j], V = dsm[j + (1u << sh)];
                butterfly_fwd(U, V, s.tw[m + i], q, qinv, twoq);
                dsm[j] = U;
                dsm[j + (1u << sh)] = V;
            }
            __syncthreads();
        }
        for (u32 i = threadIdx.x; i < N; i += blockDim.x) s.out[i] = bred_add(dsm[i], q, s.c.u0);
    } else {
        u32 sh = 0;
        for (u32 h = N >> 1; h >= 1; h >>= 1, ++sh) {
            for (u32 k = threadIdx.x; k < (N >> 1); k += blockDim.x) {
                const u32 i = k >> sh, jj = k & ((1u << sh) - 1);
                const u32 j = (i << (sh + 1)) + jj;
                u64 U = dsm[j], V = dsm[j + (1u << sh)];
                butterfly_inv(U, V, s.tw[h + i], q, qinv, twoq);
                dsm[j] = U;
                dsm[j + (1u << sh)] = V;
            }
            __syncthreads();
        }
        for (u32 i = threadIdx.x; i < N; i += blockDim.x) s.out[i] = mred(dsm[i], s.ninv, q, qinv);
    }
}

template <int L>
void launch_strided(bool fwd, bool literal, const NttArgs& a, dim3 grid, cudaStream_t st) {
    if (!fwd)
        ntt_inv_strided<L><<<grid, 256, 0, st>>>(a);
    else if (literal)
        ntt_fwd_strided<L, true><<<grid, 256, 0, st>>>(a);
    else
        ntt_fwd_strided<L, false><<<grid, 256, 0, st>>>(a);
}

bool literal_forward() {
    static const bool v = [] {
        const char* e = getenv("LATTIGPU_LITERAL_NTT");
        return e && e[0] == '1';
    }();
    return v;
}

}  // namespace

int lg_launch_ntt(const NttArgs& args, int nlimbs, int batch, bool inverse, cudaStream_t st) {
    if (nlimbs <= 0 || batch <= 0) return 0;
    const u32 logN = args.T.logN, N = args.T.N;
    if (logN < 1 || logN > 16) return 1;
    if (args.mac.enabled && (logN <= 11 || inverse)) return 1;
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
    const bool literal = literal_forward();
    dim3 grid(batch, N / 4096, nlimbs);
    NttArgs second = args;  // the second phase runs in place on the output
    second.in = args.out;
    second.in_bstride = args.out_bstride;
    auto strided = [&](const NttArgs& a) {
        switch (L) {
            case 4: launch_strided<4>(!inverse, literal, a, grid, st); break;
            case 5: launch_strided<5>(!inverse, literal, a, grid, st); break;
            case 6: launch_strided<6>(!inverse, literal, a, grid, st); break;
            case 7: launch_strided<7>(!inverse, literal, a, grid, st); break;
            default: launch_strided<8>(!inverse, literal, a, grid, st); break;
        }
    };
    if (!inverse) {
        // with the MAC epilogue the strided phase runs in place on the input (the decomposed digit)
        NttArgs first = args;
        if (args.mac.enabled) {
            first.out = const_cast<u64*>(args.in);
            first.out_bstride = args.in_bstride;
            second.in = args.in;
            second.in_bstride = args.in_bstride;
        }
        strided(first);
        if (args.mac.enabled) {
            if (literal)
                ntt_fwd_contig<true, true><<<grid, 256, 0, st>>>(second);
            else
                ntt_fwd_contig<true, false><<<grid, 256, 0, st>>>(second);
        } else {
            if (literal)
                ntt_fwd_contig<false, true><<<grid, 256, 0, st>>>(second);
            else
                ntt_fwd_contig<false, false><<<grid, 256, 0, st>>>(second);
        }
    } else {
        ntt_inv_contig<<<grid, 256, 0, st>>>(args);
        strided(second);
    }
    lg_g_launches += 2;
    return 0;
}
