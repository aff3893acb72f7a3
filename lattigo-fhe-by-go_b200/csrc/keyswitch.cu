// keyswitch.cu -- kernel family K3e: key-switch multiply-accumulate.
//
// Replaces the inner loop of ckks.evaluator.switchKeysInPlace
// (ckks/evaluator.go:1515-1541) and bfv.evaluator.switchKeys
// (bfv/evaluator.go:784-806): for one decomposition digit,
//   acc0 += MRed(evk[i][0], d),  acc1 += MRed(evk[i][1], d)
// over the active Q limbs and the special primes in ONE launch (the reference
// runs a Q pass per key half plus a hand-written P loop), with the lazy
// accumulators reduced by BRedAdd on the reference's cadence.
#include "common.cuh"
#include "kernels.h"

namespace {

__global__ void __launch_bounds__(256) ks_mac_kernel(const KsMacArgs a) {
    const int j = blockIdx.y, bt = blockIdx.z;
    const int tl = a.map(j);
    const LimbConst k = load_limb_const(a.T, tl);
    const u32 N = a.T.N;
    const ulonglong2* d = reinterpret_cast<const ulonglong2*>(a.d + bt * a.d_bs + (size_t)j * N);
    const ulonglong2* e0 = reinterpret_cast<const ulonglong2*>(a.evk0 + (size_t)tl * N);
    const ulonglong2* e1 = reinterpret_cast<const ulonglong2*>(a.evk1 + (size_t)tl * N);
    ulonglong2* p0 = reinterpret_cast<ulonglong2*>(a.acc0 + bt * a.acc_bs + (size_t)j * N);
    ulonglong2* p1 = reinterpret_cast<ulonglong2*>(a.acc1 + bt * a.acc_bs + (size_t)j * N);
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < (N >> 1); i += gridDim.x * blockDim.x) {
        const ulonglong2 dv = d[i];
        const ulonglong2 k0 = __ldg(e0 + i), k1 = __ldg(e1 + i);
        ulonglong2 r0 = make_ulonglong2(0, 0), r1 = make_ulonglong2(0, 0);
        if (!a.first) {
            r0 = p0[i];
            r1 = p1[i];
        }
        r0.x += mred(k0.x, dv.x, k.q, k.qinv);
        r0.y += mred(k0.y, dv.y, k.q, k.qinv);
        r1.x += mred(k1.x, dv.x, k.q, k.qinv);
        r1.y += mred(k1.y, dv.y, k.q, k.qinv);
        if (a.reduce) {
            r0.x = bred_add(r0.x, k.q, k.u0);
            r0.y = bred_add(r0.y, k.q, k.u0);
            r1.x = bred_add(r1.x, k.q, k.u0);
            r1.y = bred_add(r1.y, k.q, k.u0);
        }
        p0[i] = r0;
        p1[i] = r1;
    }
}

// Degree-1 x degree-1 tensor product of MulRelin (ckks/evaluator.go:1076-1095) in one pass:
//   c00 = MForm(a0), c01 = MForm(a1)
//   c0 = MRed(c00, b0); c1 = CRed(MRed(c00, b1) + MRed(c01, b0)); c2 = MRed(c01, b1)
// (squaring branch :1080-1085: c1 = CRed(2 * MRed(c00, b1))).  4 reads + 3 writes per coefficient
// instead of the reference's six passes.  Outputs may alias inputs (same-index access only).
__global__ void __launch_bounds__(256) tensor_kernel(const TensorArgs a) {
    const int j = blockIdx.y, bt = blockIdx.z;
    const int limb = a.limb0 + j * (a.lstep ? a.lstep : 1);
    const LimbConst k = load_limb_const(a.T, limb);
    const u32 N = a.T.N;
    const size_t off = (size_t)limb * N;
    const ulonglong2* a0 = reinterpret_cast<const ulonglong2*>(a.a0 + bt * a.a_bs[0] + off);
    const ulonglong2* a1 = reinterpret_cast<const ulonglong2*>(a.a1 + bt * a.a_bs[1] + off);
    const ulonglong2* b0 = reinterpret_cast<const ulonglong2*>(a.b0 + bt * a.b_bs[0] + off);
    const ulonglong2* b1 = reinterpret_cast<const ulonglong2*>(a.b1 + bt * a.b_bs[1] + off);
    ulonglong2* c0 = reinterpret_cast<ulonglong2*>(a.c0 + bt * a.c_bs[0] + off);
    ulonglong2* c1 = reinterpret_cast<ulonglong2*>(a.c1 + bt * a.c_bs[1] + off);
    ulonglong2* c2 = reinterpret_cast<ulonglong2*>(a.c2 + bt * a.c_bs[2] + (a.c2_compact ? (size_t)j * N : off));
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < (N >> 1); i += gridDim.x * blockDim.x) {
        const ulonglong2 x0 = a0[i], x1 = a1[i], y0 = b0[i], y1 = b1[i];
        ulonglong2 r0, r1, r2;
#define LG_TENSOR(f)                                                                \
    {                                                                               \
        const u64 m0 = mform(x0.f, k.q, k.u0, k.u1), m1 = mform(x1.f, k.q, k.u0, k.u1); \
        r0.f = mred(m0, y0.f, k.q, k.qinv);                                         \
        const u64 t = mred(m0, y1.f, k.q, k.qinv);                                  \
        const u64 s1 = a.square ? t + t : t + mred(m1, y0.f, k.q, k.qinv);              \
        r1.f = a.nomod ? s1 : cred(s1, k.q);                                        \
        r2.f = mred(m1, y1.f, k.q, k.qinv);                                         \
    }
        LG_TENSOR(x)
        LG_TENSOR(y)
#undef LG_TENSOR
        c0[i] = r0;
        c1[i] = r1;
        c2[i] = r2;
    }
}

// Hoisted key-switch multiply-accumulate: four consecutive outputs per thread (256-bit key loads), the
// digit values gathered through the index table.  The NTT-domain automorphism maps aligned blocks onto
// aligned blocks (the high bits of index[i] depend only on the high bits of i), so the gathers of a CTA
// stay inside one 8 KiB window of the source limb.
template <bool LAZYACC>
__device__ __forceinline__ void ks_hoisted_body(const KsHoistArgs& a, const LimbConst& k, int tl) {
    const int j = blockIdx.z, bt = blockIdx.x;
    const u32 N = a.T.N;
    const u32 e0 = 4 * (blockIdx.y * blockDim.x + threadIdx.x);
    if (e0 >= N) return;
    const uint4 ix = *reinterpret_cast<const uint4*>(a.index + e0);
    const u32 idx[4] = {ix.x, ix.y, ix.z, ix.w};
    const u64* d = a.D + (size_t)bt * a.d_bs + (size_t)j * N;
    const u64* key = a.evk + (size_t)tl * N + e0;
    u64 acc0[4] = {0, 0, 0, 0}, acc1[4] = {0, 0, 0, 0};
#pragma unroll 1
    for (int i = 0; i < a.beta; ++i, d += a.d_ds, key += a.evk_ds) {
        u64 x[4], k0[4], k1[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) x[e] = d[idx[e]];
        asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(k0[0]), "=l"(k0[1]), "=l"(k0[2]), "=l"(k0[3]) : "l"(key));
        asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(k1[0]), "=l"(k1[1]), "=l"(k1[2]), "=l"(k1[3])
                     : "l"(key + a.evk_hs));
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            if (LAZYACC) {
                acc0[e] += mred_constant(k0[e], x[e], k.q, k.qinv);
                acc1[e] += mred_constant(k1[e], x[e], k.q, k.qinv);
            } else {
                acc0[e] = cred(acc0[e] + mred(k0[e], x[e], k.q, k.qinv), k.q);
                acc1[e] = cred(acc1[e] + mred(k1[e], x[e], k.q, k.qinv), k.q);
            }
        }
    }
    u64* o0 = a.acc0 + (size_t)bt * a.acc_bs + (size_t)j * N + e0;
    u64* o1 = a.acc1 + (size_t)bt * a.acc_bs + (size_t)j * N + e0;
    asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(o0), "l"(bred_add(acc0[0], k.q, k.u0)),
                 "l"(bred_add(acc0[1], k.q, k.u0)), "l"(bred_add(acc0[2], k.q, k.u0)), "l"(bred_add(acc0[3], k.q, k.u0))
                 : "memory");
    asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(o1), "l"(bred_add(acc1[0], k.q, k.u0)),
                 "l"(bred_add(acc1[1], k.q, k.u0)), "l"(bred_add(acc1[2], k.q, k.u0)), "l"(bred_add(acc1[3], k.q, k.u0))
                 : "memory");
}

// The same on the FP64 pipe for moduli below 3*2^44 whose key limb is canonical (ntt.cu ACC_FP: the key out of Montgomery
// form as doubles, every term k*x brought to [0, 2q) by d64_mul, two double accumulators per coefficient, one canonical
// reduction at the end): 8 instead of about 29 instructions per term.  The digit values are canonical transforms (below q)
// except on the digit's own limbs, which are the caller's NTT-domain words: a thread that meets a word of 2^51 or more
// returns false and its four coefficients are redone on the integer path.  (Loading digit i+1 while digit i is accumulated --
// a register double buffer -- measured slower: 5990 against 6360 hoisted rotations/s, profiles/README.md.)
__device__ __forceinline__ bool ks_hoisted_body_fp(const KsHoistArgs& a, const LimbConst& k, int tl) {
    const int j = blockIdx.z, bt = blockIdx.x;
    const u32 N = a.T.N;
    const u32 e0 = 4 * (blockIdx.y * blockDim.x + threadIdx.x);
    if (e0 >= N) return true;
    const uint4 ix = *reinterpret_cast<const uint4*>(a.index + e0);
    const u32 idx[4] = {ix.x, ix.y, ix.z, ix.w};
    const u64* d = a.D + (size_t)bt * a.d_bs + (size_t)j * N;
    const u64* key = a.evk_f + (size_t)tl * N + e0;
    const double qd = __ull2double_rn(k.q), qinvd = __ddiv_rd(1.0, qd);
    double acc0[4] = {0.0, 0.0, 0.0, 0.0}, acc1[4] = {0.0, 0.0, 0.0, 0.0};
    u64 wide = 0;
#pragma unroll 1
    for (int i = 0; i < a.beta; ++i, d += a.d_ds, key += a.evk_ds) {
        u64 x[4], k0[4], k1[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) x[e] = d[idx[e]];
        asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];" : "=l"(k0[0]), "=l"(k0[1]), "=l"(k0[2]), "=l"(k0[3]) : "l"(key));
        asm volatile("ld.global.nc.v4.b64 {%0,%1,%2,%3}, [%4];"
                     : "=l"(k1[0]), "=l"(k1[1]), "=l"(k1[2]), "=l"(k1[3])
                     : "l"(key + a.evk_hs));
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            wide |= x[e];
            const double xd = u52_to_d(x[e] & 0x000FFFFFFFFFFFFFull, 4503599627370496.0);
            const double ka = bits2d(k0[e]), kb = bits2d(k1[e]);
            acc0[e] = __dadd_rn(acc0[e], d64_mul(ka, __dmul_rd(ka, qinvd), xd, qd));
            acc1[e] = __dadd_rn(acc1[e], d64_mul(kb, __dmul_rd(kb, qinvd), xd, qd));
        }
    }
    if (wide >> 51) return false;
    const double q34 = 34.0 * qd;
    u64 r0[4], r1[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {  // 0 <= sum < 2q*beta <= 64q: brought to [0, 2q) and then below q
        r0[e] = cred(d_to_u52(d64_red(acc0[e], qinvd, qd), 4503599627370496.0), k.q);
        r1[e] = cred(d_to_u52(d64_red(acc1[e], qinvd, qd), 4503599627370496.0), k.q);
    }
    (void)q34;
    u64* o0 = a.acc0 + (size_t)bt * a.acc_bs + (size_t)j * N + e0;
    u64* o1 = a.acc1 + (size_t)bt * a.acc_bs + (size_t)j * N + e0;
    asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(o0), "l"(r0[0]), "l"(r0[1]), "l"(r0[2]), "l"(r0[3]) : "memory");
    asm volatile("st.global.v4.b64 [%0], {%1,%2,%3,%4};" ::"l"(o1), "l"(r1[0]), "l"(r1[1]), "l"(r1[2]), "l"(r1[3]) : "memory");
    return true;
}

__global__ void __launch_bounds__(256) ks_hoisted_kernel(const KsHoistArgs a) {
    const int tl = a.map(blockIdx.z);
    const LimbConst k = load_limb_const(a.T, tl);
    if (a.evk_f != nullptr && k.q < (3ull << 44) && a.beta <= 32) {
        u32 bad = 0;
        for (int i = 0; i < 2 * a.beta; ++i) bad |= a.key_bad[(size_t)i * a.nqp + tl];
        if (bad == 0 && ks_hoisted_body_fp(a, k, tl)) return;
    }
    if ((2 * k.q) <= (~0ull) / (u64)a.beta)
        ks_hoisted_body<true>(a, k, tl);
    else
        ks_hoisted_body<false>(a, k, tl);
}

// FP64 form of a switching key (ntt.cu ACC_FP): plain key words as doubles for the FP64-class limbs
__global__ void __launch_bounds__(256) swk_prepare_kernel(RingTables T, const u64* key, u64* keyf, u32* bad, int nQP) {
    const int tl = blockIdx.y, dh = blockIdx.z;
    const u64 q = T.q[tl];
    if (q >= (3ull << 44)) return;
    const u64 qinv = T.qinv[tl];
    const size_t off = ((size_t)dh * nQP + tl) * T.N;
    int isbad = 0;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < T.N; i += gridDim.x * blockDim.x) {
        const u64 w = key[off + i];
        isbad |= w >= q;
        keyf[off + i] = (u64)__double_as_longlong(__ull2double_rn(mred(w, 1, q, qinv)));  // exact: below 2^46
    }
    if (__syncthreads_or(isbad) && threadIdx.x == 0) atomicOr(bad + (size_t)dh * nQP + tl, 1u);
}

}  // namespace

int lg_launch_swk_prepare(const RingTables& T, const u64* key, u64* keyf, u32* bad, int beta, int nQP, cudaStream_t st) {
    if (beta <= 0 || nQP <= 0) return 0;
    u32 bx = (T.N + 255) / 256;
    if (bx > 16) bx = 16;
    swk_prepare_kernel<<<dim3(bx, nQP, 2 * beta), 256, 0, st>>>(T, key, keyf, bad, nQP);
    lg_g_launches += 1;
    return 0;
}

int lg_launch_ks_hoisted(const KsHoistArgs& a, int nlimbs, int batch, cudaStream_t st) {
    if (nlimbs <= 0 || batch <= 0) return 0;
    if (a.T.N < 4 || a.beta < 1) return 1;
    const u32 threads = 256, per = 4 * threads;
    ks_hoisted_kernel<<<dim3(batch, (a.T.N + per - 1) / per, nlimbs), threads, 0, st>>>(a);
    lg_g_launches += 1;
    return 0;
}

int lg_launch_tensor(const TensorArgs& a, int nlimbs, int batch, cudaStream_t st) {
    if (nlimbs <= 0 || batch <= 0) return 0;
    u32 bx = ((a.T.N >> 1) + 255) / 256;
    if (bx > 64) bx = 64;
    if (bx == 0) bx = 1;
    tensor_kernel<<<dim3(bx, nlimbs, batch), 256, 0, st>>>(a);
    lg_g_launches += 1;
    return 0;
}

int lg_launch_ks_mac(const KsMacArgs& a, int nlimbs, int batch, cudaStream_t st) {
    if (nlimbs <= 0 || batch <= 0) return 0;
    u32 bx = ((a.T.N >> 1) + 255) / 256;
    if (bx > 64) bx = 64;
    if (bx == 0) bx = 1;
    ks_mac_kernel<<<dim3(bx, nlimbs, batch), 256, 0, st>>>(a);
    lg_g_launches += 1;
    return 0;
}
