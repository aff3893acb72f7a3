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

}  // namespace

int lg_launch_ks_mac(const KsMacArgs& a, int nlimbs, int batch, cudaStream_t st) {
    if (nlimbs <= 0 || batch <= 0) return 0;
    u32 bx = ((a.T.N >> 1) + 255) / 256;
    if (bx > 64) bx = 64;
    if (bx == 0) bx = 1;
    ks_mac_kernel<<<dim3(bx, nlimbs, batch), 256, 0, st>>>(a);
    lg_g_launches += 1;
    return 0;
}
