// common.cuh -- shared device-side views used by every kernel family.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "modarith.cuh"

// Device-resident tables of one ring.Context (ring/ring_context.go:18-51),
// indexed by "table limb" = position of the prime in the context's modulus list.
struct RingTables {
    const u64* q;        // [nl]            Modulus
    const u64* qinv;     // [nl]            mredParams
    const u64* bred;     // [nl][2]         bredParams {hi, lo}
    const u64* psi;      // [nl][N]         nttPsi     (Montgomery, bit-reversed)
    const u64* psi_inv;  // [nl][N]         nttPsiInv
    const u64* ninv;     // [nl]            nttNInv
    const u64* psi_w;    // [nl][N]         nttPsi out of Montgomery form (same bit-reversed order)
    const u64* psi_ws;   // [nl][N]         floor(psi_w * 2^64 / q)  (Shoup constants of the fast forward NTT)
    const u64* psi_wd;   // [nl][N]         bits of the double RD(psi_ws) * 2^-64 <= psi_w / q (FP64-assisted forward NTT)
    const u64* psi_inv_w;   // [nl][N]      nttPsiInv out of Montgomery form
    const u64* psi_inv_ws;  // [nl][N]      its Shoup constants (fast inverse NTT)
    const u64* ninv_w;   // [nl][2]         {N^-1 mod q in plain form, its Shoup constant}
    // FP64-only transforms (moduli below 3*2^44, modarith.cuh): the same twiddles as doubles
    const u64* psi_wf;      // [nl][N]      bits of (double)psi_w
    const u64* psi_inv_wf;  // [nl][N]      bits of (double)psi_inv_w
    const u64* psi_inv_wd;  // [nl][N]      bits of RD(psi_inv_ws) * 2^-64
    const u64* ninv_f;      // [nl][2]      bits of {(double)ninv, RD(ninv / q)}
    u32 N;
    u32 logN;
    int nl;
    u64 d64_mask;  // bit tl set: table limb tl is below 3*2^44 (FP64-only transforms); host-side launch heuristics read it
};

// Maps the j-th data limb of a launch to a table limb: the first n0 data limbs
// use table limbs l0.., the rest use l1...  (Q limbs 0..level followed by the
// special primes at #Q.., as in ckks/evaluator.go:1519-1525.)
// With st > 1 consecutive data limbs are st table limbs apart: the cyclic limb ownership of the multi-GPU limb axis
// (rank r of w owns table limbs r, r+w, r+2w, ...: l0 = first own Q limb, l1 = first own special prime, st = w).
struct LimbMap {
    int n0, l0, l1;
    int st = 1;
    __host__ __device__ int operator()(int j) const { return j < n0 ? l0 + j * st : l1 + (j - n0) * st; }
};
static inline LimbMap limb_map_identity() { return LimbMap{1 << 30, 0, 0, 1}; }

__device__ __forceinline__ LimbConst load_limb_const(const RingTables& T, int tl) {
    LimbConst c;
    c.q = T.q[tl];
    c.qinv = T.qinv[tl];
    c.u0 = T.bred[2 * tl];
    c.u1 = T.bred[2 * tl + 1];
    return c;
}

#define LG_CUDA_CHECK(expr)                                                         \
    do {                                                                            \
        cudaError_t _e = (expr);                                                    \
        if (_e != cudaSuccess) {                                                    \
            lg_set_error("%s:%d: %s: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return LG_ERR_CUDA;                                                     \
        }                                                                           \
    } while (0)
