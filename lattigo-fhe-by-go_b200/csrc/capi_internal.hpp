// capi_internal.hpp -- host-side objects behind the opaque C-ABI handles.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include <atomic>
#include <memory>
#include <mutex>
#include <vector>

#include "../../include/lattigpu.h"
#include "hostmath.hpp"
#include "kernels.h"

void lg_set_error(const char* fmt, ...);
extern std::atomic<uint64_t> lg_g_launches;

#define LG_REQUIRE(cond, ...)          \
    do {                               \
        if (!(cond)) {                 \
            lg_set_error(__VA_ARGS__); \
            return LG_ERR_ARG;         \
        }                              \
    } while (0)

#define LG_TRY(expr)              \
    do {                          \
        int _rc = (expr);         \
        if (_rc != LG_OK) return _rc; \
    } while (0)

// Every handle that owns or wraps device memory remembers its CUDA device; every entry point that touches the device
// runs under a DeviceGuard for it, so a host whose threads migrate (goroutines over OS threads, SURVEY.md 8(b)
// Threading) or that drives several GPUs from one process never launches on the wrong current device.  The previous
// device of the calling thread is restored on return.  device < 0 (host-only handle) is a no-op.
int lgi_current_device();
int lgi_pointer_device(const void* p);  // device of a cudaMalloc'ed pointer (cudaPointerGetAttributes), else the current one
int lgi_expected_device();  // device of the innermost DeviceGuard of this thread (-1: none): operand checks compare with it
struct DeviceGuard {
    int prev = -1;
    int outer = -1;
    bool switched = false;
    int rc = LG_OK;
    explicit DeviceGuard(int dev);
    ~DeviceGuard();
    DeviceGuard(const DeviceGuard&) = delete;
};
#define LG_ON_DEVICE(dev)           \
    DeviceGuard _lg_guard(dev);     \
    if (_lg_guard.rc != LG_OK) return _lg_guard.rc
// operands of one call must live on one device
#define LG_SAME_DEVICE(what, d0, d1) \
    LG_REQUIRE((d0) < 0 || (d1) < 0 || (d0) == (d1), "%s: operands live on different devices (%d and %d)", what, (int)(d0), (int)(d1))

#define LG_LAUNCH_CHECK()                                                               \
    do {                                                                                \
        cudaError_t _e = cudaPeekAtLastError();                                         \
        if (_e != cudaSuccess) {                                                        \
            lg_set_error("%s:%d: kernel launch: %s", __FILE__, __LINE__, cudaGetErrorString(_e)); \
            return LG_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

// device array owned by a handle
template <class T>
struct DevArray {
    T* d = nullptr;
    size_t n = 0;
    DevArray() {}
    DevArray(const DevArray&) = delete;
    DevArray& operator=(const DevArray&) = delete;
    ~DevArray() {
        if (d) cudaFree(d);
    }
    int upload(const std::vector<T>& h) {
        n = h.size();
        if (n == 0) return LG_OK;
        LG_CUDA_CHECK(cudaMalloc(&d, n * sizeof(T)));
        LG_CUDA_CHECK(cudaMemcpy(d, h.data(), n * sizeof(T), cudaMemcpyHostToDevice));
        return LG_OK;
    }
};

// stream-ordered scratch (cudaMallocAsync): safe for concurrent evaluators that
// share read-only ring handles, as the reference's goroutine pattern requires.
struct Scratch {
    u64* d = nullptr;
    cudaStream_t st;
    explicit Scratch(cudaStream_t s) : st(s) {}
    Scratch(const Scratch&) = delete;
    int alloc(size_t words) {
        LG_CUDA_CHECK(cudaMallocAsync((void**)&d, words * sizeof(u64), st));
        return LG_OK;
    }
    ~Scratch() {
        if (d) cudaFreeAsync(d, st);
    }
};

struct lg_ring {
    int device = -1;  // CUDA device of the tables
    u64 N = 0;
    u32 logN = 0;
    int nl = 0;
    std::vector<u64> q, bred, mred, ninv, psi, psi_inv, rescale;  // host copies
    DevArray<u64> d_q, d_qinv, d_bred, d_psi, d_psi_inv, d_ninv, d_psi_w, d_psi_ws, d_psi_inv_w, d_psi_inv_ws, d_ninv_w, d_psi_wd;
    DevArray<u64> d_psi_wf, d_psi_inv_wf, d_psi_inv_wd, d_ninv_f;  // FP64-only transforms
    // rescaleParams[j-1][i] and pHalfNegQi = q_i - ((q_j - 1)/2 mod q_i) (ring_scaling.go:82,:97), both triangular [j(j-1)/2 + i]
    DevArray<u64> d_rescale, d_phalfneg;
    RingTables T;
    u64 rescale_param(int j, int i) const { return rescale[(size_t)j * (j - 1) / 2 + i]; }  // rescaleParams[j-1][i]
};

struct lg_poly {
    int device = -1;
    u64* d = nullptr;
    u64 N = 0;
    int nlimbs = 0;
    int batch = 1;
    size_t bstride = 0;  // words
    bool owns = false;
};

struct lg_galois {
    int device = -1;
    u64 N = 0;
    std::vector<u64> index;
    DevArray<u32> d_index;
};

// device modupParams (ring_basis_extension.go:20-37)
struct ModUpDev {
    int nsrc = 0, ndst = 0;
    bool small = false;  // all moduli below 2^61
    std::vector<u64> hsrc;  // host copy of the source moduli
    std::vector<u64> hdst;  // ... and of the target moduli
    // kernel choice for the first `n` sources: 0 generic, 1 modup_fast_kernel (all moduli below 2^61), 2 modup_fp_kernel
    // (additionally the source moduli sum to less than 2^48), 3 modup_fp2_kernel with *shift = SH (sum below 2^(52+SH)
    // and the first remainder, below (2^SH + 8 * 2^-52 * sum + 2) * p, fits 64 bits for every target p); see basisext.cu
    int fast_level(int n, int* shift = nullptr) const {
        if (shift) *shift = 0;
        if (!small) return 0;
        if (n > 4) return 1;
        unsigned __int128 sum = 0;
        for (int i = 0; i < n && i < (int)hsrc.size(); ++i) sum += hsrc[i];
        if (sum < ((unsigned __int128)1 << 48)) return 2;
        int sh = 0;
        while (((sum + 1) >> (52 + sh)) != 0) ++sh;  // the first quotient, below sum + 1, must fit the mantissa at 2^sh
        if (sh > 11) return 1;
        const unsigned __int128 bound = ((unsigned __int128)1 << sh) + (sum >> 49) + 3;
        u64 pmax = 0;
        for (u64 p : hdst) pmax = p > pmax ? p : pmax;
        if (bound * pmax >= ((unsigned __int128)1 << 64)) return 1;
        if (shift) *shift = sh;
        return 3;
    }
    DevArray<u64> srcQ, srcQinv, qib, qispj, qpjinv, dstQ, dstQinv, dstU0;
    ModUpTables M;
    int build(const u64* Q, int nq, const u64* P, int np);
};

struct lg_extender {
    const lg_ring* Q = nullptr;
    const lg_ring* P = nullptr;
    ModUpDev qp, pq;
    std::vector<u64> moddown_pq;  // per Q limb: MForm(P^-1 mod q_i)
    DevArray<u64> d_moddown_pq;
    std::vector<u64> moddown_qp;  // per P limb: MForm(Q^-1 mod p_j)
};

struct lg_decomposer {
    int device = -1;
    u64 N = 0;
    int nQ = 0, nP = 0, alpha = 0, beta = 0;
    std::vector<int> xalpha;
    std::vector<std::vector<std::unique_ptr<ModUpDev>>> modup;  // [beta][xalpha-1]
};

struct lg_swk {
    int device = -1;
    u64* d = nullptr;
    u64 N = 0;
    int beta = 0, nQP = 0;
    bool owns = false;
    const u64* key(int digit, int half) const { return d + ((size_t)(digit * 2 + half) * nQP) * N; }
    // FP64 form for the limbs whose moduli take the FP64-only transforms (ntt.cu, ACC_FP): the key words out of Montgomery
    // form as doubles, same layout, and one flag per (digit, half, limb) raised when a word of that limb is not canonical
    // (such limbs keep the integer accumulators).  Built by lgi_swk_prepare on first use: a key is immutable once an
    // evaluator has used it -- lg_swk_invalidate after writing to it (lg_swk_poly invalidates: views exist to fill keys).
    mutable u64* d_f = nullptr;
    mutable u32* d_bad = nullptr;
    mutable bool prepared = false;
    // host side of the same: fp_ok[tl] = table limb tl may take ks_fused_tma_kernel (FP64-class modulus and no flag raised
    // for any (digit, half)), and the TMA descriptor of d_f (128 opaque bytes, a CUtensorMap; has_map = it could be encoded)
    mutable std::vector<unsigned char> fp_ok;
    alignas(64) mutable unsigned char keymap[128];
    mutable bool has_map = false;
    mutable std::mutex mu;
    ~lg_swk() {
        if (d_f) cudaFree(d_f);
        if (d_bad) cudaFree(d_bad);
    }
};
int lgi_swk_prepare(const lg_swk* k, const lg_ring* QP, cudaStream_t st);

// RotateHoisted precomputation: the NTT-domain digits of value[1] (ckks/evaluator.go:1258-1273)
struct lg_hoisted {
    int device = -1;
    u64* d = nullptr;  // [beta][batch][level+1+nP][N], allocated and freed in stream order on st
    cudaStream_t st = nullptr;
    u64 N = 0;
    int level = 0, beta = 0, batch = 0, nd = 0;
    size_t d_bs = 0, d_ds = 0;
};

struct lg_ckks_eval {
    const lg_ring* Q = nullptr;
    const lg_ring* P = nullptr;
    std::unique_ptr<lg_ring> QP;  // concatenated tables (ckks.go:73 contextQP)
    std::unique_ptr<lg_extender> ext;
    std::unique_ptr<lg_decomposer> dec;
    int alpha = 0;
};

// internal (non-ABI) helpers shared between translation units
int lgi_ring_build_device(lg_ring* r);
int lgi_ntt(const lg_ring* r, LimbMap map, int nl, int batch, const u64* in, size_t in_bs, u64* out, size_t out_bs,
            bool inverse, int skip0, int skip1, cudaStream_t st, bool in_range = false, const NttTail* tail = nullptr,
            const NttBcast* bcast = nullptr);
// the forward transform can carry the (x - y) * s tail of ModDown / rescaling in its last phase (logN >= 12)
bool lgi_ntt_tail_ok(const lg_ring* r);
int lgi_ew(int op, const lg_ring* r, LimbMap map, int nl, int batch, const u64* a, size_t a_bs, const u64* b,
           size_t b_bs, u64* c, size_t c_bs, const u64* scalars, int nscalars, cudaStream_t st);
int lgi_moddown_tail_ntt(const lg_extender* e, int level, int batch, const u64* p1Q, size_t p1Q_bs, u64* p1P,
                         size_t p1P_bs, u64* p2, size_t p2_bs, bool ntt, cudaStream_t st, bool accumulate = false,
                         bool p_in_range = false);
int lgi_moddown_pair_ntt(const lg_extender* e, int level, int batch, u64* acc0, u64* acc1, size_t acc_bs, int p_off, u64* out0,
                         size_t out0_bs, bool add0, u64* out1, size_t out1_bs, bool add1, bool ntt, cudaStream_t st,
                         bool p_in_range);
// lazy_out: the converted limbs may be left in [0, 2p) (callers whose only reader is the forward NTT)
int lgi_decompose(const lg_decomposer* d, int level, int crt, int batch, const u64* p0, size_t p0_bs, u64* outQ,
                  size_t outQ_bs, u64* outP, size_t outP_bs, cudaStream_t st, bool lazy_out = false);
int lgi_keyswitch_digits(const lg_ring* QP, const lg_ring* Q, LimbMap qp_map, const lg_decomposer* dec, int level, int beta,
                         int batch, const u64* coef, size_t coef_bs, const u64* nttd, size_t nttd_bs, const lg_swk* evk,
                         u64* d, u64* acc0, u64* acc1, size_t d_bs, int cadence, cudaStream_t st);
int lgi_concat_ring(const lg_ring* Q, const lg_ring* P, std::unique_ptr<lg_ring>& out);
int lgi_modup_launch(const ModUpDev& m, u64 N, int batch, const u64* in, size_t in_bs, int nsrc, u64* out, size_t out_bs,
                     int ndst, int tgt0, cudaStream_t st, bool lazy_out = false);
int lgi_div_by_last_modulus(const lg_ring* r, int nl, int batch, u64* p0, size_t bs, bool round, bool ntt,
                            cudaStream_t st);
