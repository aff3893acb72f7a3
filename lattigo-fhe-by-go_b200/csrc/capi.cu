// capi.cu -- C-ABI host layer, part 1: library, ring.Context, ring.Poly, NTT,
// coefficient-wise ops, Galois permutations and RNS rescaling.
// Host-side mirror of ring/ring_context.go, ring/ring_object.go, ring/ring.go,
// ring/ring_galois.go and ring/ring_scaling.go; all arithmetic on polynomial
// data runs in the CUDA kernels -- there is no CPU fallback.
#include <math.h>
#include <string.h>

#include <stdlib.h>

#include <mutex>

#include "capi_internal.hpp"

std::atomic<uint64_t> lg_g_launches{0};

static thread_local char g_err[512] = "";

void lg_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

static inline cudaStream_t cs(lg_stream_t s) { return (cudaStream_t)s; }

LgSwitches& lg_switches() {
    static LgSwitches sw;
    static std::once_flag once;
    std::call_once(once, [] {
        auto flag = [](const char* name, bool any_value) {
            const char* e = getenv(name);
            return (e && (any_value || e[0] == '1')) ? 1 : 0;
        };
        sw.literal_ntt = flag("LATTIGPU_LITERAL_NTT", false);
        sw.ks_acc64 = flag("LATTIGPU_KS_ACC64", false);
        sw.no_fp_mac = flag("LATTIGPU_NO_FP_MAC", true);
        sw.no_ks_tma = flag("LATTIGPU_NO_KS_TMA", true);
        sw.no_aux_streams = flag("LATTIGPU_NO_AUX_STREAMS", true);
        sw.no_strided_tma = flag("LATTIGPU_NO_STRIDED_TMA", true);
        if (const char* e = getenv("LATTIGPU_TILE_FASTEST")) sw.tile_fastest = atoi(e) ? 1 : 0;
        sw.no_d64_ntt = flag("LATTIGPU_NO_D64_NTT", true);
        sw.reverse_walk = flag("LATTIGPU_REVERSE_WALK", true);
        sw.no_fp_modup = flag("LATTIGPU_NO_FP_MODUP", true);
        sw.no_lazy_modup = flag("LATTIGPU_NO_LAZY_MODUP", true);
        sw.modup_cpt2 = flag("LATTIGPU_MODUP_CPT2", true);
        sw.no_wide_modup = flag("LATTIGPU_NO_WIDE_MODUP", true);
        sw.no_tail_canon = flag("LATTIGPU_NO_TAIL_CANON", true);
        sw.no_fused_tail = flag("LATTIGPU_NO_FUSED_TAIL", true);
        if (const char* e = getenv("LATTIGPU_KS_KEY_PF")) sw.ks_key_pf = atoi(e);
        if (const char* e = getenv("LATTIGPU_TAIL_PF")) sw.tail_pf = atoi(e);
        if (const char* e = getenv("LATTIGPU_KS_SCRATCH_WORDS")) sw.ks_scratch_words = strtoull(e, nullptr, 10);
        if (const char* e = getenv("LATTIGPU_NTT_L2_BYTES")) sw.ntt_l2_bytes = strtoull(e, nullptr, 10);
        if (const char* e = getenv("LATTIGPU_NTT_L2_STREAMS")) sw.ntt_l2_streams = atoi(e);
    });
    return sw;
}

int lgi_current_device() {
    int dev = -1;
    if (cudaGetDevice(&dev) != cudaSuccess) {
        cudaGetLastError();
        return -1;
    }
    return dev;
}
int lgi_pointer_device(const void* p) {
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) {
        cudaGetLastError();
        return lgi_current_device();
    }
    if (at.type == cudaMemoryTypeDevice || at.type == cudaMemoryTypeManaged) return at.device;
    return lgi_current_device();
}
static thread_local int g_expected_dev = -1;
int lgi_expected_device() { return g_expected_dev; }
// first use of a device: keep stream-ordered scratch cached in the pool instead of returning it to the driver
static void device_init_once(int dev) {
    static std::atomic<uint64_t> done{0};
    if (dev < 0 || dev >= 64 || (done.load(std::memory_order_acquire) >> dev & 1)) return;
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
        uint64_t thr = UINT64_MAX;
        cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    done.fetch_or(1ull << dev, std::memory_order_release);
}
DeviceGuard::DeviceGuard(int dev) {
    outer = g_expected_dev;
    if (dev < 0) return;
    g_expected_dev = dev;
    if (cudaGetDevice(&prev) != cudaSuccess) {
        lg_set_error("cudaGetDevice: %s", cudaGetErrorString(cudaGetLastError()));
        rc = LG_ERR_CUDA;
        return;
    }
    if (prev != dev) {
        if (cudaSetDevice(dev) != cudaSuccess) {
            lg_set_error("cudaSetDevice(%d): %s", dev, cudaGetErrorString(cudaGetLastError()));
            rc = LG_ERR_CUDA;
            return;
        }
        switched = true;
    }
}
DeviceGuard::~DeviceGuard() {
    g_expected_dev = outer;
    if (switched) cudaSetDevice(prev);
}

extern "C" {

const char* lg_last_error(void) { return g_err; }
int lg_debug_set_switch(const char* name, uint64_t value) {
    LG_REQUIRE(name, "lg_debug_set_switch: null name");
    LgSwitches& sw = lg_switches();
    const int v = value ? 1 : 0;
    if (!strcmp(name, "literal_ntt")) sw.literal_ntt = v;
    else if (!strcmp(name, "ks_acc64")) sw.ks_acc64 = v;
    else if (!strcmp(name, "no_fp_mac")) sw.no_fp_mac = v;
    else if (!strcmp(name, "no_ks_tma")) sw.no_ks_tma = v;
    else if (!strcmp(name, "no_aux_streams")) sw.no_aux_streams = v;
    else if (!strcmp(name, "no_strided_tma")) sw.no_strided_tma = v;
    else if (!strcmp(name, "tile_fastest")) sw.tile_fastest = v;
    else if (!strcmp(name, "no_d64_ntt")) sw.no_d64_ntt = v;
    else if (!strcmp(name, "reverse_walk")) sw.reverse_walk = v;
    else if (!strcmp(name, "no_fp_modup")) sw.no_fp_modup = v;
    else if (!strcmp(name, "no_lazy_modup")) sw.no_lazy_modup = v;
    else if (!strcmp(name, "modup_cpt2")) sw.modup_cpt2 = v;
    else if (!strcmp(name, "no_wide_modup")) sw.no_wide_modup = v;
    else if (!strcmp(name, "no_tail_canon")) sw.no_tail_canon = v;
    else if (!strcmp(name, "no_fused_tail")) sw.no_fused_tail = v;
    else if (!strcmp(name, "ks_key_pf")) sw.ks_key_pf = (int)value;
    else if (!strcmp(name, "tail_pf")) sw.tail_pf = (int)value;
    else if (!strcmp(name, "ks_scratch_words")) sw.ks_scratch_words = value ? value : ((uint64_t)6 << 27);
    else if (!strcmp(name, "ntt_l2_bytes")) sw.ntt_l2_bytes = value;
    else if (!strcmp(name, "ntt_l2_streams")) sw.ntt_l2_streams = (int)value;
    else {
        lg_set_error("lg_debug_set_switch: unknown switch '%s'", name);
        return LG_ERR_ARG;
    }
    return LG_OK;
}
const char* lg_version(void) { return "lattigpu 0.1 (sm_100a)"; }
uint64_t lg_launch_count(void) { return lg_g_launches.load(); }

int lg_device_count(int* count) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        cudaGetLastError();
        n = 0;
    }
    if (count) *count = n;
    if (n == 0) {
        lg_set_error("no CUDA device available (lattigpu has no CPU fallback)");
        return LG_ERR_NODEVICE;
    }
    return LG_OK;
}
int lg_set_device(int device) {
    LG_CUDA_CHECK(cudaSetDevice(device));
    device_init_once(device);
    return LG_OK;
}
int lg_stream_create(lg_stream_t* stream) {
    cudaStream_t s;
    LG_CUDA_CHECK(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (lg_stream_t)s;
    return LG_OK;
}
int lg_stream_destroy(lg_stream_t stream) {
    LG_CUDA_CHECK(cudaStreamDestroy(cs(stream)));
    return LG_OK;
}
int lg_stream_sync(lg_stream_t stream) {
    LG_CUDA_CHECK(cudaStreamSynchronize(cs(stream)));
    return LG_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// ring.Context
// ---------------------------------------------------------------------------

int lgi_ring_build_device(lg_ring* r) {
    std::vector<u64> qinv(r->mred);
    LG_TRY(r->d_q.upload(r->q));
    LG_TRY(r->d_qinv.upload(qinv));
    LG_TRY(r->d_bred.upload(r->bred));
    LG_TRY(r->d_psi.upload(r->psi));
    LG_TRY(r->d_psi_inv.upload(r->psi_inv));
    LG_TRY(r->d_ninv.upload(r->ninv));
    if (!r->rescale.empty()) {
        std::vector<u64> neg(r->rescale.size());
        for (int j = 1; j < r->nl; ++j)
            for (int i = 0; i < j; ++i) neg[(size_t)j * (j - 1) / 2 + i] = r->q[i] - (((r->q[j] - 1) >> 1) % r->q[i]);
        LG_TRY(r->d_rescale.upload(r->rescale));
        LG_TRY(r->d_phalfneg.upload(neg));
    }
    // twiddles of the fast transforms: psi / psi^-1 / N^-1 out of Montgomery form and their Shoup constants
    // (derived from the reference's tables, so the root choice stays the reference's)
    // RD(s) * 2^-64 as a double for s = floor(plain * 2^64 / q): never above plain/q, less than 2^-52 below it
    auto ratio_bits = [](u64 s) {
        double d = (double)s;
        if ((unsigned __int128)d > (unsigned __int128)s) d = nextafter(d, 0.0);
        d = ldexp(d, -64);
        u64 bits;
        memcpy(&bits, &d, sizeof(bits));
        return bits;
    };
    auto double_bits = [](u64 v) {  // exact: every table word is below 2^53 where this form is used
        const double d = (double)v;
        u64 bits;
        memcpy(&bits, &d, sizeof(bits));
        return bits;
    };
    auto shoup_tables = [&](const std::vector<u64>& mont, DevArray<u64>& dw, DevArray<u64>& dws, DevArray<u64>& dwd,
                            DevArray<u64>& dwf) -> int {
        std::vector<u64> w(mont.size()), ws(mont.size()), wd(mont.size()), wf(mont.size());
        const size_t per = mont.size() / (size_t)r->nl;
        for (int i = 0; i < r->nl; ++i) {
            const u64 qi = r->q[i], qi_inv = r->mred[i];
            for (size_t j = 0; j < per; ++j) {
                const u64 plain = lgh::mred(mont[(size_t)i * per + j], 1, qi, qi_inv);  // InvMForm
                w[(size_t)i * per + j] = plain;
                const u64 s = (u64)((((unsigned __int128)plain) << 64) / qi);
                ws[(size_t)i * per + j] = s;
                wd[(size_t)i * per + j] = ratio_bits(s);
                wf[(size_t)i * per + j] = double_bits(plain);
            }
        }
        LG_TRY(dw.upload(w));
        LG_TRY(dws.upload(ws));
        LG_TRY(dwd.upload(wd));
        LG_TRY(dwf.upload(wf));
        return LG_OK;
    };
    LG_TRY(shoup_tables(r->psi, r->d_psi_w, r->d_psi_ws, r->d_psi_wd, r->d_psi_wf));
    LG_TRY(shoup_tables(r->psi_inv, r->d_psi_inv_w, r->d_psi_inv_ws, r->d_psi_inv_wd, r->d_psi_inv_wf));
    {
        std::vector<u64> nw(2 * (size_t)r->nl), nf(2 * (size_t)r->nl);
        for (int i = 0; i < r->nl; ++i) {
            const u64 plain = lgh::mred(r->ninv[i], 1, r->q[i], r->mred[i]);
            nw[2 * i] = plain;
            nw[2 * i + 1] = (u64)((((unsigned __int128)plain) << 64) / r->q[i]);
            nf[2 * i] = double_bits(plain);
            nf[2 * i + 1] = ratio_bits(nw[2 * i + 1]);
        }
        LG_TRY(r->d_ninv_w.upload(nw));
        LG_TRY(r->d_ninv_f.upload(nf));
    }
    r->T.psi_wf = r->d_psi_wf.d;
    r->T.psi_inv_wf = r->d_psi_inv_wf.d;
    r->T.psi_inv_wd = r->d_psi_inv_wd.d;
    r->T.ninv_f = r->d_ninv_f.d;
    r->T.psi_w = r->d_psi_w.d;
    r->T.psi_ws = r->d_psi_ws.d;
    r->T.psi_wd = r->d_psi_wd.d;
    r->T.psi_inv_w = r->d_psi_inv_w.d;
    r->T.psi_inv_ws = r->d_psi_inv_ws.d;
    r->T.ninv_w = r->d_ninv_w.d;
    r->T.q = r->d_q.d;
    r->T.qinv = r->d_qinv.d;
    r->T.bred = r->d_bred.d;
    r->T.psi = r->d_psi.d;
    r->T.psi_inv = r->d_psi_inv.d;
    r->T.ninv = r->d_ninv.d;
    r->T.N = (u32)r->N;
    r->T.logN = r->logN;
    r->T.nl = r->nl;
    r->T.d64_mask = 0;
    for (int i = 0; i < r->nl && i < 64; ++i)
        if (r->q[i] < (3ull << 44)) r->T.d64_mask |= 1ull << i;
    return LG_OK;
}

static int ring_check_dims(uint64_t N, int nlimbs) {
    // ring_context.go:71-73 (panics when N is not a power of two)
    LG_REQUIRE(N >= 2 && (N & (N - 1)) == 0, "invalid ring degree %llu (must be a power of 2, >= 2)",
               (unsigned long long)N);
    LG_REQUIRE(N <= (1u << 16), "ring degree %llu above the supported maximum 2^16", (unsigned long long)N);
    LG_REQUIRE(nlimbs >= 1 && nlimbs <= LG_MAX_LIMBS, "number of moduli %d outside [1,%d]", nlimbs, LG_MAX_LIMBS);
    return LG_OK;
}

extern "C" {

int lg_ring_create(uint64_t N, int nlimbs, const uint64_t* moduli, lg_ring** out) {
    LG_REQUIRE(out && moduli, "null argument");
    LG_TRY(ring_check_dims(N, nlimbs));
    int ndev;
    LG_TRY(lg_device_count(&ndev));
    std::unique_ptr<lg_ring> r(new lg_ring);
    r->device = lgi_current_device();
    device_init_once(r->device);
    r->N = N;
    r->logN = lgh::log2u(N);
    r->nl = nlimbs;
    r->q.assign(moduli, moduli + nlimbs);
    r->bred.resize(2 * nlimbs);
    r->mred.resize(nlimbs);
    r->ninv.resize(nlimbs);
    r->psi.resize((size_t)nlimbs * N);
    r->psi_inv.resize((size_t)nlimbs * N);
    // GenNTTParams, ring_context.go:139-146
    for (int i = 0; i < nlimbs; ++i) {
        const u64 qi = moduli[i];
        if (!lgh::is_prime(qi) || (qi & ((N << 1) - 1)) != 1) {
            lg_set_error("warning : provided modulus does not allow NTT");
            return LG_ERR_NTT;
        }
    }
    for (int i = 0; i < nlimbs; ++i) {
        const u64 qi = moduli[i];
        lgh::bred_params(qi, r->bred[2 * i], r->bred[2 * i + 1]);
        r->mred[i] = lgh::mred_params(qi);
    }
    // rescaleParams, ring_context.go:148-158
    r->rescale.resize((size_t)nlimbs * (nlimbs - 1) / 2);
    for (int j = 1; j < nlimbs; ++j)
        for (int i = 0; i < j; ++i)
            r->rescale[(size_t)j * (j - 1) / 2 + i] =
                lgh::mform(lgh::powmod(moduli[j] % moduli[i], moduli[i] - 2, moduli[i]), moduli[i]);
    // psi tables, ring_context.go:168-204
    for (int i = 0; i < nlimbs; ++i) {
        const u64 qi = moduli[i], qinv = r->mred[i];
        r->ninv[i] = lgh::mform(lgh::powmod(N, qi - 2, qi), qi);
        const u64 g = lgh::primitive_root(qi);
        const u64 power = (qi - 1) / (N << 1);
        const u64 psi = lgh::mform(lgh::powmod(g, power, qi), qi);
        const u64 psi_inv = lgh::mform(lgh::powmod(g, (qi - 1) - power, qi), qi);
        u64* tp = r->psi.data() + (size_t)i * N;
        u64* ti = r->psi_inv.data() + (size_t)i * N;
        tp[0] = ti[0] = lgh::mform(1, qi);
        for (u64 j = 1; j < N; ++j) {
            const u64 prev = lgh::bitrev(j - 1, r->logN), next = lgh::bitrev(j, r->logN);
            tp[next] = lgh::mred(tp[prev], psi, qi, qinv);
            ti[next] = lgh::mred(ti[prev], psi_inv, qi, qinv);
        }
    }
    LG_TRY(lgi_ring_build_device(r.get()));
    *out = r.release();
    return LG_OK;
}

int lg_ring_create_from_tables(uint64_t N, int nlimbs, const uint64_t* moduli, const uint64_t* bred,
                               const uint64_t* mred, const uint64_t* psi, const uint64_t* psi_inv,
                               const uint64_t* ninv, const uint64_t* rescale, lg_ring** out) {
    LG_REQUIRE(out && moduli && bred && mred && psi && psi_inv && ninv, "null argument");
    LG_TRY(ring_check_dims(N, nlimbs));
    int ndev;
    LG_TRY(lg_device_count(&ndev));
    std::unique_ptr<lg_ring> r(new lg_ring);
    r->device = lgi_current_device();
    device_init_once(r->device);
    r->N = N;
    r->logN = lgh::log2u(N);
    r->nl = nlimbs;
    r->q.assign(moduli, moduli + nlimbs);
    r->bred.assign(bred, bred + 2 * nlimbs);
    r->mred.assign(mred, mred + nlimbs);
    r->ninv.assign(ninv, ninv + nlimbs);
    r->psi.assign(psi, psi + (size_t)nlimbs * N);
    r->psi_inv.assign(psi_inv, psi_inv + (size_t)nlimbs * N);
    const size_t nres = (size_t)nlimbs * (nlimbs - 1) / 2;
    if (rescale) r->rescale.assign(rescale, rescale + nres);
    LG_TRY(lgi_ring_build_device(r.get()));
    *out = r.release();
    return LG_OK;
}

int lg_ring_destroy(lg_ring* ring) {
    if (!ring) return LG_OK;
    LG_ON_DEVICE(ring->device);
    delete ring;
    return LG_OK;
}
uint64_t lg_ring_n(const lg_ring* ring) { return ring ? ring->N : 0; }
int lg_ring_nlimbs(const lg_ring* ring) { return ring ? ring->nl : 0; }

int lg_ring_get_tables(const lg_ring* r, uint64_t* moduli, uint64_t* bred, uint64_t* mred, uint64_t* psi,
                       uint64_t* psi_inv, uint64_t* ninv, uint64_t* rescale) {
    LG_REQUIRE(r, "null ring");
    if (moduli) memcpy(moduli, r->q.data(), r->q.size() * 8);
    if (bred) memcpy(bred, r->bred.data(), r->bred.size() * 8);
    if (mred) memcpy(mred, r->mred.data(), r->mred.size() * 8);
    if (psi) memcpy(psi, r->psi.data(), r->psi.size() * 8);
    if (psi_inv) memcpy(psi_inv, r->psi_inv.data(), r->psi_inv.size() * 8);
    if (ninv) memcpy(ninv, r->ninv.data(), r->ninv.size() * 8);
    if (rescale && !r->rescale.empty()) memcpy(rescale, r->rescale.data(), r->rescale.size() * 8);
    return LG_OK;
}

// host-only helpers of ring/utils.go
int lg_is_prime(uint64_t num) { return lgh::is_prime(num) ? 1 : 0; }
uint64_t lg_primitive_root(uint64_t q) { return lgh::primitive_root(q); }
int lg_generate_ntt_primes(uint64_t logQ, uint64_t logN, uint64_t levels, uint64_t* primes) {
    // ring/utils.go:133-175 ("logQ must be between 1 and 60" panics there)
    LG_REQUIRE(logQ >= 1 && logQ <= 60, "logQ must be between 1 and 60");
    LG_REQUIRE(primes || levels == 0, "null output");
    const u64 two_n = (u64)2 << logN;
    u64 x = ((u64)1 << logQ) + 1, y = x, n = 0;
    while (n < levels) {
        if (lgh::is_prime(x)) primes[n++] = x;
        x += two_n;
        if (n < levels && two_n > y) {  // :161-170 (unreachable for logQ > logN+1, kept for fidelity)
            y -= two_n;
            if (lgh::is_prime(y)) primes[n++] = y;
        }
    }
    return LG_OK;
}

// ---------------------------------------------------------------------------
// ring.Poly
// ---------------------------------------------------------------------------

int lg_poly_create(uint64_t N, int nlimbs, int batch, lg_poly** out) {
    LG_REQUIRE(out, "null argument");
    LG_REQUIRE(N >= 2 && (N & (N - 1)) == 0 && nlimbs >= 1 && batch >= 1, "invalid polynomial shape");
    std::unique_ptr<lg_poly> p(new lg_poly);
    p->device = lgi_current_device();
    p->N = N;
    p->nlimbs = nlimbs;
    p->batch = batch;
    p->bstride = (size_t)nlimbs * N;
    p->owns = true;
    const size_t bytes = (size_t)batch * p->bstride * sizeof(u64);
    LG_CUDA_CHECK(cudaMalloc((void**)&p->d, bytes));
    LG_CUDA_CHECK(cudaMemset(p->d, 0, bytes));
    *out = p.release();
    return LG_OK;
}
int lg_poly_wrap(void* device_ptr, uint64_t N, int nlimbs, int batch, lg_poly** out) {
    LG_REQUIRE(out && device_ptr, "null argument");
    LG_REQUIRE(((uintptr_t)device_ptr & 31) == 0, "device pointer must be 32-byte aligned");
    LG_REQUIRE(N >= 2 && (N & (N - 1)) == 0 && nlimbs >= 1 && batch >= 1, "invalid polynomial shape");
    lg_poly* p = new lg_poly;
    p->device = lgi_pointer_device(device_ptr);
    p->d = (u64*)device_ptr;
    p->N = N;
    p->nlimbs = nlimbs;
    p->batch = batch;
    p->bstride = (size_t)nlimbs * N;
    p->owns = false;
    *out = p;
    return LG_OK;
}
int lg_poly_view(const lg_poly* parent, int limb0, int nlimbs, lg_poly** out) {
    LG_REQUIRE(parent && out, "null argument");
    LG_REQUIRE(limb0 >= 0 && nlimbs >= 1 && limb0 + nlimbs <= parent->nlimbs, "view [%d,%d) outside %d limbs", limb0,
               limb0 + nlimbs, parent->nlimbs);
    lg_poly* p = new lg_poly(*parent);
    p->d = parent->d + (size_t)limb0 * parent->N;
    p->nlimbs = nlimbs;
    p->owns = false;
    *out = p;
    return LG_OK;
}
int lg_poly_destroy(lg_poly* p) {
    if (!p) return LG_OK;
    LG_ON_DEVICE(p->device);
    if (p->owns && p->d) cudaFree(p->d);
    delete p;
    return LG_OK;
}
uint64_t lg_poly_n(const lg_poly* p) { return p ? p->N : 0; }
int lg_poly_nlimbs(const lg_poly* p) { return p ? p->nlimbs : 0; }
int lg_poly_batch(const lg_poly* p) { return p ? p->batch : 0; }
void* lg_poly_device_ptr(const lg_poly* p) { return p ? p->d : nullptr; }
size_t lg_poly_batch_stride(const lg_poly* p) { return p ? p->bstride : 0; }

static int poly_xfer(const lg_poly* p, int batch0, int nbatch, int limb0, int nl, u64* host, bool up, cudaStream_t st,
                     bool sync = true) {
    LG_REQUIRE(p && host, "null argument");
    LG_REQUIRE(batch0 >= 0 && nbatch >= 1 && batch0 + nbatch <= p->batch, "batch range out of bounds");
    LG_REQUIRE(limb0 >= 0 && nl >= 1 && limb0 + nl <= p->nlimbs, "limb range out of bounds");
    LG_ON_DEVICE(p->device);
    const size_t row = (size_t)nl * p->N * sizeof(u64);
    u64* dev = p->d + (size_t)batch0 * p->bstride + (size_t)limb0 * p->N;
    if (up)
        LG_CUDA_CHECK(cudaMemcpy2DAsync(dev, p->bstride * sizeof(u64), host, row, row, nbatch, cudaMemcpyHostToDevice, st));
    else
        LG_CUDA_CHECK(cudaMemcpy2DAsync(host, row, dev, p->bstride * sizeof(u64), row, nbatch, cudaMemcpyDeviceToHost, st));
    if (sync) LG_CUDA_CHECK(cudaStreamSynchronize(st));  // cgo: the Go buffer may move after return
    return LG_OK;
}
int lg_poly_upload(lg_poly* p, int batch0, int nbatch, int limb0, int nl, const uint64_t* host, lg_stream_t s) {
    return poly_xfer(p, batch0, nbatch, limb0, nl, const_cast<u64*>(host), true, cs(s));
}
int lg_poly_download(const lg_poly* p, int batch0, int nbatch, int limb0, int nl, uint64_t* host, lg_stream_t s) {
    return poly_xfer(p, batch0, nbatch, limb0, nl, host, false, cs(s));
}
int lg_poly_upload_async(lg_poly* p, int batch0, int nbatch, int limb0, int nl, const uint64_t* host, lg_stream_t s) {
    return poly_xfer(p, batch0, nbatch, limb0, nl, const_cast<u64*>(host), true, cs(s), false);
}
int lg_poly_download_async(const lg_poly* p, int batch0, int nbatch, int limb0, int nl, uint64_t* host, lg_stream_t s) {
    return poly_xfer(p, batch0, nbatch, limb0, nl, host, false, cs(s), false);
}
// Wire format of ring.Poly (ring/ring_object.go:146-289): data[0] = log2(N), data[1] = number of moduli, then the
// coefficients limb-major as big-endian 64-bit words.  The byte swap runs on the device.
uint64_t lg_poly_get_data_len(const lg_poly* p, int nl, int with_metadata) {  // GetDataLen :186-192
    if (!p) return 0;
    return (uint64_t)(with_metadata ? 2 : 0) + (((uint64_t)nl * p->N) << 3);
}
int lg_poly_write_to(const lg_poly* p, int batch_index, int nl, uint8_t* data, uint64_t len, int with_metadata, lg_stream_t s) {
    LG_REQUIRE(p && data, "WriteTo: null argument");
    LG_REQUIRE(batch_index >= 0 && batch_index < p->batch, "WriteTo: batch index out of range");
    LG_REQUIRE(nl >= 0 && nl <= p->nlimbs && nl <= 255, "WriteTo: limb count out of range");
    LG_REQUIRE(len >= lg_poly_get_data_len(p, nl, with_metadata), "Data array is too small to write ring.Poly");  // :165-168
    LG_ON_DEVICE(p->device);
    size_t off = 0;
    if (with_metadata) {
        data[0] = (uint8_t)lgh::log2u(p->N);  // :169-170
        data[1] = (uint8_t)nl;
        off = 2;
    }
    const size_t words = (size_t)nl * p->N;
    if (words == 0) return LG_OK;
    Scratch tmp(cs(s));
    LG_TRY(tmp.alloc(words));
    lg_launch_bswap64(p->d + (size_t)batch_index * p->bstride, tmp.d, words, cs(s));
    LG_LAUNCH_CHECK();
    LG_CUDA_CHECK(cudaMemcpyAsync(data + off, tmp.d, words * sizeof(u64), cudaMemcpyDeviceToHost, cs(s)));
    LG_CUDA_CHECK(cudaStreamSynchronize(cs(s)));
    return LG_OK;
}
int lg_poly_decode(lg_poly* p, int batch_index, const uint8_t* data, uint64_t len, int with_metadata, int nl, lg_stream_t s) {
    LG_REQUIRE(p && data, "DecodePoly: null argument");
    LG_REQUIRE(batch_index >= 0 && batch_index < p->batch, "DecodePoly: batch index out of range");
    LG_ON_DEVICE(p->device);
    size_t off = 0;
    if (with_metadata) {  // UnmarshalBinary :257-274
        LG_REQUIRE(len >= 2, "error : invalid polynomial encoding");
        const uint64_t N = 1ull << (data[0] & 63);
        nl = data[1];
        LG_REQUIRE(((len - 2) >> 3) == N * (uint64_t)nl, "error : invalid polynomial encoding");
        LG_REQUIRE(N == p->N, "DecodePoly: encoded degree 2^%u differs from the polynomial's", (unsigned)data[0]);
        off = 2;
    } else {
        LG_REQUIRE(len >= (((uint64_t)nl * p->N) << 3), "DecodeCoeffs: data array too small");
    }
    LG_REQUIRE(nl >= 0 && nl <= p->nlimbs, "DecodePoly: %d moduli encoded, polynomial has %d", nl, p->nlimbs);
    const size_t words = (size_t)nl * p->N;
    if (words == 0) return LG_OK;
    Scratch tmp(cs(s));
    LG_TRY(tmp.alloc(words));
    LG_CUDA_CHECK(cudaMemcpyAsync(tmp.d, data + off, words * sizeof(u64), cudaMemcpyHostToDevice, cs(s)));
    lg_launch_bswap64(tmp.d, p->d + (size_t)batch_index * p->bstride, words, cs(s));
    LG_LAUNCH_CHECK();
    LG_CUDA_CHECK(cudaStreamSynchronize(cs(s)));
    return LG_OK;
}
int lg_poly_zero(lg_poly* p, lg_stream_t s) {
    LG_REQUIRE(p, "null argument");
    LG_ON_DEVICE(p->device);
    LG_CUDA_CHECK(cudaMemset2DAsync(p->d, p->bstride * sizeof(u64), 0, (size_t)p->nlimbs * p->N * sizeof(u64), p->batch,
                                    cs(s)));
    return LG_OK;
}
int lg_poly_copy(const lg_poly* src, int nl, lg_poly* dst, lg_stream_t s) {
    LG_REQUIRE(src && dst, "null argument");
    LG_REQUIRE(src->N == dst->N && src->batch == dst->batch, "shape mismatch");
    LG_REQUIRE(nl >= 1 && nl <= src->nlimbs && nl <= dst->nlimbs, "limb count out of range");
    if (src->d == dst->d) return LG_OK;  // ring_object.go:87 (p0 != p1)
    LG_SAME_DEVICE("Copy", src->device, dst->device);
    LG_ON_DEVICE(dst->device);
    const size_t row = (size_t)nl * src->N * sizeof(u64);
    LG_CUDA_CHECK(cudaMemcpy2DAsync(dst->d, dst->bstride * sizeof(u64), src->d, src->bstride * sizeof(u64), row, src->batch,
                                    cudaMemcpyDeviceToDevice, cs(s)));
    return LG_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// internal launch helpers
// ---------------------------------------------------------------------------

int lgi_ntt(const lg_ring* r, LimbMap map, int nl, int batch, const u64* in, size_t in_bs, u64* out, size_t out_bs,
            bool inverse, int skip0, int skip1, cudaStream_t st, bool in_range, const NttTail* tail,
            const NttBcast* bcast) {
    NttArgs a;
    memset(&a, 0, sizeof(a));
    if (tail) a.tail = *tail;
    if (bcast) a.bcast = *bcast;
    a.T = r->T;
    a.map = map;
    a.in = in;
    a.out = out;
    a.in_bstride = in_bs;
    a.out_bstride = out_bs;
    a.skip0 = skip0;
    a.skip1 = skip1;
    // The fast inverse butterflies equal the reference's only while no InvButterfly sum wraps, i.e. when
    // every input word is <= 2q; unless the caller vouches for that, flag the limbs that break it.
    Scratch flags(st);
    if (inverse && !in_range && r->logN >= 12 && nl > 0 && batch > 0) {
        LG_TRY(flags.alloc(((size_t)batch * nl + 1) / 2));
        lg_launch_range_flags(a, nl, batch, (u32*)flags.d, st);
        a.flags = (const u32*)flags.d;
    }
    if (lg_launch_ntt(a, nl, batch, inverse, st) != 0) {
        lg_set_error("NTT: unsupported ring degree 2^%u", r->logN);
        return LG_ERR_ARG;
    }
    LG_LAUNCH_CHECK();
    return LG_OK;
}

bool lgi_ntt_tail_ok(const lg_ring* r) {
    return !lg_switches().no_fused_tail.load(std::memory_order_relaxed) && r->logN >= 12;
}

int lgi_ew(int op, const lg_ring* r, LimbMap map, int nl, int batch, const u64* a, size_t a_bs, const u64* b,
           size_t b_bs, u64* c, size_t c_bs, const u64* scalars, int nscalars, cudaStream_t st) {
    EwArgs g;
    g.T = r->T;
    g.map = map;
    g.a = a;
    g.b = b;
    g.c = c;
    g.a_bs = a_bs;
    g.b_bs = b_bs;
    g.c_bs = c_bs;
    g.a_ls = g.b_ls = g.c_ls = r->N;
    for (int i = 0; i < nscalars && i < LG_MAX_LIMBS; ++i) g.s[i] = scalars[i];
    if (lg_launch_ew(op, g, nl, batch, st) != 0) {
        lg_set_error("elementwise: bad op %d", op);
        return LG_ERR_ARG;
    }
    LG_LAUNCH_CHECK();
    return LG_OK;
}

static int check_poly(const lg_ring* r, int nl, const lg_poly* p, const char* what) {
    LG_REQUIRE(p, "%s: null polynomial", what);
    LG_REQUIRE(p->N == r->N, "%s: polynomial degree %llu != ring degree %llu", what, (unsigned long long)p->N,
               (unsigned long long)r->N);
    LG_REQUIRE(nl <= p->nlimbs, "%s: %d limbs requested, polynomial has %d", what, nl, p->nlimbs);
    LG_SAME_DEVICE(what, r->device, p->device);
    return LG_OK;
}

static int ring_op3(int op, const lg_ring* r, int nl, const lg_poly* p1, const lg_poly* p2, lg_poly* p3,
                    lg_stream_t s, const char* what) {
    LG_REQUIRE(r, "%s: null ring", what);
    LG_REQUIRE(nl >= 1 && nl <= r->nl, "%s: %d limbs requested, ring has %d", what, nl, r->nl);
    LG_TRY(check_poly(r, nl, p1, what));
    if (p2) LG_TRY(check_poly(r, nl, p2, what));
    LG_TRY(check_poly(r, nl, p3, what));
    const int batch = p3->batch;
    LG_REQUIRE(p1->batch == batch || p1->batch == 1, "%s: batch mismatch", what);
    LG_REQUIRE(!p2 || p2->batch == batch || p2->batch == 1, "%s: batch mismatch", what);
    LG_ON_DEVICE(r->device);
    return lgi_ew(op, r, limb_map_identity(), nl, batch, p1->d, p1->batch == 1 && batch > 1 ? 0 : p1->bstride,
                  p2 ? p2->d : nullptr, p2 ? (p2->batch == 1 && batch > 1 ? 0 : p2->bstride) : 0, p3->d, p3->bstride,
                  nullptr, 0, cs(s));
}

static int ring_op_scalar(int op, const lg_ring* r, int nl, const lg_poly* p1, const u64* scalars, int nscalars,
                          lg_poly* p2, lg_stream_t s, const char* what) {
    LG_REQUIRE(r, "%s: null ring", what);
    LG_REQUIRE(nl >= 1 && nl <= r->nl, "%s: %d limbs requested, ring has %d", what, nl, r->nl);
    LG_TRY(check_poly(r, nl, p1, what));
    LG_TRY(check_poly(r, nl, p2, what));
    LG_REQUIRE(p1->batch == p2->batch, "%s: batch mismatch", what);
    LG_ON_DEVICE(r->device);
    return lgi_ew(op, r, limb_map_identity(), nl, p2->batch, p1->d, p1->bstride, nullptr, 0, p2->d, p2->bstride, scalars,
                  nscalars, cs(s));
}

extern "C" {

// ---------------------------------------------------------------------------
// NTT
// ---------------------------------------------------------------------------
static int ring_ntt(const lg_ring* r, int nl, const lg_poly* p1, lg_poly* p2, bool inv, lg_stream_t s) {
    LG_REQUIRE(r, "NTT: null ring");
    LG_REQUIRE(nl >= 1 && nl <= r->nl, "NTT: %d limbs requested, ring has %d", nl, r->nl);
    LG_TRY(check_poly(r, nl, p1, "NTT"));
    LG_TRY(check_poly(r, nl, p2, "NTT"));
    LG_REQUIRE(p1->batch == p2->batch, "NTT: batch mismatch");
    LG_ON_DEVICE(r->device);
    return lgi_ntt(r, limb_map_identity(), nl, p2->batch, p1->d, p1->bstride, p2->d, p2->bstride, inv, 0, 0, cs(s));
}
int lg_ring_ntt(const lg_ring* r, int nl, const lg_poly* p1, lg_poly* p2, lg_stream_t s) {
    return ring_ntt(r, nl, p1, p2, false, s);
}
int lg_ring_invntt(const lg_ring* r, int nl, const lg_poly* p1, lg_poly* p2, lg_stream_t s) {
    return ring_ntt(r, nl, p1, p2, true, s);
}
static int ring_ntt_limb(const lg_ring* r, int tl, const lg_poly* p1, int l1, lg_poly* p2, int l2, bool inv,
                         lg_stream_t s) {
    LG_REQUIRE(r && p1 && p2, "NTT: null argument");
    LG_REQUIRE(tl >= 0 && tl < r->nl, "NTT: table limb %d out of range", tl);
    LG_REQUIRE(l1 >= 0 && l1 < p1->nlimbs && l2 >= 0 && l2 < p2->nlimbs, "NTT: limb out of range");
    LG_REQUIRE(p1->N == r->N && p2->N == r->N && p1->batch == p2->batch, "NTT: shape mismatch");
    LG_SAME_DEVICE("NTT", r->device, p1->device);
    LG_SAME_DEVICE("NTT", r->device, p2->device);
    LG_ON_DEVICE(r->device);
    LimbMap m{1 << 30, tl, 0};
    return lgi_ntt(r, m, 1, p2->batch, p1->d + (size_t)l1 * r->N, p1->bstride, p2->d + (size_t)l2 * r->N, p2->bstride, inv,
                   0, 0, cs(s));
}
int lg_ring_ntt_limb(const lg_ring* r, int tl, const lg_poly* p1, int l1, lg_poly* p2, int l2, lg_stream_t s) {
    return ring_ntt_limb(r, tl, p1, l1, p2, l2, false, s);
}
int lg_ring_invntt_limb(const lg_ring* r, int tl, const lg_poly* p1, int l1, lg_poly* p2, int l2, lg_stream_t s) {
    return ring_ntt_limb(r, tl, p1, l1, p2, l2, true, s);
}

// ---------------------------------------------------------------------------
// coefficient-wise ops
// ---------------------------------------------------------------------------
#define LG_OP3(name, op)                                                                                          \
    int lg_ring_##name(const lg_ring* r, int nl, const lg_poly* p1, const lg_poly* p2, lg_poly* p3, lg_stream_t s) { \
        LG_REQUIRE(p2, #name ": null polynomial");                                                                 \
        return ring_op3(op, r, nl, p1, p2, p3, s, #name);                                                          \
    }
#define LG_OP2(name, op)                                                                       \
    int lg_ring_##name(const lg_ring* r, int nl, const lg_poly* p1, lg_poly* p2, lg_stream_t s) { \
        return ring_op3(op, r, nl, p1, nullptr, p2, s, #name);                                 \
    }
LG_OP3(add, EW_ADD)
LG_OP3(add_nomod, EW_ADD_NOMOD)
LG_OP3(sub, EW_SUB)
LG_OP3(sub_nomod, EW_SUB_NOMOD)
LG_OP2(neg, EW_NEG)
LG_OP2(reduce, EW_REDUCE)
LG_OP3(mul_coeffs, EW_MUL_BARRETT)
LG_OP3(mul_coeffs_and_add, EW_MUL_BARRETT_ADD)
LG_OP3(mul_coeffs_and_add_nomod, EW_MUL_BARRETT_ADD_NOMOD)
LG_OP3(mul_coeffs_constant, EW_MUL_BARRETT_CONSTANT)
LG_OP3(mul_coeffs_montgomery, EW_MULMONT)
LG_OP3(mul_coeffs_montgomery_and_add, EW_MULMONT_ADD)
LG_OP3(mul_coeffs_montgomery_and_add_nomod, EW_MULMONT_ADD_NOMOD)
LG_OP3(mul_coeffs_montgomery_constant_and_add_nomod, EW_MULMONT_CONSTANT_ADD_NOMOD)
LG_OP3(mul_coeffs_montgomery_and_sub, EW_MULMONT_SUB)
LG_OP3(mul_coeffs_montgomery_and_sub_nomod, EW_MULMONT_SUB_NOMOD)
LG_OP3(mul_coeffs_montgomery_constant, EW_MULMONT_CONSTANT)
LG_OP2(mform, EW_MFORM)
LG_OP2(invmform, EW_INVMFORM)

int lg_ring_mod(const lg_ring* r, int nl, const lg_poly* p1, uint64_t m, lg_poly* p2, lg_stream_t s) {
    LG_REQUIRE(m != 0, "Mod: zero modulus");
    u64 sc[2], lo;
    sc[0] = m;
    lgh::bred_params(m, sc[1], lo);
    return ring_op_scalar(EW_MOD, r, nl, p1, sc, 2, p2, s, "Mod");
}
int lg_ring_and(const lg_ring* r, int nl, const lg_poly* p1, uint64_t m, lg_poly* p2, lg_stream_t s) {
    return ring_op_scalar(EW_AND, r, nl, p1, &m, 1, p2, s, "AND");
}
int lg_ring_or(const lg_ring* r, int nl, const lg_poly* p1, uint64_t m, lg_poly* p2, lg_stream_t s) {
    return ring_op_scalar(EW_OR, r, nl, p1, &m, 1, p2, s, "OR");
}
int lg_ring_xor(const lg_ring* r, int nl, const lg_poly* p1, uint64_t m, lg_poly* p2, lg_stream_t s) {
    return ring_op_scalar(EW_XOR, r, nl, p1, &m, 1, p2, s, "XOR");
}
int lg_ring_add_scalar(const lg_ring* r, int nl, lg_poly* p1, const uint64_t* scalar, lg_stream_t s) {
    LG_REQUIRE(scalar, "AddScalar: null scalar");
    return ring_op_scalar(EW_ADD_SCALAR, r, nl, p1, scalar, nl, p1, s, "AddScalar");
}
int lg_ring_sub_scalar(const lg_ring* r, int nl, lg_poly* p1, const uint64_t* scalar, lg_stream_t s) {
    LG_REQUIRE(scalar, "SubScalar: null scalar");
    return ring_op_scalar(EW_SUB_SCALAR, r, nl, p1, scalar, nl, p1, s, "SubScalar");
}
int lg_ring_mul_scalar(const lg_ring* r, int nl, const lg_poly* p1, const uint64_t* scalar, lg_poly* p2, lg_stream_t s) {
    LG_REQUIRE(scalar, "MulScalar: null scalar");
    return ring_op_scalar(EW_MUL_SCALAR, r, nl, p1, scalar, nl, p2, s, "MulScalar");
}
static int ring_op_halves(int op, const lg_ring* r, int nl, const lg_poly* p1, const uint64_t* lo, const uint64_t* hi,
                          lg_poly* p2, lg_stream_t s, const char* what) {
    LG_REQUIRE(r && lo && hi, "%s: null argument", what);
    LG_REQUIRE(nl >= 1 && nl <= r->nl, "%s: %d limbs requested, ring has %d", what, nl, r->nl);
    LG_TRY(check_poly(r, nl, p1, what));
    LG_TRY(check_poly(r, nl, p2, what));
    LG_REQUIRE(p1->batch == p2->batch, "%s: batch mismatch", what);
    LG_ON_DEVICE(r->device);
    EwArgs g;
    g.T = r->T;
    g.map = limb_map_identity();
    g.a = p1->d;
    g.b = nullptr;
    g.c = p2->d;
    g.a_bs = p1->bstride;
    g.b_bs = 0;
    g.c_bs = p2->bstride;
    g.a_ls = g.b_ls = g.c_ls = r->N;
    for (int i = 0; i < nl; ++i) {
        g.s[i] = lo[i];
        g.shi[i] = hi[i];
    }
    if (lg_launch_ew(op, g, nl, p2->batch, cs(s)) != 0) {
        lg_set_error("%s: bad op", what);
        return LG_ERR_ARG;
    }
    LG_LAUNCH_CHECK();
    return LG_OK;
}
int lg_ring_add_scalar_halves(const lg_ring* r, int nl, const lg_poly* p1, const uint64_t* lo, const uint64_t* hi, lg_poly* p2,
                              lg_stream_t s) {
    return ring_op_halves(EW_ADD_SCALAR2, r, nl, p1, lo, hi, p2, s, "AddConst");
}
int lg_ring_mul_scalar_montgomery_halves(const lg_ring* r, int nl, const lg_poly* p1, const uint64_t* lo, const uint64_t* hi,
                                         lg_poly* p2, lg_stream_t s) {
    return ring_op_halves(EW_MUL_SCALAR_MONT2, r, nl, p1, lo, hi, p2, s, "MultByConst");
}
int lg_ring_mul_scalar_montgomery_halves_and_add(const lg_ring* r, int nl, const lg_poly* p1, const uint64_t* lo,
                                                 const uint64_t* hi, lg_poly* p2, lg_stream_t s) {
    return ring_op_halves(EW_MUL_SCALAR_MONT2_ADD, r, nl, p1, lo, hi, p2, s, "MultByConstAndAdd");
}
int lg_ring_mul_by_pow2(const lg_ring* r, int nl, const lg_poly* p1, uint64_t pow2, lg_poly* p2, lg_stream_t s) {
    // ring.go:629-653: MForm(p1, p2) followed by PowerOf2 of p1's words.  When
    // p1 != p2 the MForm result is overwritten, so only the aliased call sees it.
    LG_REQUIRE(pow2 < 64, "MulByPow2: shift %llu out of range", (unsigned long long)pow2);
    if (p1 && p2 && p1->d == p2->d) LG_TRY(ring_op3(EW_MFORM, r, nl, p1, nullptr, p2, s, "MulByPow2"));
    return ring_op_scalar(EW_MUL_POW2, r, nl, p1, &pow2, 1, p2, s, "MulByPow2");
}
int lg_ring_mul_by_vector_montgomery(const lg_ring* r, int nl, const lg_poly* p1, const lg_poly* vec, lg_poly* p2,
                                     lg_stream_t s) {
    LG_REQUIRE(r && vec && p1 && p2, "MulByVectorMontgomery: null argument");
    LG_TRY(check_poly(r, nl, p1, "MulByVectorMontgomery"));
    LG_TRY(check_poly(r, nl, p2, "MulByVectorMontgomery"));
    LG_REQUIRE(vec->N == r->N && nl <= r->nl && p1->batch == p2->batch, "MulByVectorMontgomery: shape mismatch");
    LG_SAME_DEVICE("MulByVectorMontgomery", r->device, vec->device);
    LG_ON_DEVICE(r->device);
    EwArgs g;
    g.T = r->T;
    g.map = limb_map_identity();
    g.a = p1->d;
    g.b = vec->d;
    g.c = p2->d;
    g.a_bs = p1->bstride;
    g.b_bs = 0;
    g.c_bs = p2->bstride;
    g.a_ls = g.c_ls = r->N;
    g.b_ls = 0;
    lg_launch_ew(EW_MULVEC, g, nl, p2->batch, cs(s));
    LG_LAUNCH_CHECK();
    return LG_OK;
}
int lg_ring_mul_by_vector_montgomery_and_add_nomod(const lg_ring* r, int nl, const lg_poly* p1, const lg_poly* vec,
                                                   lg_poly* p2, lg_stream_t s) {
    LG_REQUIRE(r && vec && p1 && p2, "MulByVectorMontgomeryAndAddNoMod: null argument");
    LG_TRY(check_poly(r, nl, p1, "MulByVectorMontgomeryAndAddNoMod"));
    LG_TRY(check_poly(r, nl, p2, "MulByVectorMontgomeryAndAddNoMod"));
    LG_REQUIRE(vec->N == r->N && nl <= r->nl && p1->batch == p2->batch, "shape mismatch");
    LG_SAME_DEVICE("MulByVectorMontgomeryAndAddNoMod", r->device, vec->device);
    LG_ON_DEVICE(r->device);
    EwArgs g;
    g.T = r->T;
    g.map = limb_map_identity();
    g.a = p1->d;
    g.b = vec->d;
    g.c = p2->d;
    g.a_bs = p1->bstride;
    g.b_bs = 0;
    g.c_bs = p2->bstride;
    g.a_ls = g.c_ls = r->N;
    g.b_ls = 0;
    lg_launch_ew(EW_MULVEC_ADD_NOMOD, g, nl, p2->batch, cs(s));
    LG_LAUNCH_CHECK();
    return LG_OK;
}

// ---------------------------------------------------------------------------
// permutations
// ---------------------------------------------------------------------------
static int make_galois(std::vector<u64>&& idx, u64 N, lg_galois** out) {
    std::unique_ptr<lg_galois> g(new lg_galois);
    g->device = lgi_current_device();
    g->N = N;
    g->index = std::move(idx);
    std::vector<u32> i32(N);
    for (u64 i = 0; i < N; ++i) {
        LG_REQUIRE(g->index[i] < N, "Galois index %llu out of range", (unsigned long long)g->index[i]);
        i32[i] = (u32)g->index[i];
    }
    LG_TRY(g->d_index.upload(i32));
    *out = g.release();
    return LG_OK;
}
static std::vector<u64> permute_ntt_index(u64 genpow, u64 N) {
    // ring_galois.go:37-48
    const unsigned logN = lgh::log2u(N);
    const u64 mask = (N << 1) - 1;
    std::vector<u64> idx(N);
    for (u64 i = 0; i < N; ++i) {
        const u64 t1 = 2 * lgh::bitrev(i, logN) + 1;
        const u64 t2 = (((genpow * t1) & mask) - 1) >> 1;
        idx[i] = lgh::bitrev(t2, logN);
    }
    return idx;
}
int lg_galois_create(uint64_t gen, uint64_t power, uint64_t N, lg_galois** out) {
    LG_REQUIRE(out && N >= 2 && (N & (N - 1)) == 0, "PermuteNTTIndex: invalid argument");
    return make_galois(permute_ntt_index(lgh::powmod(gen, power, 2 * N), N), N, out);  // :31 ModExp(gen, power, 2N)
}
int lg_galois_create_from_index(const uint64_t* index, uint64_t N, lg_galois** out) {
    LG_REQUIRE(out && index && N >= 2, "invalid argument");
    return make_galois(std::vector<u64>(index, index + N), N, out);
}
int lg_galois_get_index(const lg_galois* g, uint64_t* index) {
    LG_REQUIRE(g && index, "null argument");
    memcpy(index, g->index.data(), g->N * 8);
    return LG_OK;
}
int lg_galois_destroy(lg_galois* g) {
    if (!g) return LG_OK;
    LG_ON_DEVICE(g->device);
    delete g;
    return LG_OK;
}

static int perm_args(PermArgs& a, const lg_ring* r, int nl, const lg_poly* in, lg_poly* out, const char* what) {
    LG_REQUIRE(in && out, "%s: null polynomial", what);
    LG_REQUIRE(in->N == out->N && in->batch == out->batch, "%s: shape mismatch", what);
    LG_REQUIRE(nl >= 1 && nl <= in->nlimbs && nl <= out->nlimbs, "%s: limb count out of range", what);
    LG_REQUIRE(in->d != out->d, "%s: not in place (ring_galois.go:54,88,105)", what);
    LG_SAME_DEVICE(what, in->device, out->device);
    if (r) {
        LG_REQUIRE(r->N == in->N && nl <= r->nl, "%s: ring mismatch", what);
        LG_SAME_DEVICE(what, r->device, in->device);
        a.T = r->T;
    } else {
        memset(&a.T, 0, sizeof(a.T));
        a.T.N = (u32)in->N;
        a.T.logN = lgh::log2u(in->N);
    }
    a.map = limb_map_identity();
    a.in = in->d;
    a.out = out->d;
    a.in_bs = in->bstride;
    a.out_bs = out->bstride;
    a.index = nullptr;
    a.gen = 0;
    return LG_OK;
}
int lg_ring_permute_ntt_with_index(int nl, const lg_poly* in, const lg_galois* g, lg_poly* out, lg_stream_t s) {
    PermArgs a;
    LG_REQUIRE(g, "PermuteNTTWithIndex: null index");
    LG_TRY(perm_args(a, nullptr, nl, in, out, "PermuteNTTWithIndex"));
    LG_REQUIRE(g->N == in->N, "PermuteNTTWithIndex: index length mismatch");
    LG_SAME_DEVICE("PermuteNTTWithIndex", g->device, in->device);
    LG_ON_DEVICE(in->device);
    a.index = g->d_index.d;
    lg_launch_permute_ntt(a, nl, out->batch, cs(s));
    LG_LAUNCH_CHECK();
    return LG_OK;
}
int lg_ring_permute_ntt(int nl, const lg_poly* in, uint64_t gen, lg_poly* out, lg_stream_t s) {
    // ring_galois.go:55-84: the index is rebuilt per call from gen (not exponentiated)
    LG_REQUIRE(in, "PermuteNTT: null polynomial");
    LG_ON_DEVICE(in->device);
    lg_galois* g = nullptr;
    LG_TRY(make_galois(permute_ntt_index(gen, in->N), in->N, &g));
    int rc = lg_ring_permute_ntt_with_index(nl, in, g, out, s);
    if (rc == LG_OK) cudaStreamSynchronize(cs(s));  // the temporary index must outlive the kernel
    lg_galois_destroy(g);
    return rc;
}
int lg_ring_permute(const lg_ring* r, int nl, const lg_poly* in, uint64_t gen, lg_poly* out, lg_stream_t s) {
    PermArgs a;
    LG_REQUIRE(r, "Permute: null ring");
    LG_TRY(perm_args(a, r, nl, in, out, "Permute"));
    LG_ON_DEVICE(r->device);
    a.gen = gen;
    lg_launch_permute_coeff(a, nl, out->batch, cs(s));
    LG_LAUNCH_CHECK();
    return LG_OK;
}
int lg_ring_mult_by_monomial(const lg_ring* r, int nl, const lg_poly* p1, uint64_t deg, lg_poly* p2, lg_stream_t s) {
    PermArgs a;
    LG_REQUIRE(r, "MultByMonomial: null ring");
    LG_TRY(perm_args(a, r, nl, p1, p2, "MultByMonomial"));
    LG_ON_DEVICE(r->device);
    a.gen = deg;
    lg_launch_mult_by_monomial(a, nl, p2->batch, cs(s));
    LG_LAUNCH_CHECK();
    return LG_OK;
}
int lg_ring_bitreverse(const lg_ring* r, int nl, const lg_poly* p1, lg_poly* p2, lg_stream_t s) {
    PermArgs a;
    LG_REQUIRE(r, "BitReverse: null ring");
    LG_TRY(perm_args(a, r, nl, p1, p2, "BitReverse"));
    LG_ON_DEVICE(r->device);
    lg_launch_bitreverse(a, nl, p2->batch, cs(s));
    LG_LAUNCH_CHECK();
    return LG_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------
// RNS rescaling, ring/ring_scaling.go
// ---------------------------------------------------------------------------

// One division by the last of `nl` active moduli on a [batch][..][N] buffer.
int lgi_div_by_last_modulus(const lg_ring* r, int nl, int batch, u64* p0, size_t bs, bool round, bool ntt,
                            cudaStream_t st) {
    LG_REQUIRE(nl >= 2 && nl <= r->nl, "DivByLastModulus: needs between 2 and %d active limbs, got %d", r->nl, nl);
    LG_REQUIRE(!r->rescale.empty(), "DivByLastModulus: ring was created without rescaleParams");
    const int level = nl - 1;
    const u64 N = r->N;
    u64* last = p0 + (size_t)level * N;
    if (ntt) {  // :17 / :80  InvNTT of the last limb, in place
        LimbMap m{1 << 30, level, 0};
        LG_TRY(lgi_ntt(r, m, 1, batch, last, bs, last, bs, true, 0, 0, st));
    }
    Scratch tmp(st);
    LG_TRY(tmp.alloc((size_t)batch * level * N));
    const size_t tbs = (size_t)level * N;
    if (ntt && lgi_ntt_tail_ok(r)) {
        // :17-30 / :80-109 in one transform: the first phase reads the last limb for every target limb (+ pHalf terms),
        // the last phase applies (x - y) * q_last^-1 straight from its registers
        NttBcast bc;
        memset(&bc, 0, sizeof(bc));
        bc.enabled = 1;
        const size_t row = (size_t)level * (level - 1) / 2;
        if (round) {
            // :82-88 once, in place on the last limb as the reference does: last = CRed(last + pHalf, q_last)
            const u64 phalf = (r->q[level] - 1) >> 1;
            LimbMap m{1 << 30, level, 0};
            LG_TRY(lgi_ew(EW_ADD_SCALAR, r, m, 1, batch, last, bs, nullptr, 0, last, bs, &phalf, 1, st));
            bc.add = r->d_phalfneg.d + row;  // :97 pHalfNegQi
        }
        NttTail t;
        memset(&t, 0, sizeof(t));
        t.enabled = 1;
        t.split = batch;
        t.a[0] = p0;
        t.out[0] = p0;
        t.a_bs[0] = t.out_bs[0] = bs;
        t.s = r->d_rescale.d + row;  // rescaleParams[level-1][.]
        return lgi_ntt(r, limb_map_identity(), level, batch, last, bs, tmp.d, tbs, false, 0, 0, st, false, &t, &bc);
    }
    FanoutArgs f;
    f.N = (u32)N;
    f.in = last;
    f.in_bs = bs;
    f.nruns = 1;
    f.out[0] = tmp.d;
    f.out_bs[0] = tbs;
    f.ndst[0] = level;
    f.mode = round ? 1 : 0;
    f.phalf = 0;
    f.plast = r->q[level];
    if (round) {
        const u64 phalf = (r->q[level] - 1) >> 1;  // :82
        f.phalf = phalf;
        for (int i = 0; i < level; ++i) f.add[i] = r->q[i] - (phalf % r->q[i]);  // :97 pHalfNegQi
    }
    lg_launch_fanout(f, batch, st);
    LG_LAUNCH_CHECK();
    if (ntt)  // :21 / :105
        LG_TRY(lgi_ntt(r, limb_map_identity(), level, batch, tmp.d, tbs, tmp.d, tbs, false, 0, 0, st));
    else  // :48 / :143  BRedAdd of the broadcast limb
        LG_TRY(lgi_ew(EW_REDUCE, r, limb_map_identity(), level, batch, tmp.d, tbs, nullptr, 0, tmp.d, tbs, nullptr, 0, st));
    std::vector<u64> sc(level);
    for (int i = 0; i < level; ++i) sc[i] = r->rescale_param(level, i);
    // :30 / :50 / :109 / :144
    return lgi_ew(EW_SUB_MULMONT_SCALAR, r, limb_map_identity(), level, batch, p0, bs, tmp.d, tbs, p0, bs, sc.data(), level,
                  st);
}

static int ring_div(const lg_ring* r, int nl, lg_poly* p0, int nb, bool round, bool ntt, bool many_ntt, lg_stream_t s,
                    const char* what) {
    LG_REQUIRE(r, "%s: null ring", what);
    LG_TRY(check_poly(r, nl, p0, what));
    LG_REQUIRE(nb >= 1 && nb < nl, "%s: cannot drop %d of %d limbs", what, nb, nl);
    LG_ON_DEVICE(r->device);
    if (many_ntt)  // :58 / :153
        LG_TRY(lgi_ntt(r, limb_map_identity(), nl, p0->batch, p0->d, p0->bstride, p0->d, p0->bstride, true, 0, 0, cs(s)));
    for (int k = 0; k < nb; ++k) LG_TRY(lgi_div_by_last_modulus(r, nl - k, p0->batch, p0->d, p0->bstride, round, ntt, cs(s)));
    if (many_ntt)  // :60 / :155
        LG_TRY(lgi_ntt(r, limb_map_identity(), nl - nb, p0->batch, p0->d, p0->bstride, p0->d, p0->bstride, false, 0, 0,
                       cs(s)));
    return LG_OK;
}

extern "C" {
int lg_ring_div_floor_by_last_modulus_ntt(const lg_ring* r, int nl, lg_poly* p0, lg_stream_t s) {
    return ring_div(r, nl, p0, 1, false, true, false, s, "DivFloorByLastModulusNTT");
}
int lg_ring_div_floor_by_last_modulus(const lg_ring* r, int nl, lg_poly* p0, lg_stream_t s) {
    return ring_div(r, nl, p0, 1, false, false, false, s, "DivFloorByLastModulus");
}
int lg_ring_div_floor_by_last_modulus_many_ntt(const lg_ring* r, int nl, lg_poly* p0, int nb, lg_stream_t s) {
    return ring_div(r, nl, p0, nb, false, false, true, s, "DivFloorByLastModulusManyNTT");
}
int lg_ring_div_floor_by_last_modulus_many(const lg_ring* r, int nl, lg_poly* p0, int nb, lg_stream_t s) {
    return ring_div(r, nl, p0, nb, false, false, false, s, "DivFloorByLastModulusMany");
}
int lg_ring_div_round_by_last_modulus_ntt(const lg_ring* r, int nl, lg_poly* p0, lg_stream_t s) {
    return ring_div(r, nl, p0, 1, true, true, false, s, "DivRoundByLastModulusNTT");
}
int lg_ring_div_round_by_last_modulus(const lg_ring* r, int nl, lg_poly* p0, lg_stream_t s) {
    return ring_div(r, nl, p0, 1, true, false, false, s, "DivRoundByLastModulus");
}
int lg_ring_div_round_by_last_modulus_many_ntt(const lg_ring* r, int nl, lg_poly* p0, int nb, lg_stream_t s) {
    return ring_div(r, nl, p0, nb, true, false, true, s, "DivRoundByLastModulusManyNTT");
}
int lg_ring_div_round_by_last_modulus_many(const lg_ring* r, int nl, lg_poly* p0, int nb, lg_stream_t s) {
    return ring_div(r, nl, p0, nb, true, false, false, s, "DivRoundByLastModulusMany");
}
}  // extern "C"
