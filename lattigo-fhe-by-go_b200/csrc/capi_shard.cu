// capi_shard.cu -- C-ABI host layer, part 4: multi-GPU paths (one process per GPU).
//
// (1) lg_comm: the ranks of one node.  Two transports:
//       * peer memory over NVLink: every rank owns an exchange buffer that its peers map (CUDA IPC between
//         processes, plain pointers inside one process) and an arrival-flag array; the data path never calls a
//         library collective -- the consuming kernels LOAD the limbs they need straight from the owner's buffer,
//         and a one-CTA barrier kernel (release store into every peer's flag array, acquire spin on the own one)
//         orders producers and consumers in stream order;
//       * NCCL (resolved with dlopen so that the library the host process already uses is shared) for the one
//         true reduction of the path, AggregateShares.
// (2) Limb axis (BASELINE config 4): one CKKS ciphertext (or a small batch) with its RNS limbs
//     spread CYCLICALLY over the ranks -- rank r of w owns table limbs r, r+w, r+2w, ... of Q || P -- so ownership
//     is balanced at every level and does not move when a limb is dropped.  Ciphertexts stay limb-resident between
//     ops: a rank's polynomial buffers are full size but only its own limbs are meaningful.  NTT, tensor,
//     multiply-accumulate and the ModDown / rescale tails are limb-local; limbs cross NVLink exactly where a basis
//     extension needs every source limb:
//         c2 (coefficient domain)    -> read by DecomposeAndSplit of every rank   (ckks/evaluator.go:1503-1513)
//         special-prime accumulators -> read by the ModDown basis extension       (ring_basis_extension.go:219-226)
//         the last limb              -> read by every rank's rescale              (ring_scaling.go:80-103)
//     The replicated-in / replicated-out entry points of round 1 remain (resident op + gather of the result limbs).
// (3) Party axis (config 5): AggregateShares of the dckks/dbfv protocols is an all-reduce(sum, u64) followed by one
//     Reduce, which equals the reference's pairwise CRed-add chain (dckks/publickey_gen.go:45-47) for up to 8
//     canonical shares of < 2^61.
#include <dlfcn.h>
#include <string.h>

#include <mutex>

#include "capi_internal.hpp"

static inline cudaStream_t cs(lg_stream_t s) { return (cudaStream_t)s; }

// ---- NCCL, resolved at run time (ABI-stable subset of nccl.h) -------------------------------------
namespace {
typedef struct {
    char internal[128];
} NcclUniqueId;
typedef void* NcclComm;
enum { kNcclUint64 = 5, kNcclSum = 0 };

struct NcclApi {
    void* handle = nullptr;
    int (*GetUniqueId)(NcclUniqueId*) = nullptr;
    int (*CommInitRank)(NcclComm*, int, NcclUniqueId, int) = nullptr;
    int (*CommDestroy)(NcclComm) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
};

NcclApi* nccl() {
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) {
            api.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
            if (api.handle) break;
        }
        if (api.handle) {
#define LG_SYM(field, name) *(void**)(&api.field) = dlsym(api.handle, name)
            LG_SYM(GetUniqueId, "ncclGetUniqueId");
            LG_SYM(CommInitRank, "ncclCommInitRank");
            LG_SYM(CommDestroy, "ncclCommDestroy");
            LG_SYM(GroupStart, "ncclGroupStart");
            LG_SYM(GroupEnd, "ncclGroupEnd");
            LG_SYM(AllReduce, "ncclAllReduce");
            LG_SYM(GetErrorString, "ncclGetErrorString");
#undef LG_SYM
        }
    });
    return &api;
}

#define LG_NCCL_CHECK(expr)                                                                      \
    do {                                                                                         \
        int _r = (expr);                                                                         \
        if (_r != 0) {                                                                           \
            lg_set_error("%s:%d: %s: %s", __FILE__, __LINE__, #expr,                             \
                         nccl()->GetErrorString ? nccl()->GetErrorString(_r) : "nccl error");    \
            return LG_ERR_CUDA;                                                                  \
        }                                                                                        \
    } while (0)

constexpr int kMaxRanks = 16;
}  // namespace

struct lg_comm {
    int device = -1;  // the rank's GPU
    int world = 1, rank = 0;
    NcclComm comm = nullptr;
    // peer-memory exchange (limb axis)
    u64* xbuf[kMaxRanks] = {};    // exchange buffer of every rank as mapped here; xbuf[rank] is the own one
    u32* flags[kMaxRanks] = {};   // flags[r]: rank r's arrival array [world] (the tail of its exchange buffer: one IPC
                                  // mapping covers both), slot s is written by rank s
    bool ipc[kMaxRanks] = {};     // mapped with cudaIpcOpenMemHandle (closed on destroy)
    size_t xwords = 0;            // capacity of one exchange buffer in words; each half serves every other op
    u32** d_flag_tab = nullptr;   // device copy of flags[]
    u32* d_err = nullptr;         // set by a barrier that timed out
    uint32_t epoch = 0;           // barriers issued so far
    uint64_t opseq = 0;           // sharded ops issued so far: op n exchanges through half n & 1
    size_t bump = 0;              // words of the current half handed out to the current op
    bool peers_ready() const {
        for (int r = 0; r < world; ++r)
            if (!xbuf[r] || !flags[r]) return false;
        return d_flag_tab != nullptr;
    }
};

namespace {

// ---- cross-GPU barrier in stream order ----------------------------------------------------------------------
// Thread t publishes this rank's arrival (epoch) into rank t's flag array and waits until rank t's arrival shows in
// the own array.  Everything this rank enqueued before the barrier has completed when the kernel starts (stream
// order), so its exchange-buffer contents are in its L2 / HBM -- the point of coherence peers read through; consumers
// are launched after the barrier kernel, and a kernel launch invalidates L1, so they cannot see stale peer lines.
// The spin is bounded (5 s of %globaltimer): a lost peer raises the error flag instead of hanging the GPU.
__global__ void xbarrier_kernel(u32* const* flag_tab, int world, int rank, u32 epoch, u32* err) {
    const int t = threadIdx.x;
    if (t >= world || t == rank) return;
    __threadfence_system();
    volatile u32* dst = flag_tab[t] + rank;
    *dst = epoch;
    __threadfence_system();
    volatile u32* src = flag_tab[rank] + t;
    unsigned long long t0, t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    while ((int)(*src - epoch) < 0) {
        __nanosleep(200);
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > 5000000000ull) {
            atomicExch(err, 1u);
            break;
        }
    }
    __threadfence_system();
}

int xbarrier(lg_comm* c, cudaStream_t st) {
    if (c->world == 1) return LG_OK;
    c->epoch += 1;
    xbarrier_kernel<<<1, 32, 0, st>>>(c->d_flag_tab, c->world, c->rank, c->epoch, c->d_err);
    lg_g_launches += 1;
    LG_LAUNCH_CHECK();
    return LG_OK;
}

// ---- exchange-buffer bookkeeping: op n uses half n & 1, regions are bump-allocated inside the half -------------
// Reuse of a half is safe: before rank A writes half h again (op n+2) it has passed a barrier of op n+1, which every
// peer reaches only after all its reads of op n have completed (stream order).  Every op issues at least one barrier.
int op_begin(lg_comm* c) {
    LG_REQUIRE(c->world == 1 || c->peers_ready(),
               "limb-sharded op: exchange buffers are not set up (lg_comm_xbuf_alloc, then lg_comm_xbuf_open / _attach for every peer)");
    c->opseq += 1;
    c->bump = 0;
    return LG_OK;
}
int xalloc(lg_comm* c, size_t words, size_t* off) {
    if (c->world == 1 && !c->xbuf[0]) {  // single rank without an exchange buffer: allocate one lazily
        lg_set_error("limb-sharded op on a single rank needs lg_comm_xbuf_alloc as well");
        return LG_ERR_ARG;
    }
    const size_t half = c->xwords / 2;
    words = (words + 3) & ~(size_t)3;  // keep 32-byte alignment
    LG_REQUIRE(c->bump + words <= half,
               "limb-sharded op needs %zu exchange words per half, %zu reserved: call lg_comm_xbuf_alloc with a larger size",
               c->bump + words, half);
    *off = (size_t)(c->opseq & 1) * half + c->bump;
    c->bump += words;
    return LG_OK;
}

// cyclic ownership over the table limbs of Q || P
struct Own {
    int w, r;
    int nQ;    // limbs of the Q ring (table offset of the special primes)
    int n0q;   // own Q limbs among [0, nl)
    int k0;    // first own special prime (index within P)
    int n0p;   // own special primes
    int n() const { return n0q + n0p; }
    LimbMap qp_map() const { return LimbMap{n0q, r, nQ + k0, w}; }
    LimbMap q_map() const { return LimbMap{1 << 30, r, 0, w}; }
    LimbMap p_map() const { return LimbMap{1 << 30, k0, 0, w}; }
};
int count_own(int n, int w, int first) { return n > first ? (n - first + w - 1) / w : 0; }
Own make_own(const lg_comm* c, int nQ, int nP, int nl) {
    Own o;
    o.w = c->world;
    o.r = c->rank;
    o.nQ = nQ;
    o.n0q = count_own(nl, o.w, o.r);
    o.k0 = ((o.r - nQ) % o.w + o.w) % o.w;
    o.n0p = count_own(nP, o.w, o.k0);
    return o;
}

int launch_ntt(const NttArgs& a, int nlimbs, int batch, bool inverse, cudaStream_t st) {
    if (nlimbs <= 0 || batch <= 0) return LG_OK;
    LG_REQUIRE(lg_launch_ntt(a, nlimbs, batch, inverse, st) == 0, "sharded NTT: unsupported ring degree");
    LG_LAUNCH_CHECK();
    return LG_OK;
}

// strided copy of limb sets: dst limb k <- src limb k, limb strides in words
int copy_limbs(const lg_ring* R, int nlimbs, int batch, const u64* src, size_t src_bs, size_t src_ls, u64* dst, size_t dst_bs,
               size_t dst_ls, cudaStream_t st) {
    if (nlimbs <= 0) return LG_OK;
    EwArgs g;
    g.T = R->T;
    g.map = LimbMap{1 << 30, 0, 0, 0};  // the copy uses no table: every data limb reads table limb 0
    g.a = src;
    g.b = nullptr;
    g.c = dst;
    g.a_bs = src_bs;
    g.b_bs = 0;
    g.c_bs = dst_bs;
    g.a_ls = src_ls;
    g.b_ls = 0;
    g.c_ls = dst_ls;
    LG_REQUIRE(lg_launch_ew(EW_COPY, g, nlimbs, batch, st) == 0, "copy: launch failed");
    LG_LAUNCH_CHECK();
    return LG_OK;
}

// Decompose(AndSplit) (ring_basis_extension.go:476-713) of digit `crt` for the target limbs this rank owns, the
// source limbs read from their owners' exchange buffers (region c2_off, layout [batch][nl][N] by global limb).
int decompose_own(const lg_comm* c, const lg_decomposer* d, const Own& o, int level, int crt, int batch, size_t c2_off, u64* Di,
                  size_t d_bs, cudaStream_t st) {
    const int nl = level + 1;
    const int alphai = d->xalpha[crt];
    const int p0idxst = crt * d->alpha;
    const int p0idxed = p0idxst + alphai;
    const u64 N = d->N;
    const size_t c2_bs = (size_t)nl * N;
    auto src_limb = [&](int j) { return c->xbuf[j % c->world] + c2_off + (size_t)j * N; };
    if ((p0idxed > level + 1 && (level + 1) % d->nP == 1) || alphai == 1) {  // :489 / :613
        FanoutArgs f;
        f.N = (u32)N;
        f.in = src_limb(p0idxst);
        f.in_bs = c2_bs;
        f.nruns = 1;
        f.out[0] = Di;  // own Q targets then own special primes: consecutive compact slots
        f.out_bs[0] = d_bs;
        f.ndst[0] = o.n();
        f.mode = 0;
        f.phalf = f.plast = 0;
        lg_launch_fanout(f, batch, st);
        LG_LAUNCH_CHECK();
        return LG_OK;
    }
    const int index = (level >= alphai + crt * d->alpha) ? d->xalpha[crt] - 2 : (level - 1) % d->alpha;  // :503-507 / :631-635
    LG_REQUIRE(index >= 0 && index < (int)d->modup[crt].size(), "Decompose: no parameters for digit %d index %d", crt, index);
    const ModUpDev& m = *d->modup[crt][index];
    ModUpArgs a;
    memset(&a, 0, sizeof(a));
    a.M = m.M;
    a.N = (u32)N;
    a.nsrc = index + 2;
    LG_REQUIRE(a.nsrc <= 4, "sharded Decompose: at most 4 limbs per digit");
    for (int s = 0; s < a.nsrc; ++s) a.src[s] = src_limb(p0idxst + s);
    a.in = a.src[0];
    a.in_bs = c2_bs;
    a.nruns = 2;
    a.out[0] = Di;
    a.out_bs[0] = d_bs;
    a.ndst[0] = o.n0q;
    a.tgt0[0] = o.r;
    a.out[1] = Di + (size_t)o.n0q * N;
    a.out_bs[1] = d_bs;
    a.ndst[1] = o.n0p;
    a.tgt0[1] = d->nQ + o.k0;
    a.tstep = o.w;
    a.fast = m.fast_level(a.nsrc, &a.fp_shift);
    LG_REQUIRE(a.fast >= 1, "sharded Decompose: moduli of 61 bits and more are not supported");
    a.lazy_out = 1;  // read by the forward NTT alone
    LG_REQUIRE(lg_launch_modup(a, batch, st) == 0, "Decompose: too many source limbs");
    LG_LAUNCH_CHECK();
    return LG_OK;
}

// Limb-resident switchKeysInPlace (ckks/evaluator.go:1475-1558).  cx: this rank's own Q limbs of the NTT-domain input,
// own limb k at cx + b*cx_bs + k*cx_ls.  out0/out1: user-layout polynomials (limb j at + j*N); the own Q limbs are
// written, or accumulated into with add0/add1.
int switch_keys_resident(lg_ckks_eval* e, lg_comm* c, int level, int batch, const u64* cx, size_t cx_bs, size_t cx_ls,
                         const lg_swk* evk, u64* out0, size_t out0_bs, u64* out1, size_t out1_bs, bool add0, bool add1,
                         bool cx_in_range, cudaStream_t st) {
    const lg_ring* Q = e->Q;
    const lg_ring* P = e->P;
    const lg_ring* QP = e->QP.get();
    const u64 N = Q->N;
    const int nQ = Q->nl, nP = P->nl, nl = level + 1;
    LG_REQUIRE(Q->logN >= 12, "sharded key switch needs N >= 2^12");
    LG_REQUIRE(level >= 0 && level < nQ, "switchKeys: level %d out of range", level);
    LG_REQUIRE(evk && evk->N == N && evk->nQP == nQ + nP, "switchKeys: switching key shape mismatch");
    LG_SAME_DEVICE("switchKeys", Q->device, evk->device);
    LG_REQUIRE(nP <= 4, "sharded key switch: at most 4 special primes");
    const int alpha = e->alpha, beta = (nl + alpha - 1) / alpha;
    LG_REQUIRE(beta <= evk->beta, "switchKeys: key has %d digits, %d needed", evk->beta, beta);
    const Own o = make_own(c, nQ, nP, nl);
    const int w = o.w, r = o.r;
    const size_t wN = (size_t)w * N;

    size_t c2_off, pacc_off;
    LG_TRY(xalloc(c, (size_t)batch * nl * N, &c2_off));
    LG_TRY(xalloc(c, (size_t)2 * batch * nP * N, &pacc_off));
    u64* xown = c->xbuf[c->rank];

    // :1503 c2 = InvNTT(cx) on the own Q limbs, written where the peers read it (global limb position)
    {
        NttArgs a;
        memset(&a, 0, sizeof(a));
        a.T = Q->T;
        a.map = o.q_map();
        a.in = cx;
        a.in_bstride = cx_bs;
        a.in_ls = cx_ls;
        a.out = xown + c2_off + (size_t)r * N;
        a.out_bstride = (size_t)nl * N;
        a.out_ls = wN;
        Scratch flags(st);
        if (!cx_in_range && o.n0q > 0) {
            LG_TRY(flags.alloc(((size_t)batch * o.n0q + 1) / 2));
            lg_launch_range_flags(a, o.n0q, batch, (u32*)flags.d, st);
            a.flags = (const u32*)flags.d;
        }
        LG_TRY(launch_ntt(a, o.n0q, batch, true, st));
    }
    LG_TRY(xbarrier(c, st));

    // :1511-1552 digit loop on the own target limbs (compact scratch: own Q limbs, then own special primes)
    const int no = o.n();
    const size_t d_bs = (size_t)no * N;
    Scratch acc(st), D(st), tmp(st);
    LG_TRY(acc.alloc((size_t)2 * batch * d_bs + 4));
    u64* acc0 = acc.d;
    u64* acc1 = acc.d + (size_t)batch * d_bs;
    if (no > 0) {
        LG_TRY(D.alloc((size_t)beta * batch * d_bs));
        const size_t d_ds = (size_t)batch * d_bs;
        // independent per-digit basis extensions: for a few ciphertexts they overlap on auxiliary streams (capi_ext.cu)
        LgAux* aux = ((size_t)batch * (N / 2 / 128) < 2 * 148 && beta > 1 && !lg_switches().no_aux_streams.load(std::memory_order_relaxed))
                         ? lg_aux_streams()
                         : nullptr;
        if (aux) lg_aux_fork(aux, st, LG_AUX_STREAMS);
        for (int i = 0; i < beta; ++i) {
            const int rc = decompose_own(c, e->dec.get(), o, level, i, batch, c2_off, D.d + (size_t)i * d_ds, d_bs,
                                         aux ? aux->s[i % LG_AUX_STREAMS] : st);
            if (rc != LG_OK) {
                if (aux) lg_aux_join(aux, st, LG_AUX_STREAMS);
                return rc;
            }
        }
        if (aux) lg_aux_join(aux, st, LG_AUX_STREAMS);
        NttArgs a;
        memset(&a, 0, sizeof(a));
        a.T = QP->T;
        a.map = o.qp_map();
        a.in = D.d;
        a.out = D.d;
        a.in_bstride = a.out_bstride = d_bs;
        a.skip_alpha = alpha;  // the digit's own limbs come from the NTT-domain input
        a.skip_div = batch;
        a.skip_nl = nl;
        LG_REQUIRE(lg_launch_ntt_fwd_strided(a, no, beta * batch, st) == 0, "switchKeys: strided NTT launch failed");
        LG_LAUNCH_CHECK();
        KsFusedArgs k;
        memset(&k, 0, sizeof(k));
        k.T = QP->T;
        k.map = o.qp_map();
        if (!lg_switches().no_fp_mac.load(std::memory_order_relaxed)) {
            LG_TRY(lgi_swk_prepare(evk, QP, st));
            k.evk_f = evk->d_f;
            k.key_bad = evk->d_bad;
            if (evk->has_map) {
                k.h_keymap = evk->keymap;
                k.h_fp_ok = evk->fp_ok.data();
            }
        }
        k.D = D.d;
        k.d_ds = d_ds;
        k.d_bs = d_bs;
        k.cx = cx;
        k.cx_bs = cx_bs;
        k.cx_ls = cx_ls;
        k.evk = evk->key(0, 0);
        k.evk_ds = (size_t)(evk->key(1, 0) - evk->key(0, 0));
        k.evk_hs = (size_t)(evk->key(0, 1) - evk->key(0, 0));
        k.acc0 = acc0;
        k.acc1 = acc1;
        k.acc_bs = d_bs;
        k.beta = beta;
        k.alpha = alpha;
        k.nl = nl;
        LG_REQUIRE(lg_launch_ks_fused(k, no, batch, st) == 0, "switchKeys: fused digit loop launch failed");
        LG_LAUNCH_CHECK();
    }

    // :1556-1557 ModDownSplitedNTTPQ of both accumulators: InvNTT of the own special-prime limbs into the exchange
    // buffer (the accumulators are canonical: no range check), every rank then extends P -> its own Q limbs
    {
        NttArgs a;
        memset(&a, 0, sizeof(a));
        a.T = P->T;
        a.map = o.p_map();
        a.in = acc0 + (size_t)o.n0q * N;
        a.in_bstride = d_bs;
        a.out = xown + pacc_off + (size_t)o.k0 * N;
        a.out_bstride = (size_t)nP * N;
        a.out_ls = wN;
        LG_TRY(launch_ntt(a, o.n0p, 2 * batch, true, st));
    }
    LG_TRY(xbarrier(c, st));
    if (o.n0q > 0) {
        const size_t t_bs = (size_t)o.n0q * N;
        LG_TRY(tmp.alloc((size_t)2 * batch * t_bs));
        ModUpArgs a;
        memset(&a, 0, sizeof(a));
        a.M = e->ext->pq.M;
        a.N = (u32)N;
        a.nsrc = nP;
        for (int k = 0; k < nP; ++k) a.src[k] = c->xbuf[(nQ + k) % w] + pacc_off + (size_t)k * N;
        a.in = a.src[0];
        a.in_bs = (size_t)nP * N;
        a.nruns = 1;
        a.out[0] = tmp.d;
        a.out_bs[0] = t_bs;
        a.ndst[0] = o.n0q;
        a.tgt0[0] = r;
        a.tstep = w;
        a.fast = e->ext->pq.fast_level(a.nsrc, &a.fp_shift);
        LG_REQUIRE(a.fast >= 1, "sharded ModDown: moduli of 61 bits and more are not supported");
        a.lazy_out = 1;
        LG_REQUIRE(lg_launch_modup(a, 2 * batch, st) == 0, "modUpExact: too many source limbs");
        LG_LAUNCH_CHECK();
        // forward NTT whose last phase applies (acc - NTT(t)) * P^-1 (+ add) from its registers
        NttArgs n;
        memset(&n, 0, sizeof(n));
        n.T = Q->T;
        n.map = o.q_map();
        n.in = tmp.d;
        n.out = tmp.d;
        n.in_bstride = n.out_bstride = t_bs;
        n.tail.enabled = 1;
        n.tail.split = batch;
        n.tail.add[0] = add0 ? 1 : 0;
        n.tail.add[1] = add1 ? 1 : 0;
        n.tail.a_canon = lg_switches().no_tail_canon.load(std::memory_order_relaxed) ? 0 : 1;
        n.tail.a[0] = acc0;
        n.tail.a[1] = acc1;
        n.tail.a_bs[0] = n.tail.a_bs[1] = d_bs;
        n.tail.a_ls = N;
        n.tail.out[0] = out0 + (size_t)r * N;
        n.tail.out[1] = out1 + (size_t)r * N;
        n.tail.out_bs[0] = out0_bs;
        n.tail.out_bs[1] = out1_bs;
        n.tail.out_ls = wN;
        n.tail.s = e->ext->d_moddown_pq.d;
        LG_TRY(launch_ntt(n, o.n0q, 2 * batch, false, st));
    }
    return LG_OK;
}

// Rescale loop body (ckks/evaluator.go:955-960, DivRoundByLastModulusNTT ring_scaling.go:72-114) on limb-resident
// polynomials: the owner of the last limb inverse-transforms it into its exchange buffer, every rank copies it and runs
// its own target limbs (first NTT phase reads the copy for every target, last phase applies (x - y) * q_last^-1).
int rescale_resident(lg_ckks_eval* e, lg_comm* c, int nl, int batch, u64* p0, size_t p0_bs, u64* p1, size_t p1_bs, cudaStream_t st) {
    const lg_ring* Q = e->Q;
    const u64 N = Q->N;
    LG_REQUIRE(nl >= 2 && nl <= Q->nl, "cannot Rescale: input Ciphertext already at level 0");
    LG_REQUIRE(Q->logN >= 12, "sharded rescale needs N >= 2^12");
    LG_REQUIRE(!Q->rescale.empty(), "Rescale: ring was created without rescaleParams");
    const int level = nl - 1, w = c->world, r = c->rank, owner = level % w;
    const size_t wN = (size_t)w * N;
    size_t last_off;
    LG_TRY(xalloc(c, (size_t)2 * batch * N, &last_off));
    u64* polys[2] = {p0, p1};
    const size_t pbs[2] = {p0_bs, p1_bs};
    if (r == owner) {
        const LimbMap lm{1 << 30, level, 0, 1};
        const u64 phalf = (Q->q[level] - 1) >> 1;
        for (int i = 0; i < 2; ++i) {
            u64* dst = c->xbuf[r] + last_off + (size_t)i * batch * N;
            LG_TRY(lgi_ntt(Q, lm, 1, batch, polys[i] + (size_t)level * N, pbs[i], dst, N, true, 0, 0, st));  // :80
            LG_TRY(lgi_ew(EW_ADD_SCALAR, Q, lm, 1, batch, dst, N, nullptr, 0, dst, N, &phalf, 1, st));         // :82-88
        }
    }
    LG_TRY(xbarrier(c, st));
    const int n = count_own(level, w, r);
    if (n > 0) {
        Scratch last(st), tmp(st);
        LG_TRY(last.alloc((size_t)2 * batch * N));
        LG_TRY(copy_limbs(Q, 1, 2 * batch, c->xbuf[owner] + last_off, N, N, last.d, N, N, st));  // one NVLink read per word
        LG_TRY(tmp.alloc((size_t)2 * batch * n * N));
        const size_t row = (size_t)level * (level - 1) / 2;
        NttArgs a;
        memset(&a, 0, sizeof(a));
        a.T = Q->T;
        a.map = LimbMap{1 << 30, r, 0, w};
        a.in = last.d;
        a.in_bstride = N;
        a.out = tmp.d;
        a.out_bstride = (size_t)n * N;
        a.bcast.enabled = 1;
        a.bcast.add = Q->d_phalfneg.d + row;  // :97 pHalfNegQi, by table limb
        a.tail.enabled = 1;
        a.tail.split = batch;
        for (int i = 0; i < 2; ++i) {
            a.tail.a[i] = polys[i] + (size_t)r * N;
            a.tail.out[i] = polys[i] + (size_t)r * N;
            a.tail.a_bs[i] = a.tail.out_bs[i] = pbs[i];
        }
        a.tail.a_ls = a.tail.out_ls = wN;
        a.tail.s = Q->d_rescale.d + row;  // rescaleParams[level-1][.], by table limb
        LG_TRY(launch_ntt(a, n, 2 * batch, false, st));
    }
    return LG_OK;
}

// replicate the first nl limbs of a limb-resident polynomial on every rank
int gather_limbs(const lg_ring* R, lg_comm* c, int nl, int batch, u64* p, size_t p_bs, size_t region_off, cudaStream_t st) {
    const u64 N = R->N;
    const int w = c->world, r = c->rank;
    const size_t wN = (size_t)w * N, g_bs = (size_t)nl * N;
    LG_TRY(copy_limbs(R, count_own(nl, w, r), batch, p + (size_t)r * N, p_bs, wN, c->xbuf[r] + region_off + (size_t)r * N, g_bs, wN, st));
    LG_TRY(xbarrier(c, st));
    for (int q = 0; q < w; ++q) {
        if (q == r) continue;
        LG_TRY(copy_limbs(R, count_own(nl, w, q), batch, c->xbuf[q] + region_off + (size_t)q * N, g_bs, wN, p + (size_t)q * N, p_bs, wN,
                          st));
    }
    return LG_OK;
}

int check_p(const lg_poly* p, u64 N, int nl, int batch, const char* what) {
    LG_REQUIRE(p, "%s: null polynomial", what);
    LG_REQUIRE(p->N == N, "%s: degree mismatch", what);
    LG_REQUIRE(p->nlimbs >= nl, "%s: polynomial has %d limbs, %d needed", what, p->nlimbs, nl);
    LG_REQUIRE(batch < 0 || p->batch == batch, "%s: batch mismatch", what);
    LG_SAME_DEVICE(what, lgi_expected_device(), p->device);
    return LG_OK;
}

// tensor of MulRelin (:1076-1095) on the own limbs: c0 -> out0, c1 -> out1 (user layout), c2 -> compact scratch
int tensor_own(lg_ckks_eval* e, const lg_comm* c, int nl, const lg_poly* a0, const lg_poly* a1, const lg_poly* b0, const lg_poly* b1,
               lg_poly* out0, lg_poly* out1, u64* c2, size_t c2_bs, cudaStream_t st) {
    const int n = count_own(nl, c->world, c->rank);
    if (n <= 0) return LG_OK;
    TensorArgs t;
    memset(&t, 0, sizeof(t));
    t.T = e->Q->T;
    t.a0 = a0->d;
    t.a1 = a1->d;
    t.b0 = b0->d;
    t.b1 = b1->d;
    t.c0 = out0->d;
    t.c1 = out1->d;
    t.c2 = c2;
    t.a_bs[0] = a0->bstride;
    t.a_bs[1] = a1->bstride;
    t.b_bs[0] = b0->bstride;
    t.b_bs[1] = b1->bstride;
    t.c_bs[0] = out0->bstride;
    t.c_bs[1] = out1->bstride;
    t.c_bs[2] = c2_bs;
    t.square = (a0->d == b0->d && a1->d == b1->d) ? 1 : 0;
    t.limb0 = c->rank;
    t.lstep = c->world;
    t.c2_compact = 1;
    lg_launch_tensor(t, n, a0->batch, st);
    LG_LAUNCH_CHECK();
    return LG_OK;
}

int gather_pair(lg_ckks_eval* e, lg_comm* c, int nl, lg_poly* p0, lg_poly* p1, cudaStream_t st) {
    if (c->world == 1) return LG_OK;
    size_t off;
    const size_t words = (size_t)p0->batch * nl * e->Q->N;
    LG_TRY(xalloc(c, words, &off));
    LG_TRY(gather_limbs(e->Q, c, nl, p0->batch, p0->d, p0->bstride, off, st));
    LG_TRY(xalloc(c, words, &off));
    return gather_limbs(e->Q, c, nl, p1->batch, p1->d, p1->bstride, off, st);
}

}  // namespace

extern "C" {

int lg_comm_get_unique_id(uint8_t* id128) {
    LG_REQUIRE(id128, "null argument");
    NcclApi* n = nccl();
    LG_REQUIRE(n->handle && n->GetUniqueId, "NCCL library not found (libnccl.so.2)");
    NcclUniqueId id;
    LG_NCCL_CHECK(n->GetUniqueId(&id));
    memcpy(id128, id.internal, 128);
    return LG_OK;
}
int lg_comm_create(int world, int rank, const uint8_t* id128, lg_comm** out) {
    LG_REQUIRE(out && world >= 1 && world <= kMaxRanks && rank >= 0 && rank < world, "lg_comm_create: invalid argument");
    std::unique_ptr<lg_comm> c(new lg_comm);
    c->device = lgi_current_device();
    c->world = world;
    c->rank = rank;
    if (world > 1 && id128) {  // without an id the handle serves the peer-memory paths only
        NcclApi* n = nccl();
        LG_REQUIRE(n->handle && n->CommInitRank, "NCCL library not found (libnccl.so.2)");
        NcclUniqueId id;
        memcpy(id.internal, id128, 128);
        LG_NCCL_CHECK(n->CommInitRank(&c->comm, world, id, rank));
    }
    *out = c.release();
    return LG_OK;
}
int lg_comm_destroy(lg_comm* c) {
    if (!c) return LG_OK;
    LG_ON_DEVICE(c->device);
    if (c->comm && nccl()->CommDestroy) nccl()->CommDestroy(c->comm);
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank || !c->ipc[r]) continue;
        if (c->xbuf[r]) cudaIpcCloseMemHandle(c->xbuf[r]);
    }
    if (c->xbuf[c->rank]) cudaFree(c->xbuf[c->rank]);
    if (c->d_flag_tab) cudaFree(c->d_flag_tab);
    if (c->d_err) cudaFree(c->d_err);
    delete c;
    return LG_OK;
}
int lg_comm_world(const lg_comm* c) { return c ? c->world : 0; }
int lg_comm_rank(const lg_comm* c) { return c ? c->rank : -1; }

// the ownership rule of the limb axis: limb t of Q || P belongs to rank t mod world   (host only)
int lg_comm_limb_owner(int limb, int world) { return (world >= 1 && limb >= 0) ? limb % world : -1; }

// ---- exchange buffers ---------------------------------------------------------------------------------------
static int comm_publish_tab(lg_comm* c) {
    for (int r = 0; r < c->world; ++r)
        if (!c->xbuf[r] || !c->flags[r]) return LG_OK;  // not complete yet
    LG_CUDA_CHECK(cudaMemcpy(c->d_flag_tab, c->flags, kMaxRanks * sizeof(u32*), cudaMemcpyHostToDevice));
    return LG_OK;
}
int lg_comm_xbuf_alloc(lg_comm* c, size_t words, uint8_t* handle128) {
    LG_REQUIRE(c && words >= 8, "lg_comm_xbuf_alloc: invalid argument");
    LG_REQUIRE(!c->xbuf[c->rank], "lg_comm_xbuf_alloc: already allocated");
    LG_ON_DEVICE(c->device);
    words = (words + 7) & ~(size_t)7;
    LG_CUDA_CHECK(cudaMalloc((void**)&c->xbuf[c->rank], (words + kMaxRanks) * sizeof(u64)));
    c->flags[c->rank] = reinterpret_cast<u32*>(c->xbuf[c->rank] + words);
    LG_CUDA_CHECK(cudaMemset(c->flags[c->rank], 0, kMaxRanks * sizeof(u64)));
    LG_CUDA_CHECK(cudaMalloc((void**)&c->d_err, sizeof(u32)));
    LG_CUDA_CHECK(cudaMemset(c->d_err, 0, sizeof(u32)));
    LG_CUDA_CHECK(cudaMalloc((void**)&c->d_flag_tab, kMaxRanks * sizeof(u32*)));
    c->xwords = words;
    if (handle128) {  // cudaIpcMemHandle_t (64 bytes) of the allocation, then its size in words
        cudaIpcMemHandle_t h0;
        static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
        LG_CUDA_CHECK(cudaIpcGetMemHandle(&h0, c->xbuf[c->rank]));
        memset(handle128, 0, 128);
        memcpy(handle128, &h0, 64);
        const uint64_t w64 = words;
        memcpy(handle128 + 64, &w64, 8);
    }
    return comm_publish_tab(c);
}
int lg_comm_xbuf_open(lg_comm* c, int peer, const uint8_t* handle128) {
    LG_REQUIRE(c && handle128 && peer >= 0 && peer < c->world && peer != c->rank, "lg_comm_xbuf_open: invalid argument");
    LG_REQUIRE(c->xbuf[c->rank], "lg_comm_xbuf_open: call lg_comm_xbuf_alloc first");
    LG_ON_DEVICE(c->device);
    cudaIpcMemHandle_t h0;
    uint64_t w64 = 0;
    memcpy(&h0, handle128, 64);
    memcpy(&w64, handle128 + 64, 8);
    LG_REQUIRE(w64 == c->xwords, "lg_comm_xbuf_open: rank %d reserved %llu exchange words, this rank %zu", peer,
               (unsigned long long)w64, c->xwords);
    LG_CUDA_CHECK(cudaIpcOpenMemHandle((void**)&c->xbuf[peer], h0, cudaIpcMemLazyEnablePeerAccess));
    c->flags[peer] = reinterpret_cast<u32*>(c->xbuf[peer] + c->xwords);
    c->ipc[peer] = true;
    return comm_publish_tab(c);
}
int lg_comm_xbuf_attach(lg_comm* c, int peer, const lg_comm* peer_comm) {
    LG_REQUIRE(c && peer_comm && peer >= 0 && peer < c->world && peer != c->rank && peer_comm->rank == peer,
               "lg_comm_xbuf_attach: invalid argument");
    LG_REQUIRE(c->xbuf[c->rank] && peer_comm->xbuf[peer], "lg_comm_xbuf_attach: both ranks must have called lg_comm_xbuf_alloc");
    LG_REQUIRE(c->xwords == peer_comm->xwords, "lg_comm_xbuf_attach: exchange buffers differ in size");
    LG_ON_DEVICE(c->device);
    if (peer_comm->device != c->device) {
        cudaError_t err = cudaDeviceEnablePeerAccess(peer_comm->device, 0);
        if (err != cudaSuccess && err != cudaErrorPeerAccessAlreadyEnabled) LG_CUDA_CHECK(err);
        cudaGetLastError();
    } else {
        // ranks emulated on one device run on different streams of one memory pool: the pool must not make one
        // stream wait for another to reuse freed scratch (a rank waiting in a barrier would never be released)
        cudaMemPool_t pool;
        if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess) {
            int off = 0;
            cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &off);
            cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowOpportunistic, &off);
        }
    }
    c->xbuf[peer] = peer_comm->xbuf[peer];
    c->flags[peer] = peer_comm->flags[peer];
    c->ipc[peer] = false;
    return comm_publish_tab(c);
}
size_t lg_comm_xbuf_words(const lg_comm* c) { return c ? c->xwords : 0; }
// words of exchange buffer that MulRelin + Rescale + gather of both polynomials need for `batch` ciphertexts (host only)
size_t lg_comm_xbuf_words_needed(uint64_t N, int nQ, int nP, int batch) {
    const size_t half = (size_t)batch * N * ((size_t)nQ + 2 * (size_t)nP + 2 + 2 * (size_t)nQ) + 64;
    return 2 * half;
}
int lg_comm_check(lg_comm* c, lg_stream_t s) {
    LG_REQUIRE(c, "lg_comm_check: null argument");
    LG_ON_DEVICE(c->device);
    LG_CUDA_CHECK(cudaStreamSynchronize(cs(s)));
    if (!c->d_err) return LG_OK;
    u32 err = 0;
    LG_CUDA_CHECK(cudaMemcpy(&err, c->d_err, sizeof(u32), cudaMemcpyDeviceToHost));
    if (err) {
        lg_set_error("limb-axis barrier timed out: a peer rank did not arrive within 5 s");
        return LG_ERR_CUDA;
    }
    return LG_OK;
}

// AggregateShares of the dckks/dbfv protocols across ranks (one party, or one group of parties, per GPU):
// p = Reduce(sum over ranks of p).  Equals the reference's chain of context.Add (dckks/publickey_gen.go:45-47) for
// canonical shares; world <= 8 and q < 2^61 keep the 64-bit sum from overflowing.
int lg_comm_aggregate_shares(const lg_comm* c, const lg_ring* r, int nl, lg_poly* p, lg_stream_t s) {
    LG_REQUIRE(c && r && p, "AggregateShares: null argument");
    LG_REQUIRE(c->world <= 8, "AggregateShares: at most 8 ranks (64-bit lazy sum)");
    LG_REQUIRE(nl >= 1 && nl <= r->nl && nl <= p->nlimbs && p->N == r->N, "AggregateShares: shape mismatch");
    LG_SAME_DEVICE("AggregateShares", c->device, r->device);
    LG_SAME_DEVICE("AggregateShares", c->device, p->device);
    LG_ON_DEVICE(c->device);
    if (c->world > 1) {
        LG_REQUIRE(c->comm, "AggregateShares: the communicator was created without a NCCL id");
        NcclApi* n = nccl();
        LG_NCCL_CHECK(n->GroupStart());
        int first = 0;  // an error inside the group must not leave it open
        for (int bt = 0; bt < p->batch && !first; ++bt) {
            u64* ptr = p->d + (size_t)bt * p->bstride;
            first = n->AllReduce(ptr, ptr, (size_t)nl * r->N, kNcclUint64, kNcclSum, c->comm, cs(s));
        }
        const int end = n->GroupEnd();
        LG_NCCL_CHECK(first);
        LG_NCCL_CHECK(end);
    }
    return lgi_ew(EW_REDUCE, r, limb_map_identity(), nl, p->batch, p->d, p->bstride, nullptr, 0, p->d, p->bstride, nullptr, 0,
                  cs(s));
}

// ---- limb-resident ops ------------------------------------------------------------------------------------------
int lg_comm_gather_limbs(lg_comm* c, const lg_ring* r, int nl, lg_poly* p, lg_stream_t s) {
    LG_REQUIRE(c && r && p, "GatherLimbs: null argument");
    LG_REQUIRE(nl >= 1 && nl <= p->nlimbs && p->N == r->N, "GatherLimbs: shape mismatch");
    LG_SAME_DEVICE("GatherLimbs", c->device, p->device);
    LG_ON_DEVICE(c->device);
    if (c->world == 1) return LG_OK;
    LG_TRY(op_begin(c));
    size_t off;
    LG_TRY(xalloc(c, (size_t)p->batch * nl * r->N, &off));
    return gather_limbs(r, c, nl, p->batch, p->d, p->bstride, off, cs(s));
}

int lg_ckks_switch_keys_in_place_resident(lg_ckks_eval* e, lg_comm* c, int level, const lg_poly* cx, const lg_swk* evk, lg_poly* p0,
                                          lg_poly* p1, lg_stream_t s) {
    LG_REQUIRE(e && c, "switchKeysInPlace: null argument");
    LG_SAME_DEVICE("switchKeysInPlace", c->device, e->Q->device);
    LG_ON_DEVICE(c->device);
    const u64 N = e->Q->N;
    LG_TRY(check_p(cx, N, level + 1, -1, "switchKeysInPlace"));
    LG_TRY(check_p(p0, N, level + 1, cx->batch, "switchKeysInPlace"));
    LG_TRY(check_p(p1, N, level + 1, cx->batch, "switchKeysInPlace"));
    LG_TRY(op_begin(c));
    return switch_keys_resident(e, c, level, cx->batch, cx->d + (size_t)c->rank * N, cx->bstride, (size_t)c->world * N, evk, p0->d,
                                p0->bstride, p1->d, p1->bstride, false, false, false, cs(s));
}

// MulRelin (ckks/evaluator.go:1016-1133) followed by nrescale Rescale steps (:933-968) on limb-resident ciphertexts
int lg_ckks_mul_relin_rescale_resident(lg_ckks_eval* e, lg_comm* c, int level, const lg_poly* a0, const lg_poly* a1,
                                       const lg_poly* b0, const lg_poly* b1, const lg_swk* rlk, lg_poly* out0, lg_poly* out1,
                                       int nrescale, lg_stream_t s) {
    LG_REQUIRE(e && c, "MulRelin: null argument");
    LG_SAME_DEVICE("MulRelin", c->device, e->Q->device);
    LG_ON_DEVICE(c->device);
    const lg_ring* Q = e->Q;
    const u64 N = Q->N;
    const int nl = level + 1;
    LG_REQUIRE(level >= 0 && nl <= Q->nl, "MulRelin: level %d out of range", level);
    LG_REQUIRE(nrescale >= 0 && nrescale < nl, "cannot Rescale: input Ciphertext already at level 0");
    LG_TRY(check_p(a0, N, nl, -1, "MulRelin"));
    const int batch = a0->batch;
    LG_TRY(check_p(a1, N, nl, batch, "MulRelin"));
    LG_TRY(check_p(b0, N, nl, batch, "MulRelin"));
    LG_TRY(check_p(b1, N, nl, batch, "MulRelin"));
    LG_TRY(check_p(out0, N, nl, batch, "MulRelin"));
    LG_TRY(check_p(out1, N, nl, batch, "MulRelin"));
    cudaStream_t st = cs(s);
    LG_TRY(op_begin(c));
    const int n0q = count_own(nl, c->world, c->rank);
    const size_t c2_bs = (size_t)(n0q > 0 ? n0q : 1) * N;
    Scratch w(st);
    LG_TRY(w.alloc((size_t)batch * c2_bs));
    LG_TRY(tensor_own(e, c, nl, a0, a1, b0, b1, out0, out1, w.d, c2_bs, st));
    // :1098-1104 relinearise c2 (canonical: MRed output of the tensor) and add into c0, c1
    LG_TRY(switch_keys_resident(e, c, level, batch, w.d, c2_bs, N, rlk, out0->d, out0->bstride, out1->d, out1->bstride, true, true,
                                true, st));
    for (int k = 0; k < nrescale; ++k)
        LG_TRY(rescale_resident(e, c, nl - k, batch, out0->d, out0->bstride, out1->d, out1->bstride, st));
    return LG_OK;
}

int lg_ckks_rescale_resident(lg_ckks_eval* e, lg_comm* c, int nl, lg_poly* c0, lg_poly* c1, lg_stream_t s) {
    LG_REQUIRE(e && c, "Rescale: null argument");
    LG_SAME_DEVICE("Rescale", c->device, e->Q->device);
    LG_ON_DEVICE(c->device);
    LG_TRY(check_p(c0, e->Q->N, nl, -1, "Rescale"));
    LG_TRY(check_p(c1, e->Q->N, nl, c0->batch, "Rescale"));
    LG_TRY(op_begin(c));
    return rescale_resident(e, c, nl, c0->batch, c0->d, c0->bstride, c1->d, c1->bstride, cs(s));
}

// ---- replicated-in / replicated-out forms (round 1 interface): the resident op, then the result limbs are gathered ---
int lg_ckks_switch_keys_in_place_sharded(lg_ckks_eval* e, lg_comm* c, int level, const lg_poly* cx, const lg_swk* evk,
                                         lg_poly* p0, lg_poly* p1, lg_stream_t s) {
    LG_TRY(lg_ckks_switch_keys_in_place_resident(e, c, level, cx, evk, p0, p1, s));
    LG_ON_DEVICE(c->device);
    return gather_pair(e, c, level + 1, p0, p1, cs(s));
}

int lg_ckks_mul_relin_sharded(lg_ckks_eval* e, lg_comm* c, int level, const lg_poly* a0, const lg_poly* a1, const lg_poly* b0,
                              const lg_poly* b1, const lg_swk* rlk, lg_poly* out0, lg_poly* out1, lg_stream_t s) {
    LG_TRY(lg_ckks_mul_relin_rescale_resident(e, c, level, a0, a1, b0, b1, rlk, out0, out1, 0, s));
    LG_ON_DEVICE(c->device);
    return gather_pair(e, c, level + 1, out0, out1, cs(s));
}

int lg_ckks_rescale_sharded(lg_ckks_eval* e, lg_comm* c, int nl, lg_poly* c0, lg_poly* c1, lg_stream_t s) {
    LG_TRY(lg_ckks_rescale_resident(e, c, nl, c0, c1, s));
    LG_ON_DEVICE(c->device);
    return gather_pair(e, c, nl - 1, c0, c1, cs(s));
}

}  // extern "C"
