// capi_shard.cu -- C-ABI host layer, part 4: multi-GPU paths.
//
// (1) lg_comm: a thin handle over NCCL (resolved with dlopen so that the library the host process
//     already uses -- e.g. the one torch.distributed loaded -- is shared).
// (2) Limb-sharded CKKS key switch / MulRelin / Rescale for ONE ciphertext (or a small batch) spread
//     over the GPUs of a node (BASELINE config 4, SURVEY.md 8(e) "limb axis"): every rank holds the
//     full ciphertext (replicated in, replicated out) and owns a contiguous block of the
//     level+1+#P data limbs.  NTT, multiply-accumulate and the ModDown tail are limb-local; the only
//     exchanges are all-gathers over NVLink exactly where a basis extension needs every source limb:
//     the coefficient-domain c2 before DecomposeAndSplit, the special-prime accumulators before
//     ModDown, and the result limbs at the end.
// (3) Share aggregation for the dckks/dbfv protocols (config 5, "party axis"): AggregateShares is an
//     all-reduce(sum, u64) followed by one Reduce, which equals the reference's pairwise CRed-add
//     chain (dckks/publickey_gen.go:45-47) for up to 8 canonical shares of < 2^61.
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
    int (*Broadcast)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
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
            LG_SYM(Broadcast, "ncclBroadcast");
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

struct Range {
    int b, e;
    int n() const { return e - b; }
};
Range own_range(int n, int world, int rank) {
    return Range{(int)((long long)rank * n / world), (int)((long long)(rank + 1) * n / world)};
}
Range clip(Range r, int lo, int hi) {  // intersection with [lo,hi), shifted to start at lo
    Range o{r.b < lo ? lo : r.b, r.e > hi ? hi : r.e};
    if (o.e < o.b) o.e = o.b;
    return Range{o.b - lo, o.e - lo};
}
LimbMap sub_map(LimbMap m, int b) {
    if (b < m.n0) return LimbMap{m.n0 - b, m.l0 + b, m.l1};
    return LimbMap{0, 0, m.l1 + (b - m.n0)};
}
}  // namespace

struct lg_comm {
    int device = -1;  // the rank's GPU
    int world = 1, rank = 0;
    NcclComm comm = nullptr;
};

namespace {

// In-place all-gather of limb blocks: rank r contributes limbs ranges[r] of every batch entry of
// `base` (limb stride N, batch stride bstride).  Uneven blocks -> grouped broadcasts.
int allgather_limbs(const lg_comm* c, u64* base, size_t bstride, int batch, u64 N, const std::vector<Range>& ranges,
                    cudaStream_t st) {
    if (c->world == 1) return LG_OK;
    NcclApi* n = nccl();
    LG_NCCL_CHECK(n->GroupStart());
    int first = 0;  // an error inside the group must not leave it open
    for (int r = 0; r < c->world && !first; ++r) {
        if (ranges[r].n() <= 0) continue;
        for (int bt = 0; bt < batch && !first; ++bt) {
            u64* p = base + (size_t)bt * bstride + (size_t)ranges[r].b * N;
            first = n->Broadcast(p, p, (size_t)ranges[r].n() * N, kNcclUint64, r, c->comm, st);
        }
    }
    const int end = n->GroupEnd();
    LG_NCCL_CHECK(first);
    LG_NCCL_CHECK(end);
    return LG_OK;
}

std::vector<Range> all_ranges(const lg_comm* c, int n, int lo, int hi) {
    std::vector<Range> v;
    for (int r = 0; r < c->world; ++r) v.push_back(clip(own_range(n, c->world, r), lo, hi));
    return v;
}

// Decompose(AndSplit) (ring_basis_extension.go:476-713) restricted to the target limbs this rank
// owns: Q targets q (indices into 0..level) and special primes p (indices into 0..nP-1).
// dbuf: [batch][level+1 | nP][N].
int decompose_range(const lg_decomposer* d, int level, int crt, int batch, const u64* p0, size_t p0_bs, u64* dbuf, size_t d_bs,
                    Range q, Range p, cudaStream_t st) {
    const int nl = level + 1;
    const int alphai = d->xalpha[crt];
    const int p0idxst = crt * d->alpha;
    const int p0idxed = p0idxst + alphai;
    const u64 N = d->N;
    u64* outQ = dbuf + (size_t)q.b * N;
    u64* outP = dbuf + (size_t)(nl + p.b) * N;
    if ((p0idxed > level + 1 && (level + 1) % d->nP == 1) || alphai == 1) {
        FanoutArgs f;
        f.N = (u32)N;
        f.in = p0 + (size_t)p0idxst * N;
        f.in_bs = p0_bs;
        f.nruns = 2;
        f.out[0] = outQ;
        f.out_bs[0] = d_bs;
        f.ndst[0] = q.n();
        f.out[1] = outP;
        f.out_bs[1] = d_bs;
        f.ndst[1] = p.n();
        f.mode = 0;
        f.phalf = f.plast = 0;
        lg_launch_fanout(f, batch, st);
        LG_LAUNCH_CHECK();
        return LG_OK;
    }
    int index = (level >= alphai + crt * d->alpha) ? d->xalpha[crt] - 2 : (level - 1) % d->alpha;
    LG_REQUIRE(index >= 0 && index < (int)d->modup[crt].size(), "Decompose: no parameters for digit %d index %d", crt, index);
    const ModUpDev& m = *d->modup[crt][index];
    ModUpArgs a;
    memset(&a, 0, sizeof(a));
    a.M = m.M;
    a.N = (u32)N;
    a.nsrc = index + 2;
    a.in = p0 + (size_t)p0idxst * N;
    a.in_bs = p0_bs;
    a.nruns = 2;
    a.out[0] = outQ;
    a.out_bs[0] = d_bs;
    a.ndst[0] = q.n();
    a.tgt0[0] = q.b;
    a.out[1] = outP;
    a.out_bs[1] = d_bs;
    a.ndst[1] = p.n();
    a.tgt0[1] = d->nQ + p.b;
    a.fast = m.fast_level(a.nsrc, &a.fp_shift);
    LG_REQUIRE(lg_launch_modup(a, batch, st) == 0, "Decompose: too many source limbs");
    LG_LAUNCH_CHECK();
    return LG_OK;
}

// Limb-sharded switchKeysInPlace (ckks/evaluator.go:1475-1558).  cx must hold valid NTT-domain data in
// this rank's own Q limbs (the other limbs are not read).  out0/out1: the rank's own Q limbs are
// written (or accumulated into); `gather_out` then replicates them on every rank.
int switch_keys_sharded(lg_ckks_eval* e, const lg_comm* c, int level, int batch, const u64* cx, size_t cx_bs, const lg_swk* evk,
                        u64* out0, size_t out0_bs, u64* out1, size_t out1_bs, bool add0, bool add1, bool gather_out,
                        cudaStream_t st) {
    const lg_ring* Q = e->Q;
    const lg_ring* P = e->P;
    const lg_ring* QP = e->QP.get();
    const u64 N = Q->N;
    const int nQ = Q->nl, nP = P->nl, nl = level + 1, nd = nl + nP;
    LG_REQUIRE(Q->logN >= 12, "sharded key switch needs N >= 2^12");
    LG_REQUIRE(level >= 0 && level < nQ, "switchKeys: level %d out of range", level);
    LG_REQUIRE(evk && evk->N == N && evk->nQP == nQ + nP, "switchKeys: switching key shape mismatch");
    const int alpha = e->alpha, beta = (nl + alpha - 1) / alpha;
    LG_REQUIRE(beta <= evk->beta, "switchKeys: key has %d digits, %d needed", evk->beta, beta);
    const Range mine = own_range(nd, c->world, c->rank);
    const Range myq = clip(mine, 0, nl), myp = clip(mine, nl, nd);
    const LimbMap qp_map{nl, 0, nQ};
    const LimbMap idm = limb_map_identity();

    Scratch c2(st), acc(st), tmp(st);
    LG_TRY(c2.alloc((size_t)batch * nl * N));
    LG_TRY(acc.alloc((size_t)2 * batch * nd * N));
    LG_TRY(tmp.alloc((size_t)2 * batch * nl * N));
    const size_t c2_bs = (size_t)nl * N, d_bs = (size_t)nd * N;
    u64* acc0 = acc.d;
    u64* acc1 = acc.d + (size_t)batch * d_bs;

    // :1503 c2 = InvNTT(cx) on the own Q limbs, then all-gather: DecomposeAndSplit needs every source limb
    if (myq.n() > 0)
        LG_TRY(lgi_ntt(Q, sub_map(idm, myq.b), myq.n(), batch, cx + (size_t)myq.b * N, cx_bs, c2.d + (size_t)myq.b * N, c2_bs,
                       true, 0, 0, st));
    LG_TRY(allgather_limbs(c, c2.d, c2_bs, batch, N, all_ranges(c, nd, 0, nl), st));

    // :1511-1552 digit loop on the own target limbs: every digit decomposed and taken through the strided NTT
    // phase, then the fused contiguous-phase + multiply-accumulate kernel over the own limbs (register
    // accumulators across the digits, as on one GPU)
    if (mine.n() > 0) {
        Scratch D(st);
        LG_TRY(D.alloc((size_t)beta * batch * d_bs));
        const size_t d_ds = (size_t)batch * d_bs;
        for (int i = 0; i < beta; ++i)
            LG_TRY(decompose_range(e->dec.get(), level, i, batch, c2.d, c2_bs, D.d + (size_t)i * d_ds, d_bs, myq, myp, st));
        NttArgs a;
        memset(&a, 0, sizeof(a));
        a.T = QP->T;
        a.map = sub_map(qp_map, mine.b);
        a.in = D.d + (size_t)mine.b * N;
        a.out = D.d + (size_t)mine.b * N;
        a.in_bstride = a.out_bstride = d_bs;
        a.skip_alpha = alpha;
        a.skip_div = batch;
        a.skip_nl = nl;
        a.skip_limb0 = mine.b;
        LG_REQUIRE(lg_launch_ntt_fwd_strided(a, mine.n(), beta * batch, st) == 0, "switchKeys: strided NTT launch failed");
        LG_LAUNCH_CHECK();
        KsFusedArgs k;
        memset(&k, 0, sizeof(k));
        k.T = QP->T;
        k.map = sub_map(qp_map, mine.b);
        k.D = D.d + (size_t)mine.b * N;
        k.d_ds = d_ds;
        k.d_bs = d_bs;
        k.cx = cx + (size_t)mine.b * N;
        k.cx_bs = cx_bs;
        k.evk = evk->key(0, 0);
        k.evk_ds = (size_t)(evk->key(1, 0) - evk->key(0, 0));
        k.evk_hs = (size_t)(evk->key(0, 1) - evk->key(0, 0));
        k.acc0 = acc0 + (size_t)mine.b * N;
        k.acc1 = acc1 + (size_t)mine.b * N;
        k.acc_bs = d_bs;
        k.beta = beta;
        k.alpha = alpha;
        k.nl = nl;
        k.limb0 = mine.b;
        LG_REQUIRE(lg_launch_ks_fused(k, mine.n(), batch, st) == 0, "switchKeys: fused digit loop launch failed");
        LG_LAUNCH_CHECK();
    }

    // :1556-1557 ModDownSplitedNTTPQ: InvNTT of the own special-prime limbs, all-gather them (modUpExact
    // P -> Q needs every P limb), then the own Q limbs
    if (myp.n() > 0)
        LG_TRY(lgi_ntt(P, sub_map(idm, myp.b), myp.n(), 2 * batch, acc0 + (size_t)(nl + myp.b) * N, d_bs,
                       acc0 + (size_t)(nl + myp.b) * N, d_bs, true, 0, 0, st));
    LG_TRY(allgather_limbs(c, acc0 + (size_t)nl * N, d_bs, 2 * batch, N, all_ranges(c, nd, nl, nd), st));
    if (myq.n() > 0) {
        const size_t t_bs = (size_t)nl * N;
        ModUpArgs a;
        memset(&a, 0, sizeof(a));
        a.M = e->ext->pq.M;
        a.N = (u32)N;
        a.nsrc = nP;
        a.in = acc0 + (size_t)nl * N;
        a.in_bs = d_bs;
        a.nruns = 1;
        a.out[0] = tmp.d + (size_t)myq.b * N;
        a.out_bs[0] = t_bs;
        a.ndst[0] = myq.n();
        a.tgt0[0] = myq.b;
        a.fast = e->ext->pq.fast_level(a.nsrc, &a.fp_shift);
        LG_REQUIRE(lg_launch_modup(a, 2 * batch, st) == 0, "modUpExact: too many source limbs");
        LG_LAUNCH_CHECK();
        u64* t0 = tmp.d + (size_t)myq.b * N;
        LG_TRY(lgi_ntt(Q, sub_map(idm, myq.b), myq.n(), 2 * batch, t0, t_bs, t0, t_bs, false, 0, 0, st));
        const u64* sc = e->ext->moddown_pq.data() + myq.b;
        LG_TRY(lgi_ew(add0 ? EW_SUB_MULMONT_SCALAR_ADD : EW_SUB_MULMONT_SCALAR, Q, sub_map(idm, myq.b), myq.n(), batch,
                      acc0 + (size_t)myq.b * N, d_bs, t0, t_bs, out0 + (size_t)myq.b * N, out0_bs, sc, myq.n(), st));
        LG_TRY(lgi_ew(add1 ? EW_SUB_MULMONT_SCALAR_ADD : EW_SUB_MULMONT_SCALAR, Q, sub_map(idm, myq.b), myq.n(), batch,
                      acc1 + (size_t)myq.b * N, d_bs, t0 + (size_t)batch * t_bs, t_bs, out1 + (size_t)myq.b * N, out1_bs, sc,
                      myq.n(), st));
    }
    if (gather_out) {
        const std::vector<Range> rq = all_ranges(c, nd, 0, nl);
        LG_TRY(allgather_limbs(c, out0, out0_bs, batch, N, rq, st));
        LG_TRY(allgather_limbs(c, out1, out1_bs, batch, N, rq, st));
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
    LG_REQUIRE(out && world >= 1 && rank >= 0 && rank < world, "lg_comm_create: invalid argument");
    std::unique_ptr<lg_comm> c(new lg_comm);
    c->device = lgi_current_device();
    c->world = world;
    c->rank = rank;
    if (world > 1) {
        LG_REQUIRE(id128, "lg_comm_create: null unique id");
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
    delete c;
    return LG_OK;
}
int lg_comm_world(const lg_comm* c) { return c ? c->world : 0; }
int lg_comm_rank(const lg_comm* c) { return c ? c->rank : -1; }

// the ownership rule: rank r owns limbs [r*n/world, (r+1)*n/world)   (host only)
int lg_comm_limb_range(int nlimbs, int world, int rank, int* begin, int* end) {
    LG_REQUIRE(begin && end && world >= 1 && rank >= 0 && rank < world && nlimbs >= 0, "lg_comm_limb_range: invalid argument");
    const Range r = own_range(nlimbs, world, rank);
    *begin = r.b;
    *end = r.e;
    return LG_OK;
}

// AggregateShares of the dckks/dbfv protocols across ranks (one party per GPU): p = Reduce(sum over
// ranks of p).  Equals the reference's chain of context.Add (dckks/publickey_gen.go:45-47) for
// canonical shares; world <= 8 and q < 2^61 keep the 64-bit sum from overflowing.
int lg_comm_aggregate_shares(const lg_comm* c, const lg_ring* r, int nl, lg_poly* p, lg_stream_t s) {
    LG_REQUIRE(c && r && p, "AggregateShares: null argument");
    LG_REQUIRE(c->world <= 8, "AggregateShares: at most 8 ranks (64-bit lazy sum)");
    LG_REQUIRE(nl >= 1 && nl <= r->nl && nl <= p->nlimbs && p->N == r->N, "AggregateShares: shape mismatch");
    LG_SAME_DEVICE("AggregateShares", c->device, r->device);
    LG_SAME_DEVICE("AggregateShares", c->device, p->device);
    LG_ON_DEVICE(c->device);
    if (c->world > 1) {
        NcclApi* n = nccl();
        LG_NCCL_CHECK(n->GroupStart());
        int first = 0;
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

int lg_ckks_switch_keys_in_place_sharded(lg_ckks_eval* e, const lg_comm* c, int level, const lg_poly* cx, const lg_swk* evk,
                                         lg_poly* p0, lg_poly* p1, lg_stream_t s) {
    LG_REQUIRE(e && c, "switchKeysInPlace: null argument");
    LG_SAME_DEVICE("switchKeysInPlace", c->device, e->Q->device);
    LG_ON_DEVICE(c->device);
    const u64 N = e->Q->N;
    LG_TRY(check_p(cx, N, level + 1, -1, "switchKeysInPlace"));
    LG_TRY(check_p(p0, N, level + 1, cx->batch, "switchKeysInPlace"));
    LG_TRY(check_p(p1, N, level + 1, cx->batch, "switchKeysInPlace"));
    return switch_keys_sharded(e, c, level, cx->batch, cx->d, cx->bstride, evk, p0->d, p0->bstride, p1->d, p1->bstride, false,
                               false, true, cs(s));
}

// MulRelin (ckks/evaluator.go:1016-1133) with the limbs of the ciphertext spread over the ranks of `c`.
// Inputs and outputs are replicated on every rank.
int lg_ckks_mul_relin_sharded(lg_ckks_eval* e, const lg_comm* c, int level, const lg_poly* a0, const lg_poly* a1,
                              const lg_poly* b0, const lg_poly* b1, const lg_swk* rlk, lg_poly* out0, lg_poly* out1,
                              lg_stream_t s) {
    LG_REQUIRE(e && c, "MulRelin: null argument");
    LG_SAME_DEVICE("MulRelin", c->device, e->Q->device);
    LG_ON_DEVICE(c->device);
    const lg_ring* Q = e->Q;
    const u64 N = Q->N;
    const int nl = level + 1, nd = nl + e->P->nl;
    LG_REQUIRE(level >= 0 && nl <= Q->nl, "MulRelin: level %d out of range", level);
    LG_TRY(check_p(a0, N, nl, -1, "MulRelin"));
    const int batch = a0->batch;
    LG_TRY(check_p(a1, N, nl, batch, "MulRelin"));
    LG_TRY(check_p(b0, N, nl, batch, "MulRelin"));
    LG_TRY(check_p(b1, N, nl, batch, "MulRelin"));
    LG_TRY(check_p(out0, N, nl, batch, "MulRelin"));
    LG_TRY(check_p(out1, N, nl, batch, "MulRelin"));
    cudaStream_t st = cs(s);
    const Range myq = clip(own_range(nd, c->world, c->rank), 0, nl);
    const size_t bs = (size_t)nl * N;
    Scratch w(st);
    LG_TRY(w.alloc((size_t)batch * bs));
    if (myq.n() > 0) {  // :1076-1095 tensor on the own limbs: c0 -> out0, c1 -> out1, c2 -> scratch
        TensorArgs t;
        t.T = Q->T;
        t.a0 = a0->d;
        t.a1 = a1->d;
        t.b0 = b0->d;
        t.b1 = b1->d;
        t.c0 = out0->d;
        t.c1 = out1->d;
        t.c2 = w.d;
        t.a_bs[0] = a0->bstride;
        t.a_bs[1] = a1->bstride;
        t.b_bs[0] = b0->bstride;
        t.b_bs[1] = b1->bstride;
        t.c_bs[0] = out0->bstride;
        t.c_bs[1] = out1->bstride;
        t.c_bs[2] = bs;
        t.square = (a0->d == b0->d && a1->d == b1->d) ? 1 : 0;
        t.nomod = 0;
        t.limb0 = myq.b;
        lg_launch_tensor(t, myq.n(), batch, st);
        LG_LAUNCH_CHECK();
    }
    return switch_keys_sharded(e, c, level, batch, w.d, bs, rlk, out0->d, out0->bstride, out1->d, out1->bstride, true, true,
                               true, st);
}

// Rescale loop body (ckks/evaluator.go:955-960) with the lower limbs spread over the ranks.  The last limb
// is inverse-transformed redundantly on every rank (one limb), so no exchange precedes the fan-out.
int lg_ckks_rescale_sharded(lg_ckks_eval* e, const lg_comm* c, int nl, lg_poly* c0, lg_poly* c1, lg_stream_t s) {
    LG_REQUIRE(e && c, "Rescale: null argument");
    LG_SAME_DEVICE("Rescale", c->device, e->Q->device);
    LG_ON_DEVICE(c->device);
    const lg_ring* Q = e->Q;
    const u64 N = Q->N;
    LG_TRY(check_p(c0, N, nl, -1, "Rescale"));
    LG_TRY(check_p(c1, N, nl, c0->batch, "Rescale"));
    LG_REQUIRE(nl >= 2 && nl <= Q->nl, "cannot Rescale: input Ciphertext already at level 0");
    LG_REQUIRE(Q->logN >= 12, "sharded rescale needs N >= 2^12");
    cudaStream_t st = cs(s);
    const int level = nl - 1, batch = c0->batch;
    const Range mine = own_range(level, c->world, c->rank);
    const LimbMap idm = limb_map_identity();
    Scratch tmp(st);
    LG_TRY(tmp.alloc((size_t)batch * level * N));
    const size_t tbs = (size_t)level * N;
    lg_poly* polys[2] = {c0, c1};
    for (lg_poly* p : polys) {
        u64* last = p->d + (size_t)level * N;
        LG_TRY(lgi_ntt(Q, LimbMap{1 << 30, level, 0}, 1, batch, last, p->bstride, last, p->bstride, true, 0, 0, st));  // ring_scaling.go:80
        if (mine.n() > 0) {
            FanoutArgs f;
            f.N = (u32)N;
            f.in = last;
            f.in_bs = p->bstride;
            f.nruns = 1;
            f.out[0] = tmp.d + (size_t)mine.b * N;
            f.out_bs[0] = tbs;
            f.ndst[0] = mine.n();
            f.mode = 1;
            const u64 phalf = (Q->q[level] - 1) >> 1;
            f.phalf = phalf;
            f.plast = Q->q[level];
            for (int i = 0; i < mine.n(); ++i) f.add[i] = Q->q[mine.b + i] - (phalf % Q->q[mine.b + i]);
            lg_launch_fanout(f, batch, st);
            LG_LAUNCH_CHECK();
            u64* t0 = tmp.d + (size_t)mine.b * N;
            LG_TRY(lgi_ntt(Q, sub_map(idm, mine.b), mine.n(), batch, t0, tbs, t0, tbs, false, 0, 0, st));
            std::vector<u64> sc(mine.n());
            for (int i = 0; i < mine.n(); ++i) sc[i] = Q->rescale_param(level, mine.b + i);
            LG_TRY(lgi_ew(EW_SUB_MULMONT_SCALAR, Q, sub_map(idm, mine.b), mine.n(), batch, p->d + (size_t)mine.b * N, p->bstride,
                          t0, tbs, p->d + (size_t)mine.b * N, p->bstride, sc.data(), mine.n(), st));
        }
        std::vector<Range> rq;
        for (int r = 0; r < c->world; ++r) rq.push_back(own_range(level, c->world, r));
        LG_TRY(allgather_limbs(c, p->d, p->bstride, batch, N, rq, st));
    }
    return LG_OK;
}

}  // extern "C"
