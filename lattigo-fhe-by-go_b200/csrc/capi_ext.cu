// capi_ext.cu -- C-ABI host layer, part 2: FastBasisExtender, Decomposer and the
// CKKS evaluator key-switch path.  Host-side mirror of
// ring/ring_basis_extension.go and ckks/evaluator.go:933-1591 (hot ops only).
#include <stdlib.h>
#include <string.h>

#include "capi_internal.hpp"

static inline cudaStream_t cs(lg_stream_t s) { return (cudaStream_t)s; }

// ---------------------------------------------------------------------------
// basisextenderparameters, ring_basis_extension.go:76-142.  The reference uses
// math/big for Q/q_i, its inverse mod q_i and the residues mod p_j; they are
// canonical residues and are computed here with 128-bit modular products.
// ---------------------------------------------------------------------------
int ModUpDev::build(const u64* Q, int nq, const u64* P, int np) {
    using namespace lgh;
    nsrc = nq;
    ndst = np;
    small = true;
    hsrc.assign(Q, Q + nq);
    hdst.assign(P, P + np);
    for (int i = 0; i < nq; ++i) small = small && (Q[i] >> 61) == 0;
    for (int j = 0; j < np; ++j) small = small && (P[j] >> 61) == 0;
    std::vector<u64> sQ(Q, Q + nq), sQinv(nq), vqib(nq), vqispj((size_t)nq * np), vqpj((size_t)np * (nq + 1));
    std::vector<u64> dQ(P, P + np), dQinv(np), dU0(np);
    for (int i = 0; i < nq; ++i) sQinv[i] = mred_params(Q[i]);
    for (int j = 0; j < np; ++j) {
        dQinv[j] = mred_params(P[j]);
        u64 lo;
        bred_params(P[j], dU0[j], lo);
    }
    for (int i = 0; i < nq; ++i) {
        const u64 qi = Q[i];
        u64 star = 1 % qi;
        for (int k = 0; k < nq; ++k)
            if (k != i) star = mulmod(star, Q[k] % qi, qi);
        vqib[i] = mform(powmod(star, qi - 2, qi), qi);  // :115-118 (QiBarre = QiStar^-1 mod qi)
        for (int j = 0; j < np; ++j) {
            const u64 pj = P[j];
            u64 s = 1 % pj;
            for (int k = 0; k < nq; ++k)
                if (k != i) s = mulmod(s, Q[k] % pj, pj);
            vqispj[(size_t)i * np + j] = mform(s, pj);  // :121-123
        }
    }
    for (int j = 0; j < np; ++j) {  // :128-138
        const u64 pj = P[j];
        u64 qm = 1 % pj;
        for (int k = 0; k < nq; ++k) qm = mulmod(qm, Q[k] % pj, pj);
        const u64 v = pj - qm;
        u64* row = vqpj.data() + (size_t)j * (nq + 1);
        row[0] = 0;
        for (int i = 1; i <= nq; ++i) {
            const u64 t = row[i - 1] + v;
            row[i] = t >= pj ? t - pj : t;
        }
    }
    LG_TRY(srcQ.upload(sQ));
    LG_TRY(srcQinv.upload(sQinv));
    LG_TRY(qib.upload(vqib));
    LG_TRY(qispj.upload(vqispj));
    LG_TRY(qpjinv.upload(vqpj));
    LG_TRY(dstQ.upload(dQ));
    LG_TRY(dstQinv.upload(dQinv));
    LG_TRY(dstU0.upload(dU0));
    M.srcQ = srcQ.d;
    M.srcQinv = srcQinv.d;
    M.qib = qib.d;
    M.qispj = qispj.d;
    M.qpjinv = qpjinv.d;
    M.dstQ = dstQ.d;
    M.dstQinv = dstQinv.d;
    M.dstU0 = dstU0.d;
    M.src_total = nq;
    M.dst_total = np;
    return LG_OK;
}

// modUpExact (:352-393) on `nsrc` source limbs into `ndst` target limbs tgt0..
int lgi_modup_launch(const ModUpDev& m, u64 N, int batch, const u64* in, size_t in_bs, int nsrc, u64* out,
                        size_t out_bs, int ndst, int tgt0, cudaStream_t st, bool lazy_out) {
    LG_REQUIRE(nsrc >= 1 && nsrc <= m.nsrc && ndst >= 0 && tgt0 + ndst <= m.ndst, "modUpExact: basis out of range");
    ModUpArgs a;
    memset(&a, 0, sizeof(a));
    a.M = m.M;
    a.N = (u32)N;
    a.nsrc = nsrc;
    a.in = in;
    a.in_bs = in_bs;
    a.nruns = 1;
    a.out[0] = out;
    a.out_bs[0] = out_bs;
    a.ndst[0] = ndst;
    a.tgt0[0] = tgt0;
    a.copy_out = nullptr;
    a.fast = m.fast_level(a.nsrc, &a.fp_shift);
    a.lazy_out = lazy_out ? 1 : 0;
    if (lg_launch_modup(a, batch, st) != 0) {
        lg_set_error("modUpExact: too many source limbs (%d)", nsrc);
        return LG_ERR_ARG;
    }
    LG_LAUNCH_CHECK();
    return LG_OK;
}

// genModDownParams, ring_basis_extension.go:39-53
static std::vector<u64> gen_moddown(const lg_ring* a, const lg_ring* b) {
    std::vector<u64> out(a->nl);
    for (int i = 0; i < a->nl; ++i) {
        const u64 qi = a->q[i];
        u64 m = 1 % qi;
        for (int k = 0; k < b->nl; ++k) m = lgh::mulmod(m, b->q[k] % qi, qi);
        out[i] = lgh::mform(lgh::powmod(m, qi - 2, qi), qi);
    }
    return out;
}

static int check_p(const lg_poly* p, u64 N, int nl, int batch, const char* what) {
    LG_REQUIRE(p, "%s: null polynomial", what);
    LG_REQUIRE(p->N == N, "%s: degree mismatch", what);
    LG_REQUIRE(p->nlimbs >= nl, "%s: polynomial has %d limbs, %d needed", what, p->nlimbs, nl);
    LG_REQUIRE(batch < 0 || p->batch == batch, "%s: batch mismatch", what);
    LG_SAME_DEVICE(what, lgi_expected_device(), p->device);
    return LG_OK;
}

static bool no_tail_canon() { return lg_switches().no_tail_canon.load(std::memory_order_relaxed) != 0; }

// Shared tail of every ModDown*: tmp = modUp(P part -> Q[:level+1]); optionally
// NTT(tmp); p2 = MRed(p1Q + (q - tmp), P^-1)   (:219-240, :254-273, :287-306)
// accumulate = true adds the result into p2 with CRed (the AddLvl the evaluators apply right after).
int lgi_moddown_tail_ntt(const lg_extender* e, int level, int batch, const u64* p1Q, size_t p1Q_bs, u64* p1P,
                         size_t p1P_bs, u64* p2, size_t p2_bs, bool ntt, cudaStream_t st, bool accumulate, bool p_in_range) {
    const lg_ring* Q = e->Q;
    const lg_ring* P = e->P;
    const u64 N = Q->N;
    const int nl = level + 1;
    LG_REQUIRE(level >= 0 && nl <= Q->nl, "ModDown: level %d out of range", level);
    if (ntt)  // :172 / :215 -- destroys the P part of the input, like the reference
        LG_TRY(lgi_ntt(P, limb_map_identity(), P->nl, batch, p1P, p1P_bs, p1P, p1P_bs, true, 0, 0, st, p_in_range));
    Scratch tmp(st);
    LG_TRY(tmp.alloc((size_t)batch * nl * N));
    const size_t tbs = (size_t)nl * N;
    // (with `ntt` the forward transform is the only reader of tmp: the conversion may stay in [0, 2q))
    LG_TRY(lgi_modup_launch(e->pq, N, batch, p1P, p1P_bs, P->nl, tmp.d, tbs, nl, 0, st, ntt));
    if (ntt && lgi_ntt_tail_ok(Q)) {  // :228-238 in one pass
        NttTail t;
        memset(&t, 0, sizeof(t));
        t.enabled = 1;
        t.split = batch;
        t.add[0] = accumulate ? 1 : 0;
        t.a_canon = (p_in_range && !no_tail_canon()) ? 1 : 0;  // internal callers only: canonical key-switch accumulators
        t.a[0] = p1Q;
        t.a_bs[0] = p1Q_bs;
        t.out[0] = p2;
        t.out_bs[0] = p2_bs;
        t.s = e->d_moddown_pq.d;
        return lgi_ntt(Q, limb_map_identity(), nl, batch, tmp.d, tbs, tmp.d, tbs, false, 0, 0, st, false, &t);
    }
    if (ntt) LG_TRY(lgi_ntt(Q, limb_map_identity(), nl, batch, tmp.d, tbs, tmp.d, tbs, false, 0, 0, st));
    return lgi_ew(accumulate ? EW_SUB_MULMONT_SCALAR_ADD : EW_SUB_MULMONT_SCALAR, Q, limb_map_identity(), nl, batch, p1Q,
                  p1Q_bs, tmp.d, tbs, p2, p2_bs, e->moddown_pq.data(), nl, st);
}

// Two ModDowns whose inputs sit back to back ([2][batch][...] accumulators of a key switch): InvNTT, modUpExact and
// NTT run once over 2*batch entries, only the final (x - y) * P^-1 (+ add) differs per output.
int lgi_moddown_pair_ntt(const lg_extender* e, int level, int batch, u64* acc0, u64* acc1, size_t acc_bs, int p_off, u64* out0,
                         size_t out0_bs, bool add0, u64* out1, size_t out1_bs, bool add1, bool ntt, cudaStream_t st,
                         bool p_in_range) {
    const lg_ring* Q = e->Q;
    const lg_ring* P = e->P;
    const u64 N = Q->N;
    const int nl = level + 1;
    LG_REQUIRE(level >= 0 && nl <= Q->nl, "ModDown: level %d out of range", level);
    if (acc1 != acc0 + (size_t)batch * acc_bs) {  // not contiguous: two independent passes
        LG_TRY(lgi_moddown_tail_ntt(e, level, batch, acc0, acc_bs, acc0 + (size_t)p_off * N, acc_bs, out0, out0_bs, ntt, st, add0,
                                    p_in_range));
        return lgi_moddown_tail_ntt(e, level, batch, acc1, acc_bs, acc1 + (size_t)p_off * N, acc_bs, out1, out1_bs, ntt, st, add1,
                                    p_in_range);
    }
    u64* pP = acc0 + (size_t)p_off * N;
    if (ntt) LG_TRY(lgi_ntt(P, limb_map_identity(), P->nl, 2 * batch, pP, acc_bs, pP, acc_bs, true, 0, 0, st, p_in_range));
    Scratch tmp(st);
    LG_TRY(tmp.alloc((size_t)2 * batch * nl * N));
    const size_t tbs = (size_t)nl * N;
    LG_TRY(lgi_modup_launch(e->pq, N, 2 * batch, pP, acc_bs, P->nl, tmp.d, tbs, nl, 0, st, ntt));
    if (ntt && lgi_ntt_tail_ok(Q)) {  // both tails ride on the last phase of the one transform over 2*batch entries
        NttTail t;
        memset(&t, 0, sizeof(t));
        t.enabled = 1;
        t.split = batch;
        t.add[0] = add0 ? 1 : 0;
        t.add[1] = add1 ? 1 : 0;
        t.a_canon = (p_in_range && !no_tail_canon()) ? 1 : 0;
        t.a[0] = acc0;
        t.a[1] = acc1;
        t.a_bs[0] = t.a_bs[1] = acc_bs;
        t.out[0] = out0;
        t.out[1] = out1;
        t.out_bs[0] = out0_bs;
        t.out_bs[1] = out1_bs;
        t.s = e->d_moddown_pq.d;
        return lgi_ntt(Q, limb_map_identity(), nl, 2 * batch, tmp.d, tbs, tmp.d, tbs, false, 0, 0, st, false, &t);
    }
    if (ntt) LG_TRY(lgi_ntt(Q, limb_map_identity(), nl, 2 * batch, tmp.d, tbs, tmp.d, tbs, false, 0, 0, st));
    LG_TRY(lgi_ew(add0 ? EW_SUB_MULMONT_SCALAR_ADD : EW_SUB_MULMONT_SCALAR, Q, limb_map_identity(), nl, batch, acc0, acc_bs,
                  tmp.d, tbs, out0, out0_bs, e->moddown_pq.data(), nl, st));
    return lgi_ew(add1 ? EW_SUB_MULMONT_SCALAR_ADD : EW_SUB_MULMONT_SCALAR, Q, limb_map_identity(), nl, batch, acc1, acc_bs,
                  tmp.d + (size_t)batch * tbs, tbs, out1, out1_bs, e->moddown_pq.data(), nl, st);
}

extern "C" {

int lg_extender_create(const lg_ring* ringQ, const lg_ring* ringP, lg_extender** out) {
    LG_REQUIRE(ringQ && ringP && out, "NewFastBasisExtender: null argument");
    LG_REQUIRE(ringQ->N == ringP->N, "NewFastBasisExtender: ring degrees differ");
    LG_SAME_DEVICE("NewFastBasisExtender", ringQ->device, ringP->device);
    LG_ON_DEVICE(ringQ->device);
    std::unique_ptr<lg_extender> e(new lg_extender);
    e->Q = ringQ;
    e->P = ringP;
    LG_TRY(e->qp.build(ringQ->q.data(), ringQ->nl, ringP->q.data(), ringP->nl));
    LG_TRY(e->pq.build(ringP->q.data(), ringP->nl, ringQ->q.data(), ringQ->nl));
    e->moddown_pq = gen_moddown(ringQ, ringP);
    LG_TRY(e->d_moddown_pq.upload(e->moddown_pq));
    e->moddown_qp = gen_moddown(ringP, ringQ);
    *out = e.release();
    return LG_OK;
}
int lg_extender_destroy(lg_extender* e) {
    if (!e) return LG_OK;
    LG_ON_DEVICE(e->Q->device);
    delete e;
    return LG_OK;
}

int lg_extender_modup_split_qp(const lg_extender* e, int level, const lg_poly* p1, lg_poly* p2, lg_stream_t s) {
    LG_REQUIRE(e, "ModUpSplitQP: null extender");
    LG_ON_DEVICE(e->Q->device);
    LG_TRY(check_p(p1, e->Q->N, level + 1, -1, "ModUpSplitQP"));
    LG_TRY(check_p(p2, e->Q->N, e->P->nl, p1->batch, "ModUpSplitQP"));
    return lgi_modup_launch(e->qp, e->Q->N, p1->batch, p1->d, p1->bstride, level + 1, p2->d, p2->bstride, e->P->nl, 0, cs(s));
}
int lg_extender_modup_split_pq(const lg_extender* e, int level, const lg_poly* p1, lg_poly* p2, lg_stream_t s) {
    LG_REQUIRE(e, "ModUpSplitPQ: null extender");
    LG_ON_DEVICE(e->Q->device);
    LG_TRY(check_p(p1, e->Q->N, level + 1, -1, "ModUpSplitPQ"));
    LG_TRY(check_p(p2, e->Q->N, e->Q->nl, p1->batch, "ModUpSplitPQ"));
    return lgi_modup_launch(e->pq, e->Q->N, p1->batch, p1->d, p1->bstride, level + 1, p2->d, p2->bstride, e->Q->nl, 0, cs(s));
}
int lg_extender_moddown_ntt_pq(const lg_extender* e, int level, lg_poly* p1, lg_poly* p2, lg_stream_t s) {
    LG_REQUIRE(e, "ModDownNTTPQ: null extender");
    LG_ON_DEVICE(e->Q->device);
    const int nQ = e->Q->nl, nP = e->P->nl;
    LG_TRY(check_p(p1, e->Q->N, nQ + nP, -1, "ModDownNTTPQ"));
    LG_TRY(check_p(p2, e->Q->N, level + 1, p1->batch, "ModDownNTTPQ"));
    return lgi_moddown_tail_ntt(e, level, p1->batch, p1->d, p1->bstride, p1->d + (size_t)nQ * p1->N, p1->bstride, p2->d,
                                p2->bstride, true, cs(s));
}
int lg_extender_moddown_splited_ntt_pq(const lg_extender* e, int level, const lg_poly* p1Q, lg_poly* p1P, lg_poly* p2,
                                       lg_stream_t s) {
    LG_REQUIRE(e, "ModDownSplitedNTTPQ: null extender");
    LG_ON_DEVICE(e->Q->device);
    LG_TRY(check_p(p1Q, e->Q->N, level + 1, -1, "ModDownSplitedNTTPQ"));
    LG_TRY(check_p(p1P, e->Q->N, e->P->nl, p1Q->batch, "ModDownSplitedNTTPQ"));
    LG_TRY(check_p(p2, e->Q->N, level + 1, p1Q->batch, "ModDownSplitedNTTPQ"));
    return lgi_moddown_tail_ntt(e, level, p1Q->batch, p1Q->d, p1Q->bstride, p1P->d, p1P->bstride, p2->d, p2->bstride, true,
                                cs(s));
}
int lg_extender_moddown_pq(const lg_extender* e, int level, const lg_poly* p1, lg_poly* p2, lg_stream_t s) {
    LG_REQUIRE(e, "ModDownPQ: null extender");
    LG_ON_DEVICE(e->Q->device);
    const int nP = e->P->nl;
    LG_TRY(check_p(p1, e->Q->N, level + 1 + nP, -1, "ModDownPQ"));
    LG_TRY(check_p(p2, e->Q->N, level + 1, p1->batch, "ModDownPQ"));
    // :254 the P limbs follow the level+1 active Q limbs
    return lgi_moddown_tail_ntt(e, level, p1->batch, p1->d, p1->bstride, p1->d + (size_t)(level + 1) * p1->N, p1->bstride,
                                p2->d, p2->bstride, false, cs(s));
}
int lg_extender_moddown_splited_pq(const lg_extender* e, int level, const lg_poly* p1Q, const lg_poly* p1P, lg_poly* p2,
                                   lg_stream_t s) {
    LG_REQUIRE(e, "ModDownSplitedPQ: null extender");
    LG_ON_DEVICE(e->Q->device);
    LG_TRY(check_p(p1Q, e->Q->N, level + 1, -1, "ModDownSplitedPQ"));
    LG_TRY(check_p(p1P, e->Q->N, e->P->nl, p1Q->batch, "ModDownSplitedPQ"));
    LG_TRY(check_p(p2, e->Q->N, level + 1, p1Q->batch, "ModDownSplitedPQ"));
    return lgi_moddown_tail_ntt(e, level, p1Q->batch, p1Q->d, p1Q->bstride, p1P->d, p1P->bstride, p2->d, p2->bstride,
                                false, cs(s));
}
int lg_extender_moddown_splited_qp(const lg_extender* e, int levelQ, int levelP, const lg_poly* p1Q, const lg_poly* p1P,
                                   lg_poly* p2, lg_stream_t s) {
    // :314-350: polypoolP = ModUpSplitQP(levelQ, p1Q); p2 = MRed(p1P + (p - pool), Q^-1)
    LG_REQUIRE(e, "ModDownSplitedQP: null extender");
    LG_ON_DEVICE(e->Q->device);
    const lg_ring* P = e->P;
    const u64 N = P->N;
    LG_REQUIRE(levelP >= 0 && levelP < P->nl && levelQ >= 0 && levelQ < e->Q->nl, "ModDownSplitedQP: level out of range");
    LG_TRY(check_p(p1Q, N, levelQ + 1, -1, "ModDownSplitedQP"));
    LG_TRY(check_p(p1P, N, levelP + 1, p1Q->batch, "ModDownSplitedQP"));
    LG_TRY(check_p(p2, N, levelP + 1, p1Q->batch, "ModDownSplitedQP"));
    const int batch = p1Q->batch;
    Scratch tmp(cs(s));
    LG_TRY(tmp.alloc((size_t)batch * P->nl * N));
    const size_t tbs = (size_t)P->nl * N;
    LG_TRY(lgi_modup_launch(e->qp, N, batch, p1Q->d, p1Q->bstride, levelQ + 1, tmp.d, tbs, P->nl, 0, cs(s)));
    return lgi_ew(EW_SUB_MULMONT_SCALAR, P, limb_map_identity(), levelP + 1, batch, p1P->d, p1P->bstride, tmp.d, tbs, p2->d,
                  p2->bstride, e->moddown_qp.data(), levelP + 1, cs(s));
}

}  // extern "C"

// ---------------------------------------------------------------------------
// Decomposer
// ---------------------------------------------------------------------------

// Decompose / DecomposeAndSplit (:476-713).  outQ receives limbs 0..level, outP
// the nP special-prime limbs.  In the non-trivial case the reference first copies
// the digit's own limbs and then overwrites them with the converted value (the
// second target loop starts at alpha*crt, :548 / :664), so only the conversion
// is materialised.
int lgi_decompose(const lg_decomposer* d, int level, int crt, int batch, const u64* p0, size_t p0_bs, u64* outQ,
                  size_t outQ_bs, u64* outP, size_t outP_bs, cudaStream_t st, bool lazy_out) {
    LG_REQUIRE(level >= 0 && level < d->nQ, "Decompose: level %d out of range", level);
    LG_REQUIRE(crt >= 0 && crt < d->beta, "Decompose: digit %d out of range", crt);
    const int alphai = d->xalpha[crt];
    const int p0idxst = crt * d->alpha;
    const int p0idxed = p0idxst + alphai;
    LG_REQUIRE(p0idxst <= level, "Decompose: digit %d is not active at level %d", crt, level);
    const u64 N = d->N;
    if ((p0idxed > level + 1 && (level + 1) % d->nP == 1) || alphai == 1) {  // :489 / :613
        FanoutArgs f;
        f.N = (u32)N;
        f.in = p0 + (size_t)p0idxst * N;
        f.in_bs = p0_bs;
        f.nruns = 2;
        f.out[0] = outQ;
        f.out_bs[0] = outQ_bs;
        f.ndst[0] = level + 1;
        f.out[1] = outP;
        f.out_bs[1] = outP_bs;
        f.ndst[1] = d->nP;
        f.mode = 0;
        f.phalf = f.plast = 0;
        lg_launch_fanout(f, batch, st);
        LG_LAUNCH_CHECK();
        return LG_OK;
    }
    int index;  // :503-507 / :631-635
    if (level >= alphai + crt * d->alpha)
        index = d->xalpha[crt] - 2;
    else
        index = (level - 1) % d->alpha;
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
    a.out[0] = outQ;  // targets 0..level (:528-563 / :662-692)
    a.out_bs[0] = outQ_bs;
    a.ndst[0] = level + 1;
    a.tgt0[0] = 0;
    a.out[1] = outP;  // special primes live at table index nQ.. (:565-577 / :694-709)
    a.out_bs[1] = outP_bs;
    a.ndst[1] = d->nP;
    a.tgt0[1] = d->nQ;
    if (lazy_out) {
        // key-switch callers take the digit's own limbs from the NTT-domain input (ckks/evaluator.go:1579-1584): those
        // targets are never read, so they are not computed -- the Q run splits around [p0idxst, p0idxed)
        const int own1 = p0idxed < level + 1 ? p0idxed : level + 1;
        a.nruns = 3;
        a.ndst[0] = p0idxst;
        a.out[2] = a.out[1];
        a.out_bs[2] = a.out_bs[1];
        a.ndst[2] = a.ndst[1];
        a.tgt0[2] = a.tgt0[1];
        a.out[1] = outQ + (size_t)own1 * N;
        a.out_bs[1] = outQ_bs;
        a.ndst[1] = level + 1 - own1;
        a.tgt0[1] = own1;
    }
    a.copy_out = nullptr;
    a.fast = m.fast_level(a.nsrc, &a.fp_shift);
    a.lazy_out = lazy_out ? 1 : 0;
    if (lg_launch_modup(a, batch, st) != 0) {
        lg_set_error("Decompose: too many source limbs");
        return LG_ERR_ARG;
    }
    LG_LAUNCH_CHECK();
    return LG_OK;
}

static int decomposer_build(lg_decomposer* d, u64 N, const u64* Q, int nQ, const u64* P, int nP) {
    d->N = N;
    d->nQ = nQ;
    d->nP = nP;
    d->alpha = nP;
    d->beta = (nQ + nP - 1) / nP;  // :431 ceil(len(Q)/alpha)
    d->xalpha.assign(d->beta, d->alpha);
    if (nQ % d->alpha != 0) d->xalpha[d->beta - 1] = nQ % d->alpha;
    std::vector<u64> Pi(Q, Q + nQ);
    Pi.insert(Pi.end(), P, P + nP);
    d->modup.resize(d->beta);
    for (int i = 0; i < d->beta; ++i)
        for (int j = 0; j + 1 < d->xalpha[i]; ++j) {
            std::unique_ptr<ModUpDev> m(new ModUpDev);
            LG_TRY(m->build(Q + (size_t)i * d->alpha, j + 2, Pi.data(), nQ + nP));
            d->modup[i].push_back(std::move(m));
        }
    return LG_OK;
}

extern "C" {

int lg_decomposer_create(uint64_t N, const uint64_t* Q, int nQ, const uint64_t* P, int nP, lg_decomposer** out) {
    LG_REQUIRE(Q && P && out && nQ >= 1 && nP >= 1, "NewDecomposer: invalid argument");
    LG_REQUIRE(nQ + nP <= LG_MAX_LIMBS, "NewDecomposer: too many moduli");
    std::unique_ptr<lg_decomposer> d(new lg_decomposer);
    d->device = lgi_current_device();
    LG_TRY(decomposer_build(d.get(), N, Q, nQ, P, nP));
    *out = d.release();
    return LG_OK;
}
int lg_decomposer_destroy(lg_decomposer* d) {
    if (!d) return LG_OK;
    LG_ON_DEVICE(d->device);
    delete d;
    return LG_OK;
}
int lg_decomposer_beta(const lg_decomposer* d) { return d ? d->beta : 0; }
int lg_decomposer_xalpha(const lg_decomposer* d, int i) { return (d && i >= 0 && i < d->beta) ? d->xalpha[i] : 0; }

int lg_decomposer_decompose(const lg_decomposer* d, int level, int crt, const lg_poly* p0, lg_poly* p1, lg_stream_t s) {
    LG_REQUIRE(d, "Decompose: null decomposer");
    LG_ON_DEVICE(d->device);
    LG_TRY(check_p(p0, d->N, level + 1, -1, "Decompose"));
    LG_TRY(check_p(p1, d->N, level + 1 + d->nP, p0->batch, "Decompose"));
    return lgi_decompose(d, level, crt, p0->batch, p0->d, p0->bstride, p1->d, p1->bstride,
                         p1->d + (size_t)(level + 1) * d->N, p1->bstride, cs(s));
}
int lg_decomposer_decompose_and_split(const lg_decomposer* d, int level, int crt, const lg_poly* p0, lg_poly* p1Q,
                                      lg_poly* p1P, lg_stream_t s) {
    LG_REQUIRE(d, "DecomposeAndSplit: null decomposer");
    LG_ON_DEVICE(d->device);
    LG_TRY(check_p(p0, d->N, level + 1, -1, "DecomposeAndSplit"));
    LG_TRY(check_p(p1Q, d->N, level + 1, p0->batch, "DecomposeAndSplit"));
    LG_TRY(check_p(p1P, d->N, d->nP, p0->batch, "DecomposeAndSplit"));
    return lgi_decompose(d, level, crt, p0->batch, p0->d, p0->bstride, p1Q->d, p1Q->bstride, p1P->d, p1P->bstride, cs(s));
}

}  // extern "C"

// ---------------------------------------------------------------------------
// CKKS evaluator hot ops
// ---------------------------------------------------------------------------

// Digit loop shared by the CKKS and BFV key switches (ckks/evaluator.go:1511-1552,
// bfv/evaluator.go:760-806).  For each digit i < beta: Decompose(AndSplit) the coefficient-domain
// input `coef` into d = [level+1 Q limbs | nP special-prime limbs]; NTT every limb outside the digit
// (the digit's own limbs are taken from the NTT-domain copy `nttd`); acc0/acc1 += MRed(evk[i][0/1], d)
// with BRedAdd when (i & 7) == cadence and after the last digit.

int lgi_keyswitch_digits(const lg_ring* QP, const lg_ring* Q, LimbMap qp_map, const lg_decomposer* dec, int level, int beta,
                         int batch, const u64* coef, size_t coef_bs, const u64* nttd, size_t nttd_bs, const lg_swk* evk,
                         u64* d, u64* acc0, u64* acc1, size_t d_bs, int cadence, cudaStream_t st) {
    const u64 N = Q->N;
    const int nl = level + 1, nd = nl + dec->nP, alpha = dec->alpha;
    if (Q->logN >= 12) {
        // All digits are decomposed and taken through the strided NTT phase first; one kernel then finishes
        // the transform of every digit tile by tile, multiplies by the key and keeps both accumulators in
        // registers over the whole digit loop.  The sums are reduced once at the end: the reference's
        // cadence (BRedAdd when (i & 7) == cadence and after the last digit) only bounds its lazy sums and
        // ends canonical as well, so the words are the same.
        const size_t per_entry = (size_t)beta * nd * N;                 // scratch words per batch entry
        const size_t budget = (size_t)lg_switches().ks_scratch_words.load(std::memory_order_relaxed);  // 6 GiB of words by default
        int chunk = (int)(budget / per_entry);
        if (chunk < 1) chunk = 1;
        if (chunk > batch) chunk = batch;
        Scratch D(st);
        LG_TRY(D.alloc((size_t)chunk * per_entry));
        for (int b0 = 0; b0 < batch; b0 += chunk) {
            const int cb = (batch - b0) < chunk ? (batch - b0) : chunk;
            const size_t d_ds = (size_t)cb * d_bs;
            // The basis extensions of the digits are independent; for a few ciphertexts each is a fraction of a wave, so
            // they go to auxiliary streams in turn and overlap (forked from and joined to `st`).
            LgAux* aux = (cb * (N / 2 / 128) < 2 * 148 && beta > 1 && !lg_switches().no_aux_streams.load(std::memory_order_relaxed))
                             ? lg_aux_streams()
                             : nullptr;
            if (aux) lg_aux_fork(aux, st, LG_AUX_STREAMS);
            for (int i = 0; i < beta; ++i) {
                u64* Di = D.d + (size_t)i * d_ds;
                // the digits are read by the forward NTT alone: no conditional subtraction needed
                const int rc = lgi_decompose(dec, level, i, cb, coef + (size_t)b0 * coef_bs, coef_bs, Di, d_bs, Di + (size_t)nl * N,
                                             d_bs, aux ? aux->s[i % LG_AUX_STREAMS] : st, true);
                if (rc != LG_OK) {
                    if (aux) lg_aux_join(aux, st, LG_AUX_STREAMS);
                    return rc;
                }
            }
            if (aux) lg_aux_join(aux, st, LG_AUX_STREAMS);
            NttArgs a;
            memset(&a, 0, sizeof(a));
            a.T = QP->T;
            a.map = qp_map;
            a.in = D.d;
            a.out = D.d;
            a.in_bstride = a.out_bstride = d_bs;
            a.skip_alpha = alpha;  // the digit's own limbs come from the NTT-domain input
            a.skip_div = cb;
            a.skip_nl = nl;
            if (lg_launch_ntt_fwd_strided(a, nd, beta * cb, st) != 0) {
                lg_set_error("switchKeys: unsupported ring degree 2^%u", Q->logN);
                return LG_ERR_ARG;
            }
            LG_LAUNCH_CHECK();
            KsFusedArgs k;
            memset(&k, 0, sizeof(k));
            k.T = QP->T;
            k.map = qp_map;
            if (!lg_switches().no_fp_mac.load(std::memory_order_relaxed) && evk->nQP == QP->nl) {
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
            k.cx = nttd + (size_t)b0 * nttd_bs;
            k.cx_bs = nttd_bs;
            k.evk = evk->key(0, 0);
            k.evk_ds = (size_t)(evk->key(1, 0) - evk->key(0, 0));
            k.evk_hs = (size_t)(evk->key(0, 1) - evk->key(0, 0));
            k.acc0 = acc0 + (size_t)b0 * d_bs;
            k.acc1 = acc1 + (size_t)b0 * d_bs;
            k.acc_bs = d_bs;
            k.beta = beta;
            k.alpha = alpha;
            k.nl = nl;
            if (lg_launch_ks_fused(k, nd, cb, st) != 0) {
                lg_set_error("switchKeys: fused digit loop launch failed");
                return LG_ERR_ARG;
            }
            LG_LAUNCH_CHECK();
        }
        return LG_OK;
    }
    for (int i = 0; i < beta; ++i) {
        // decomposeAndSplitNTT :1561-1591 / Decompose bfv:767
        LG_TRY(lgi_decompose(dec, level, i, batch, coef, coef_bs, d, d_bs, d + (size_t)nl * N, d_bs, st));
        const int p0idxst = i * alpha;
        int p0idxed = p0idxst + dec->xalpha[i];
        if (p0idxed > nl) p0idxed = nl;
        const int first = (i == 0);
        const int reduce = ((i & 7) == cadence) || (i == beta - 1);  // ckks :1536,:1547 / bfv :795,:803
        LG_TRY(lgi_ew(EW_COPY, Q, limb_map_identity(), p0idxed - p0idxst, batch, nttd + (size_t)p0idxst * N, nttd_bs,
                      nullptr, 0, d + (size_t)p0idxst * N, d_bs, nullptr, 0, st));
        LG_TRY(lgi_ntt(QP, qp_map, nd, batch, d, d_bs, d, d_bs, false, p0idxst, p0idxed, st));
        KsMacArgs m;
        m.T = QP->T;
        m.map = qp_map;
        m.d = d;
        m.d_bs = d_bs;
        m.evk0 = evk->key(i, 0);
        m.evk1 = evk->key(i, 1);
        m.acc0 = acc0;
        m.acc1 = acc1;
        m.acc_bs = d_bs;
        m.first = first;
        m.reduce = reduce;
        lg_launch_ks_mac(m, nd, batch, st);
        LG_LAUNCH_CHECK();
    }
    return LG_OK;
}

// switchKeysInPlace, ckks/evaluator.go:1475-1558, on raw device buffers.
// cx: [batch][>=level+1][N] NTT domain.  out0/out1: level+1 limbs each; with add0/add1 the result is
// added (CRed) into what out0/out1 already hold -- the AddLvl every caller applies next (:1103-1104,
// :1158-1159, :1187, :1470).
static int ckks_switch_keys(lg_ckks_eval* e, int level, int batch, const u64* cx, size_t cx_bs, const lg_swk* evk,
                            u64* out0, size_t out0_bs, u64* out1, size_t out1_bs, cudaStream_t st, bool add0 = false,
                            bool add1 = false, bool cx_in_range = false) {
    const lg_ring* Q = e->Q;
    const lg_ring* P = e->P;
    const lg_ring* QP = e->QP.get();
    const u64 N = Q->N;
    const int nQ = Q->nl, nP = P->nl, nl = level + 1, nd = nl + nP;
    LG_REQUIRE(level >= 0 && level < nQ, "switchKeys: level %d out of range", level);
    LG_REQUIRE(evk && evk->N == N && evk->nQP == nQ + nP, "switchKeys: switching key shape mismatch");
    LG_SAME_DEVICE("switchKeys", Q->device, evk->device);
    const int alpha = e->alpha;
    const int beta = (nl + alpha - 1) / alpha;  // :1508
    LG_REQUIRE(beta <= evk->beta, "switchKeys: key has %d digits, %d needed", evk->beta, beta);

    Scratch c2(st), d(st), acc(st);
    LG_TRY(c2.alloc((size_t)batch * nl * N));
    if (Q->logN < 12) LG_TRY(d.alloc((size_t)batch * nd * N));  // larger rings: the digit loop owns its scratch
    LG_TRY(acc.alloc((size_t)2 * batch * nd * N));
    const size_t c2_bs = (size_t)nl * N, d_bs = (size_t)nd * N;
    u64* acc0 = acc.d;
    u64* acc1 = acc.d + (size_t)batch * d_bs;
    const LimbMap qp_map{nl, 0, nQ};  // Q limbs 0..level, then the special primes at #Q.. (:1519-1525)

    // :1503  c2 = InvNTT(cx)
    LG_TRY(lgi_ntt(Q, limb_map_identity(), nl, batch, cx, cx_bs, c2.d, c2_bs, true, 0, 0, st, cx_in_range));
    // :1511-1552 digit loop (decomposeAndSplitNTT + multiply-accumulate), reduce cadence reduce&7 == 1
    LG_TRY(lgi_keyswitch_digits(QP, Q, qp_map, e->dec.get(), level, beta, batch, c2.d, c2_bs, cx, cx_bs, evk, d.d, acc0,
                                acc1, d_bs, 1, st));
    // :1556-1557
    // the accumulators are canonical, so their inverse transforms need no range check
    return lgi_moddown_pair_ntt(e->ext.get(), level, batch, acc0, acc1, d_bs, nl, out0, out0_bs, add0, out1, out1_bs, add1, true,
                                st, true);
}

int lgi_swk_prepare(const lg_swk* k, const lg_ring* QP, cudaStream_t st) {
    std::lock_guard<std::mutex> lock(k->mu);
    if (k->prepared) return LG_OK;
    const size_t words = (size_t)k->beta * 2 * k->nQP * k->N;
    if (!k->d_f) LG_CUDA_CHECK(cudaMalloc((void**)&k->d_f, words * sizeof(u64)));
    if (!k->d_bad) LG_CUDA_CHECK(cudaMalloc((void**)&k->d_bad, (size_t)k->beta * 2 * k->nQP * sizeof(u32)));
    LG_CUDA_CHECK(cudaMemsetAsync(k->d_bad, 0, (size_t)k->beta * 2 * k->nQP * sizeof(u32), st));
    lg_launch_swk_prepare(QP->T, k->d, k->d_f, k->d_bad, k->beta, k->nQP, st);
    LG_LAUNCH_CHECK();
    LG_CUDA_CHECK(cudaStreamSynchronize(st));  // once per key: other streams may use it right away
    // host side: which limbs the TMA digit loop may take, and the descriptor of d_f it loads key tiles through
    std::vector<u32> bad((size_t)k->beta * 2 * k->nQP);
    LG_CUDA_CHECK(cudaMemcpy(bad.data(), k->d_bad, bad.size() * sizeof(u32), cudaMemcpyDeviceToHost));
    k->fp_ok.assign(k->nQP, 0);
    for (int tl = 0; tl < k->nQP && tl < (int)QP->q.size(); ++tl) {
        u32 b = QP->q[tl] >= (3ull << 44);
        for (int dh = 0; dh < 2 * k->beta; ++dh) b |= bad[(size_t)dh * k->nQP + tl];
        k->fp_ok[tl] = b ? 0 : 1;
    }
    k->has_map = (words % 16 == 0) && lg_encode_key_tensor_map(k->keymap, k->d_f, words) == 0;
    k->prepared = true;
    return LG_OK;
}

int lgi_concat_ring(const lg_ring* Q, const lg_ring* P, std::unique_ptr<lg_ring>& out) {
    std::unique_ptr<lg_ring> r(new lg_ring);
    r->device = Q->device;
    r->N = Q->N;
    r->logN = Q->logN;
    r->nl = Q->nl + P->nl;
    auto cat = [](const std::vector<u64>& a, const std::vector<u64>& b) {
        std::vector<u64> c(a);
        c.insert(c.end(), b.begin(), b.end());
        return c;
    };
    r->q = cat(Q->q, P->q);
    r->bred = cat(Q->bred, P->bred);
    r->mred = cat(Q->mred, P->mred);
    r->ninv = cat(Q->ninv, P->ninv);
    r->psi = cat(Q->psi, P->psi);
    r->psi_inv = cat(Q->psi_inv, P->psi_inv);
    LG_TRY(lgi_ring_build_device(r.get()));
    out = std::move(r);
    return LG_OK;
}

extern "C" {

int lg_ckks_eval_create(const lg_ring* ringQ, const lg_ring* ringP, lg_ckks_eval** out) {
    LG_REQUIRE(ringQ && ringP && out, "NewEvaluator: null argument");
    LG_REQUIRE(ringQ->N == ringP->N, "NewEvaluator: ring degrees differ");
    LG_REQUIRE(ringQ->nl + ringP->nl <= LG_MAX_LIMBS, "NewEvaluator: too many moduli");
    LG_SAME_DEVICE("NewEvaluator", ringQ->device, ringP->device);
    LG_ON_DEVICE(ringQ->device);
    std::unique_ptr<lg_ckks_eval> e(new lg_ckks_eval);
    e->Q = ringQ;
    e->P = ringP;
    e->alpha = ringP->nl;
    LG_TRY(lgi_concat_ring(ringQ, ringP, e->QP));
    lg_extender* ext = nullptr;
    LG_TRY(lg_extender_create(ringQ, ringP, &ext));
    e->ext.reset(ext);
    lg_decomposer* dec = nullptr;
    LG_TRY(lg_decomposer_create(ringQ->N, ringQ->q.data(), ringQ->nl, ringP->q.data(), ringP->nl, &dec));
    e->dec.reset(dec);
    *out = e.release();
    return LG_OK;
}
int lg_ckks_eval_destroy(lg_ckks_eval* e) {
    if (!e) return LG_OK;
    LG_ON_DEVICE(e->Q->device);
    delete e;
    return LG_OK;
}

int lg_swk_create(uint64_t N, int beta, int nQP, const uint64_t* host, lg_swk** out) {
    LG_REQUIRE(host && out && beta >= 1 && nQP >= 1, "SwitchingKey: invalid argument");
    std::unique_ptr<lg_swk> k(new lg_swk);
    k->device = lgi_current_device();
    k->N = N;
    k->beta = beta;
    k->nQP = nQP;
    k->owns = true;
    const size_t bytes = (size_t)beta * 2 * nQP * N * sizeof(u64);
    LG_CUDA_CHECK(cudaMalloc((void**)&k->d, bytes));
    LG_CUDA_CHECK(cudaMemcpy(k->d, host, bytes, cudaMemcpyHostToDevice));
    *out = k.release();
    return LG_OK;
}
int lg_swk_wrap(void* device_ptr, uint64_t N, int beta, int nQP, lg_swk** out) {
    LG_REQUIRE(device_ptr && out && beta >= 1 && nQP >= 1, "SwitchingKey: invalid argument");
    LG_REQUIRE(((uintptr_t)device_ptr & 31) == 0, "device pointer must be 32-byte aligned");
    lg_swk* k = new lg_swk;
    k->device = lgi_pointer_device(device_ptr);
    k->d = (u64*)device_ptr;
    k->N = N;
    k->beta = beta;
    k->nQP = nQP;
    k->owns = false;
    *out = k;
    return LG_OK;
}
// an uninitialised key of the given shape (filled through lg_swk_poly views, e.g. by lg_poly_decode)
int lg_swk_alloc(uint64_t N, int beta, int nQP, lg_swk** out) {
    LG_REQUIRE(out && beta >= 1 && nQP >= 1 && N >= 1, "SwitchingKey: invalid argument");
    std::unique_ptr<lg_swk> k(new lg_swk);
    k->device = lgi_current_device();
    k->N = N;
    k->beta = beta;
    k->nQP = nQP;
    k->owns = true;
    const size_t bytes = (size_t)beta * 2 * nQP * N * sizeof(u64);
    LG_CUDA_CHECK(cudaMalloc((void**)&k->d, bytes));
    LG_CUDA_CHECK(cudaMemset(k->d, 0, bytes));
    *out = k.release();
    return LG_OK;
}
// evakey[digit][half] as a non-owning polynomial handle over QP (the key must outlive it)
int lg_swk_invalidate(lg_swk* k) {
    LG_REQUIRE(k, "SwitchingKey: null argument");
    std::lock_guard<std::mutex> lock(k->mu);
    k->prepared = false;
    return LG_OK;
}
int lg_swk_poly(const lg_swk* k, int digit, int half, lg_poly** out) {
    LG_REQUIRE(k && out, "SwitchingKey: null argument");
    {
        std::lock_guard<std::mutex> lock(k->mu);
        k->prepared = false;  // a view exists to fill the key: its derived forms are rebuilt at the next use
    }
    LG_REQUIRE(digit >= 0 && digit < k->beta && (half == 0 || half == 1), "SwitchingKey: evakey[%d][%d] out of range", digit, half);
    LG_ON_DEVICE(k->device);
    return lg_poly_wrap((void*)k->key(digit, half), k->N, k->nQP, 1, out);
}
int lg_swk_beta(const lg_swk* k) { return k ? k->beta : 0; }
int lg_swk_nlimbs(const lg_swk* k) { return k ? k->nQP : 0; }
uint64_t lg_swk_n(const lg_swk* k) { return k ? k->N : 0; }
int lg_swk_destroy(lg_swk* k) {
    if (!k) return LG_OK;
    LG_ON_DEVICE(k->device);
    if (k->owns && k->d) cudaFree(k->d);
    delete k;
    return LG_OK;
}

int lg_ckks_switch_keys_in_place(lg_ckks_eval* e, int level, const lg_poly* cx, const lg_swk* evk, lg_poly* p0, lg_poly* p1,
                                 lg_stream_t s) {
    LG_REQUIRE(e, "switchKeysInPlace: null evaluator");
    LG_ON_DEVICE(e->Q->device);
    const u64 N = e->Q->N;
    LG_TRY(check_p(cx, N, level + 1, -1, "switchKeysInPlace"));
    LG_TRY(check_p(p0, N, level + 1, cx->batch, "switchKeysInPlace"));
    LG_TRY(check_p(p1, N, level + 1, cx->batch, "switchKeysInPlace"));
    return ckks_switch_keys(e, level, cx->batch, cx->d, cx->bstride, evk, p0->d, p0->bstride, p1->d, p1->bstride, cs(s));
}

int lg_ckks_mul_relin(lg_ckks_eval* e, int level, const lg_poly* a0, const lg_poly* a1, const lg_poly* b0, const lg_poly* b1,
                      const lg_swk* rlk, lg_poly* out0, lg_poly* out1, lg_stream_t s) {
    LG_REQUIRE(e, "MulRelin: null evaluator");
    LG_ON_DEVICE(e->Q->device);
    const lg_ring* Q = e->Q;
    const u64 N = Q->N;
    const int nl = level + 1;
    LG_REQUIRE(level >= 0 && nl <= Q->nl, "MulRelin: level %d out of range", level);
    LG_TRY(check_p(a0, N, nl, -1, "MulRelin"));
    const int batch = a0->batch;
    LG_TRY(check_p(a1, N, nl, batch, "MulRelin"));
    LG_TRY(check_p(b0, N, nl, batch, "MulRelin"));
    LG_TRY(check_p(b1, N, nl, batch, "MulRelin"));
    LG_TRY(check_p(out0, N, nl, batch, "MulRelin"));
    LG_TRY(check_p(out1, N, nl, batch, "MulRelin"));
    cudaStream_t st = cs(s);
    const bool square = (a0->d == b0->d && a1->d == b1->d);  // el0 == el1, :1080
    const size_t bs = (size_t)nl * N;
    Scratch w(st);
    LG_TRY(w.alloc((size_t)batch * bs));
    u64* c2 = w.d;
    // :1076-1095 tensor product in one pass; c0, c1 land directly in the outputs (same-index aliasing
    // with the inputs is safe), c2 in scratch
    TensorArgs t;
    t.T = Q->T;
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
    t.c_bs[2] = bs;
    t.square = square ? 1 : 0;
    t.nomod = 0;
    t.limb0 = 0;
    lg_launch_tensor(t, nl, batch, st);
    LG_LAUNCH_CHECK();
    // :1098-1104 relinearise c2 and add: out0 = CRed(c0 + pool1), out1 = CRed(c1 + pool2)
    // (c2 is canonical: MRed output of the tensor kernel)
    return ckks_switch_keys(e, level, batch, c2, bs, rlk, out0->d, out0->bstride, out1->d, out1->bstride, st, true, true, true);
}

int lg_ckks_relinearize(lg_ckks_eval* e, int level, const lg_poly* c0, const lg_poly* c1, const lg_poly* c2, const lg_swk* rlk,
                        lg_poly* out0, lg_poly* out1, lg_stream_t s) {
    LG_REQUIRE(e, "Relinearize: null evaluator");
    LG_ON_DEVICE(e->Q->device);
    const lg_ring* Q = e->Q;
    const u64 N = Q->N;
    const int nl = level + 1;
    LG_REQUIRE(level >= 0 && nl <= Q->nl, "Relinearize: level %d out of range", level);
    LG_TRY(check_p(c0, N, nl, -1, "Relinearize"));
    const int batch = c0->batch;
    LG_TRY(check_p(c1, N, nl, batch, "Relinearize"));
    LG_TRY(check_p(c2, N, nl, batch, "Relinearize"));
    LG_TRY(check_p(out0, N, nl, batch, "Relinearize"));
    LG_TRY(check_p(out1, N, nl, batch, "Relinearize"));
    cudaStream_t st = cs(s);
    const LimbMap id = limb_map_identity();
    LG_REQUIRE(c2->d != out0->d && c2->d != out1->d, "Relinearize: value[2] must not alias the receiver");
    if (c0->d != out0->d) LG_TRY(lgi_ew(EW_COPY, Q, id, nl, batch, c0->d, c0->bstride, nullptr, 0, out0->d, out0->bstride, nullptr, 0, st));
    if (c1->d != out1->d) LG_TRY(lgi_ew(EW_COPY, Q, id, nl, batch, c1->d, c1->bstride, nullptr, 0, out1->d, out1->bstride, nullptr, 0, st));
    // :1156-1159
    return ckks_switch_keys(e, level, batch, c2->d, c2->bstride, rlk, out0->d, out0->bstride, out1->d, out1->bstride, st, true,
                            true);
}

int lg_ckks_rescale(lg_ckks_eval* e, int nl, lg_poly* c0, lg_poly* c1, int nb, lg_stream_t s) {
    LG_REQUIRE(e, "Rescale: null evaluator");
    LG_ON_DEVICE(e->Q->device);
    LG_TRY(check_p(c0, e->Q->N, nl, -1, "Rescale"));
    LG_TRY(check_p(c1, e->Q->N, nl, c0->batch, "Rescale"));
    LG_REQUIRE(nb >= 1 && nb < nl, "cannot Rescale: input Ciphertext already at level 0");  // ckks/evaluator.go:938
    for (int k = 0; k < nb; ++k) {  // :955-960
        LG_TRY(lgi_div_by_last_modulus(e->Q, nl - k, c0->batch, c0->d, c0->bstride, true, true, cs(s)));
        LG_TRY(lgi_div_by_last_modulus(e->Q, nl - k, c1->batch, c1->d, c1->bstride, true, true, cs(s)));
    }
    return LG_OK;
}

int lg_ckks_switch_keys(lg_ckks_eval* e, int level, const lg_poly* c0, const lg_poly* c1, const lg_swk* k, lg_poly* out0,
                        lg_poly* out1, lg_stream_t s) {
    LG_REQUIRE(e, "SwitchKeys: null evaluator");
    LG_ON_DEVICE(e->Q->device);
    const lg_ring* Q = e->Q;
    const u64 N = Q->N;
    const int nl = level + 1;
    LG_REQUIRE(level >= 0 && nl <= Q->nl, "SwitchKeys: level %d out of range", level);
    LG_TRY(check_p(c0, N, nl, -1, "SwitchKeys"));
    const int batch = c0->batch;
    LG_TRY(check_p(c1, N, nl, batch, "SwitchKeys"));
    LG_TRY(check_p(out0, N, nl, batch, "SwitchKeys"));
    LG_TRY(check_p(out1, N, nl, batch, "SwitchKeys"));
    cudaStream_t st = cs(s);
    const LimbMap id = limb_map_identity();
    LG_REQUIRE(out0->d != c1->d, "SwitchKeys: receiver value[0] must not alias input value[1]");
    if (c0->d != out0->d) LG_TRY(lgi_ew(EW_COPY, Q, id, nl, batch, c0->d, c0->bstride, nullptr, 0, out0->d, out0->bstride, nullptr, 0, st));
    // :1184-1188: value[0] += pool1, value[1] = pool2
    return ckks_switch_keys(e, level, batch, c1->d, c1->bstride, k, out0->d, out0->bstride, out1->d, out1->bstride, st, true,
                            false);
}

int lg_ckks_hoist(lg_ckks_eval* e, int level, const lg_poly* c1, lg_hoisted** out, lg_stream_t s) {
    LG_REQUIRE(e && out, "RotateHoisted: null argument");
    LG_ON_DEVICE(e->Q->device);
    const lg_ring* Q = e->Q;
    const lg_ring* QP = e->QP.get();
    const u64 N = Q->N;
    const int nQ = Q->nl, nP = e->P->nl, nl = level + 1, nd = nl + nP, alpha = e->alpha;
    LG_REQUIRE(level >= 0 && nl <= nQ, "RotateHoisted: level %d out of range", level);
    LG_REQUIRE(Q->logN >= 12, "RotateHoisted: ring degree below 2^12 is not supported by the hoisted path");
    LG_TRY(check_p(c1, N, nl, -1, "RotateHoisted"));
    const int batch = c1->batch;
    const int beta = (nl + alpha - 1) / alpha;  // :1259
    cudaStream_t st = cs(s);
    std::unique_ptr<lg_hoisted> h(new lg_hoisted);
    h->device = Q->device;
    h->N = N;
    h->level = level;
    h->beta = beta;
    h->batch = batch;
    h->nd = nd;
    h->d_bs = (size_t)nd * N;
    h->d_ds = (size_t)batch * h->d_bs;
    // stream-ordered like every other scratch: a synchronous cudaMalloc / cudaFree of the digit array (C4, batch 16:
    // 2.7 GB) costs more than the eight rotations it serves
    h->st = st;
    LG_CUDA_CHECK(cudaMallocAsync((void**)&h->d, (size_t)beta * h->d_ds * sizeof(u64), st));
    auto fail = [&](int rc) {
        cudaFreeAsync(h->d, st);
        h->d = nullptr;
        return rc;
    };
    Scratch c2(st);
    int rc = c2.alloc((size_t)batch * nl * N);
    if (rc != LG_OK) return fail(rc);
    const size_t c2_bs = (size_t)nl * N;
    // :1256  c2InvNTT = InvNTT(value[1])
    rc = lgi_ntt(Q, limb_map_identity(), nl, batch, c1->d, c1->bstride, c2.d, c2_bs, true, 0, 0, st);
    if (rc != LG_OK) return fail(rc);
    // :1267-1272 decomposeAndSplitNTT of every digit: decompose, forward NTT of the limbs outside the digit
    // (one digit-batched launch pair), the digit's own limbs copied from the NTT-domain input (:1579-1584)
    for (int i = 0; i < beta; ++i) {
        u64* Di = h->d + (size_t)i * h->d_ds;
        rc = lgi_decompose(e->dec.get(), level, i, batch, c2.d, c2_bs, Di, h->d_bs, Di + (size_t)nl * N, h->d_bs, st, true);
        if (rc != LG_OK) return fail(rc);
        const int p0 = i * alpha, p1 = (p0 + alpha < nl) ? p0 + alpha : nl;
        rc = lgi_ew(EW_COPY, Q, limb_map_identity(), p1 - p0, batch, c1->d + (size_t)p0 * N, c1->bstride, nullptr, 0,
                    Di + (size_t)p0 * N, h->d_bs, nullptr, 0, st);
        if (rc != LG_OK) return fail(rc);
    }
    NttArgs a;
    memset(&a, 0, sizeof(a));
    a.T = QP->T;
    a.map = LimbMap{nl, 0, nQ};
    a.in = h->d;
    a.out = h->d;
    a.in_bstride = a.out_bstride = h->d_bs;
    a.skip_alpha = alpha;
    a.skip_div = batch;
    a.skip_nl = nl;
    if (lg_launch_ntt(a, nd, beta * batch, false, st) != 0) {
        lg_set_error("RotateHoisted: unsupported ring degree 2^%u", Q->logN);
        return fail(LG_ERR_ARG);
    }
    if (cudaPeekAtLastError() != cudaSuccess) {
        lg_set_error("RotateHoisted: kernel launch: %s", cudaGetErrorString(cudaGetLastError()));
        return fail(LG_ERR_CUDA);
    }
    *out = h.release();
    return LG_OK;
}

int lg_hoisted_destroy(lg_hoisted* h) {
    if (!h) return LG_OK;
    LG_ON_DEVICE(h->device);
    if (h->d) cudaFreeAsync(h->d, h->st);  // after the rotations issued on the hoist's stream
    delete h;
    return LG_OK;
}

int lg_ckks_switch_key_hoisted(lg_ckks_eval* e, const lg_hoisted* h, const lg_poly* c0, const lg_galois* g, const lg_swk* k,
                               lg_poly* out0, lg_poly* out1, lg_stream_t s) {
    LG_REQUIRE(e && h && g && k, "switchKeyHoisted: null argument");
    LG_ON_DEVICE(e->Q->device);
    LG_SAME_DEVICE("switchKeyHoisted", e->Q->device, h->device);
    LG_SAME_DEVICE("switchKeyHoisted", e->Q->device, g->device);
    LG_SAME_DEVICE("switchKeyHoisted", e->Q->device, k->device);
    const lg_ring* Q = e->Q;
    const lg_ring* QP = e->QP.get();
    const u64 N = Q->N;
    const int nQ = Q->nl, nP = e->P->nl, level = h->level, nl = level + 1, nd = nl + nP;
    LG_REQUIRE(h->N == N && h->nd == nd, "switchKeyHoisted: decomposition does not match the evaluator");
    LG_REQUIRE(g->N == N, "switchKeyHoisted: index length mismatch");
    LG_REQUIRE(k->N == N && k->nQP == nQ + nP, "switchKeyHoisted: switching key shape mismatch");
    LG_REQUIRE(h->beta <= k->beta, "switchKeyHoisted: key has %d digits, %d needed", k->beta, h->beta);
    const int batch = h->batch;
    LG_TRY(check_p(c0, N, nl, batch, "switchKeyHoisted"));
    LG_TRY(check_p(out0, N, nl, batch, "switchKeyHoisted"));
    LG_TRY(check_p(out1, N, nl, batch, "switchKeyHoisted"));
    LG_REQUIRE(c0->d != out0->d, "switchKeyHoisted: PermuteNTTWithIndex is not in place (ctOut must differ from ct0)");
    cudaStream_t st = cs(s);
    // :1318-1320  ctOut.value[0] = Permute(ct0.value[0])
    PermArgs pa;
    memset(&pa.T, 0, sizeof(pa.T));
    pa.T.N = (u32)N;
    pa.T.logN = Q->logN;
    pa.map = limb_map_identity();
    pa.index = g->d_index.d;
    pa.gen = 0;
    pa.in = c0->d;
    pa.in_bs = c0->bstride;
    pa.out = out0->d;
    pa.out_bs = out0->bstride;
    lg_launch_permute_ntt(pa, nl, batch, st);
    LG_LAUNCH_CHECK();
    // :1336-1378 the digit loop on permuted digits
    Scratch acc(st);
    LG_TRY(acc.alloc((size_t)2 * batch * h->d_bs));
    u64* acc0 = acc.d;
    u64* acc1 = acc.d + (size_t)batch * h->d_bs;
    KsHoistArgs ka;
    memset(&ka, 0, sizeof(ka));
    ka.T = QP->T;
    ka.map = LimbMap{nl, 0, nQ};
    ka.D = h->d;
    ka.d_ds = h->d_ds;
    ka.d_bs = h->d_bs;
    ka.index = g->d_index.d;
    ka.evk = k->key(0, 0);
    ka.evk_ds = (size_t)(k->key(1, 0) - k->key(0, 0));
    ka.evk_hs = (size_t)(k->key(0, 1) - k->key(0, 0));
    if (!lg_switches().no_fp_mac.load(std::memory_order_relaxed) && k->nQP == QP->nl) {
        LG_TRY(lgi_swk_prepare(k, QP, st));
        ka.evk_f = k->d_f;
        ka.key_bad = k->d_bad;
        ka.nqp = k->nQP;
    }
    ka.acc0 = acc0;
    ka.acc1 = acc1;
    ka.acc_bs = h->d_bs;
    ka.beta = h->beta;
    LG_REQUIRE(lg_launch_ks_hoisted(ka, nd, batch, st) == 0, "switchKeyHoisted: launch failed");
    LG_LAUNCH_CHECK();
    // :1382-1386  ModDown both; value[0] += pool2Q, value[1] = pool3Q
    return lgi_moddown_pair_ntt(e->ext.get(), level, batch, acc0, acc1, h->d_bs, nl, out0->d, out0->bstride, true, out1->d,
                                out1->bstride, false, true, st, true);
}

int lg_ckks_permute_ntt(lg_ckks_eval* e, int level, const lg_poly* c0, const lg_poly* c1, const lg_galois* g, const lg_swk* k,
                        lg_poly* out0, lg_poly* out1, lg_stream_t s) {
    LG_REQUIRE(e && g, "permuteNTT: null argument");
    LG_ON_DEVICE(e->Q->device);
    LG_SAME_DEVICE("permuteNTT", e->Q->device, g->device);
    const lg_ring* Q = e->Q;
    const u64 N = Q->N;
    const int nl = level + 1;
    LG_REQUIRE(level >= 0 && nl <= Q->nl, "permuteNTT: level %d out of range", level);
    LG_REQUIRE(g->N == N, "permuteNTT: index length mismatch");
    LG_TRY(check_p(c0, N, nl, -1, "permuteNTT"));
    const int batch = c0->batch;
    LG_TRY(check_p(c1, N, nl, batch, "permuteNTT"));
    LG_TRY(check_p(out0, N, nl, batch, "permuteNTT"));
    LG_TRY(check_p(out1, N, nl, batch, "permuteNTT"));
    cudaStream_t st = cs(s);
    const size_t bs = (size_t)nl * N;
    Scratch w(st);
    LG_TRY(w.alloc((size_t)2 * batch * bs));
    u64* el0 = w.d;
    u64* el1 = el0 + batch * bs;
    PermArgs a;
    memset(&a.T, 0, sizeof(a.T));
    a.T.N = (u32)N;
    a.T.logN = Q->logN;
    a.map = limb_map_identity();
    a.index = g->d_index.d;
    a.gen = 0;
    a.in = c0->d;  // :1462-1463
    a.in_bs = c0->bstride;
    a.out = el0;
    a.out_bs = bs;
    lg_launch_permute_ntt(a, nl, batch, st);
    a.in = c1->d;
    a.in_bs = c1->bstride;
    a.out = el1;
    lg_launch_permute_ntt(a, nl, batch, st);
    LG_LAUNCH_CHECK();
    const LimbMap id = limb_map_identity();
    // :1468-1471: value[0] = el0 + pool1, value[1] = pool2
    LG_TRY(lgi_ew(EW_COPY, Q, id, nl, batch, el0, bs, nullptr, 0, out0->d, out0->bstride, nullptr, 0, st));
    return ckks_switch_keys(e, level, batch, el1, bs, k, out0->d, out0->bstride, out1->d, out1->bstride, st, true, false);
}

}  // extern "C"
