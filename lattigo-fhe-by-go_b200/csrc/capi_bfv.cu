// capi_bfv.cu -- C-ABI host layer, part 3: hot ops of the BFV evaluator
// (bfv/evaluator.go:278-813): tensorAndRescale (Mul), switchKeys, relinearize,
// SwitchKeys and permute (RotateColumns / RotateRows with a direct key).
// BFV ciphertexts live in the coefficient domain; the key switch shares the
// fused digit loop (decompose -> NTT -> multiply-accumulate) with CKKS.
#include <string.h>

#include "capi_internal.hpp"

static inline cudaStream_t cs(lg_stream_t s) { return (cudaStream_t)s; }

struct lg_bfv_eval {
    const lg_ring* Q = nullptr;     // contextQ
    const lg_ring* M = nullptr;     // contextQMul
    const lg_ring* P = nullptr;     // contextP
    std::unique_ptr<lg_ring> QP;    // contextQP  (bfv/bfv.go:63)
    std::unique_ptr<lg_ring> QM;    // Q || QMul tables: the tensor runs over both bases in one launch
    std::unique_ptr<lg_extender> q1q2;  // baseconverterQ1Q2 (bfv/evaluator.go:95)
    std::unique_ptr<lg_extender> q1p;   // baseconverterQ1P  (:86)
    std::unique_ptr<lg_decomposer> dec;
    int alpha = 0, beta = 0;
    u64 t = 0;
    std::vector<u64> phalf_m, phalf_q;  // pHalf = QMul >> 1 (:98) mod each prime of QMul / Q
};

namespace {

// minimal multi-word unsigned integer for pHalf = (prod QMul) >> 1
struct Big {
    std::vector<u64> w{1};
    void mul(u64 m) {
        u64 carry = 0;
        for (auto& x : w) {
            const unsigned __int128 p = (unsigned __int128)x * m + carry;
            x = (u64)p;
            carry = (u64)(p >> 64);
        }
        if (carry) w.push_back(carry);
    }
    void shr1() {
        for (size_t i = 0; i < w.size(); ++i) w[i] = (w[i] >> 1) | (i + 1 < w.size() ? w[i + 1] << 63 : 0);
    }
    u64 mod(u64 m) const {
        unsigned __int128 r = 0;
        for (size_t i = w.size(); i-- > 0;) r = ((r << 64) | w[i]) % m;
        return (u64)r;
    }
};

int check_p(const lg_poly* p, u64 N, int nl, int batch, const char* what) {
    LG_REQUIRE(p, "%s: null polynomial", what);
    LG_REQUIRE(p->N == N, "%s: degree mismatch", what);
    LG_REQUIRE(p->nlimbs >= nl, "%s: polynomial has %d limbs, %d needed", what, p->nlimbs, nl);
    LG_REQUIRE(batch < 0 || p->batch == batch, "%s: batch mismatch", what);
    LG_SAME_DEVICE(what, lgi_expected_device(), p->device);
    return LG_OK;
}

// switchKeys, bfv/evaluator.go:736-813.  cx: coefficient domain, nQ limbs.  The reference leaves
// the result in the first nQ limbs of two QP polys; here out0/out1 receive those nQ limbs (added
// with CRed when add0/add1, the context.Add every caller applies next).
int bfv_switch_keys(lg_bfv_eval* e, int batch, const u64* cx, size_t cx_bs, const lg_swk* evk, u64* out0, size_t out0_bs,
                    bool add0, u64* out1, size_t out1_bs, bool add1, cudaStream_t st) {
    const lg_ring* Q = e->Q;
    const lg_ring* QP = e->QP.get();
    const u64 N = Q->N;
    const int nQ = Q->nl, nP = e->P->nl, nd = nQ + nP, level = nQ - 1;
    LG_REQUIRE(evk && evk->N == N && evk->nQP == nd, "switchKeys: switching key shape mismatch");
    LG_SAME_DEVICE("switchKeys", Q->device, evk->device);
    LG_REQUIRE(e->beta <= evk->beta, "switchKeys: key has %d digits, %d needed", evk->beta, e->beta);
    Scratch c2(st), d(st), acc(st);
    LG_TRY(c2.alloc((size_t)batch * nQ * N));
    if (Q->logN < 12) LG_TRY(d.alloc((size_t)batch * nd * N));  // larger rings: the digit loop owns its scratch
    LG_TRY(acc.alloc((size_t)2 * batch * nd * N));
    const size_t c2_bs = (size_t)nQ * N, d_bs = (size_t)nd * N;
    u64* acc0 = acc.d;
    u64* acc1 = acc.d + (size_t)batch * d_bs;
    // :753  c2 = NTT(cx)
    LG_TRY(lgi_ntt(Q, limb_map_identity(), nQ, batch, cx, cx_bs, c2.d, c2_bs, false, 0, 0, st));
    // :760-806 digit loop; all QP limbs are active, reduce cadence reduce&7 == 7
    LG_TRY(lgi_keyswitch_digits(QP, Q, limb_map_identity(), e->dec.get(), level, e->beta, batch, cx, cx_bs, c2.d, c2_bs, evk,
                                d.d, acc0, acc1, d_bs, 7, st));
    // :808-809 InvNTT over QP of both accumulators (contiguous: one launch over 2*batch entries)
    // (the accumulators are canonical, so the inverse needs no range check)
    LG_TRY(lgi_ntt(QP, limb_map_identity(), nd, 2 * batch, acc0, d_bs, acc0, d_bs, true, 0, 0, st, true));
    // :811-812 ModDownPQ
    LG_TRY(lgi_moddown_pair_ntt(e->q1p.get(), level, batch, acc0, acc1, d_bs, nQ, out0, out0_bs, add0, out1, out1_bs, add1, false,
                                st, true));
    return LG_OK;
}

int copy_if_needed(const lg_ring* Q, int batch, const lg_poly* src, lg_poly* dst, cudaStream_t st) {
    if (src->d == dst->d) return LG_OK;
    return lgi_ew(EW_COPY, Q, limb_map_identity(), Q->nl, batch, src->d, src->bstride, nullptr, 0, dst->d, dst->bstride, nullptr,
                  0, st);
}

}  // namespace

extern "C" {

int lg_bfv_eval_create(const lg_ring* ringQ, const lg_ring* ringQMul, const lg_ring* ringP, uint64_t t, lg_bfv_eval** out) {
    LG_REQUIRE(ringQ && ringQMul && ringP && out, "NewEvaluator: null argument");
    LG_REQUIRE(ringQ->N == ringQMul->N && ringQ->N == ringP->N, "NewEvaluator: ring degrees differ");
    LG_REQUIRE(ringQ->nl + ringQMul->nl <= LG_MAX_LIMBS && ringQ->nl + ringP->nl <= LG_MAX_LIMBS, "NewEvaluator: too many moduli");
    LG_SAME_DEVICE("NewEvaluator", ringQ->device, ringQMul->device);
    LG_SAME_DEVICE("NewEvaluator", ringQ->device, ringP->device);
    LG_ON_DEVICE(ringQ->device);
    std::unique_ptr<lg_bfv_eval> e(new lg_bfv_eval);
    e->Q = ringQ;
    e->M = ringQMul;
    e->P = ringP;
    e->t = t;
    e->alpha = ringP->nl;
    e->beta = (ringQ->nl + ringP->nl - 1) / ringP->nl;  // bfv/params.go: ceil(len(Qi)/alpha)
    LG_TRY(lgi_concat_ring(ringQ, ringP, e->QP));
    LG_TRY(lgi_concat_ring(ringQ, ringQMul, e->QM));
    lg_extender* x = nullptr;
    LG_TRY(lg_extender_create(ringQ, ringQMul, &x));
    e->q1q2.reset(x);
    LG_TRY(lg_extender_create(ringQ, ringP, &x));
    e->q1p.reset(x);
    lg_decomposer* d = nullptr;
    LG_TRY(lg_decomposer_create(ringQ->N, ringQ->q.data(), ringQ->nl, ringP->q.data(), ringP->nl, &d));
    e->dec.reset(d);
    Big ph;  // :98 pHalf = QMul.ModulusBigint >> 1
    for (u64 q : ringQMul->q) ph.mul(q);
    ph.shr1();
    for (u64 q : ringQMul->q) e->phalf_m.push_back(ph.mod(q));
    for (u64 q : ringQ->q) e->phalf_q.push_back(ph.mod(q));
    *out = e.release();
    return LG_OK;
}
int lg_bfv_eval_destroy(lg_bfv_eval* e) {
    if (!e) return LG_OK;
    LG_ON_DEVICE(e->Q->device);
    delete e;
    return LG_OK;
}

// tensorAndRescale (bfv/evaluator.go:278-464) for two degree-1 ciphertexts; identical handles for
// (a0,a1) and (b0,b1) select the squaring branch (:334-349).  out0..out2 = the degree-2 result.
int lg_bfv_mul(lg_bfv_eval* e, const lg_poly* a0, const lg_poly* a1, const lg_poly* b0, const lg_poly* b1, lg_poly* out0,
               lg_poly* out1, lg_poly* out2, lg_stream_t s) {
    LG_REQUIRE(e, "Mul: null evaluator");
    LG_ON_DEVICE(e->Q->device);
    const lg_ring* Q = e->Q;
    const lg_ring* M = e->M;
    const lg_ring* QM = e->QM.get();
    const u64 N = Q->N;
    const int nQ = Q->nl, nM = M->nl, nT = nQ + nM;
    LG_TRY(check_p(a0, N, nQ, -1, "Mul"));
    const int B = a0->batch;
    const lg_poly* in[4] = {a0, a1, b0, b1};
    lg_poly* outs[3] = {out0, out1, out2};
    for (int i = 1; i < 4; ++i) LG_TRY(check_p(in[i], N, nQ, B, "Mul"));
    for (int i = 0; i < 3; ++i) LG_TRY(check_p(outs[i], N, nQ, B, "Mul"));
    cudaStream_t st = cs(s);
    const bool square = (a0->d == b0->d && a1->d == b1->d);
    const int nin = square ? 2 : 4;
    const size_t ws = (size_t)nT * N;  // one extended poly
    Scratch W(st), T(st), tmp(st);
    LG_TRY(W.alloc((size_t)nin * B * ws));  // inputs in Q||QMul, NTT domain: [nin][B][nT][N]
    LG_TRY(T.alloc((size_t)3 * B * ws));    // tensor outputs:              [3][B][nT][N]
    const LimbMap id = limb_map_identity();
    const LimbMap qmul_map{1 << 30, nQ, 0};  // data limb j -> table limb nQ + j of the Q||QMul tables
    for (int p = 0; p < nin; ++p) {
        u64* w = W.d + (size_t)p * B * ws;
        // :299 / :308 ModUpSplitQP(levelQ, ct.value[i], cQ2[i])
        LG_TRY(lgi_modup_launch(e->q1q2->qp, N, B, in[p]->d, in[p]->bstride, nQ, w + (size_t)nQ * N, ws, nM, 0, st));
        // :301 / :310 contextQ.NTT(ct.value[i], cQ1[i])
        LG_TRY(lgi_ntt(Q, id, nQ, B, in[p]->d, in[p]->bstride, w, ws, false, 0, 0, st));
    }
    // :302 / :311 contextQMul.NTT(cQ2[i], cQ2[i]) for every input at once
    LG_TRY(lgi_ntt(QM, qmul_map, nM, nin * B, W.d + (size_t)nQ * N, ws, W.d + (size_t)nQ * N, ws, false, 0, 0, st));
    // :327-367 MForm + tensor in both bases, one pass
    TensorArgs ta;
    ta.T = QM->T;
    const u64* w0 = W.d;
    const u64* w1 = W.d + (size_t)B * ws;
    ta.a0 = w0;
    ta.a1 = w1;
    ta.b0 = square ? w0 : W.d + (size_t)2 * B * ws;
    ta.b1 = square ? w1 : W.d + (size_t)3 * B * ws;
    ta.c0 = T.d;
    ta.c1 = T.d + (size_t)B * ws;
    ta.c2 = T.d + (size_t)2 * B * ws;
    for (int i = 0; i < 2; ++i) ta.a_bs[i] = ta.b_bs[i] = ws;
    for (int i = 0; i < 3; ++i) ta.c_bs[i] = ws;
    ta.square = square ? 1 : 0;
    ta.nomod = 1;
    ta.limb0 = 0;
    lg_launch_tensor(ta, nT, B, st);
    LG_LAUNCH_CHECK();
    // :424-425 InvNTT of the three outputs in both bases
    LG_TRY(lgi_ntt(QM, id, nT, 3 * B, T.d, ws, T.d, ws, true, 0, 0, st));
    // :450 ModDownSplitedQP(levelQ, levelQMul, c2Q1, c2Q2, c2Q2): pool = ModUpSplitQP(c2Q1);
    //      c2Q2 = MRed(c2Q2 + (qm - pool), Q^-1 mod qm)
    LG_TRY(tmp.alloc((size_t)3 * B * nM * N));
    const size_t ts = (size_t)nM * N;
    LG_TRY(lgi_modup_launch(e->q1q2->qp, N, 3 * B, T.d, ws, nQ, tmp.d, ts, nM, 0, st));
    u64* tM = T.d + (size_t)nQ * N;
    LG_TRY(lgi_ew(EW_SUB_MULMONT_SCALAR, M, id, nM, 3 * B, tM, ws, tmp.d, ts, tM, ws, e->q1q2->moddown_qp.data(), nM, st));
    // :457 AddScalarBigint(c2Q2, pHalf)
    LG_TRY(lgi_ew(EW_ADD_SCALAR, M, id, nM, 3 * B, tM, ws, nullptr, 0, tM, ws, e->phalf_m.data(), nM, st));
    // :458 ModUpSplitPQ(levelQMul, c2Q2, ctOut.value[i])  (into the Q part of T, free after the ModDown)
    LG_TRY(lgi_modup_launch(e->q1q2->pq, N, 3 * B, tM, ws, nM, T.d, ws, nQ, 0, st));
    // :459 SubScalarBigint(ctOut.value[i], pHalf)
    LG_TRY(lgi_ew(EW_SUB_SCALAR, Q, id, nQ, 3 * B, T.d, ws, nullptr, 0, T.d, ws, e->phalf_q.data(), nQ, st));
    // :462 MulScalar(ctOut.value[i], t), written to the receivers
    std::vector<u64> tv(nQ, e->t);
    for (int i = 0; i < 3; ++i)
        LG_TRY(lgi_ew(EW_MUL_SCALAR, Q, id, nQ, B, T.d + (size_t)i * B * ws, ws, nullptr, 0, outs[i]->d, outs[i]->bstride,
                      tv.data(), nQ, st));
    return LG_OK;
}

int lg_bfv_switch_keys_core(lg_bfv_eval* e, const lg_poly* cx, const lg_swk* evk, lg_poly* p0, lg_poly* p1, lg_stream_t s) {
    LG_REQUIRE(e, "switchKeys: null evaluator");
    LG_ON_DEVICE(e->Q->device);
    const u64 N = e->Q->N;
    const int nQ = e->Q->nl;
    LG_TRY(check_p(cx, N, nQ, -1, "switchKeys"));
    LG_TRY(check_p(p0, N, nQ, cx->batch, "switchKeys"));
    LG_TRY(check_p(p1, N, nQ, cx->batch, "switchKeys"));
    return bfv_switch_keys(e, cx->batch, cx->d, cx->bstride, evk, p0->d, p0->bstride, false, p1->d, p1->bstride, false, cs(s));
}

// relinearize, bfv/evaluator.go:480-500, for a degree-2 ciphertext (evk = evakey[0])
int lg_bfv_relinearize(lg_bfv_eval* e, const lg_poly* c0, const lg_poly* c1, const lg_poly* c2, const lg_swk* rlk, lg_poly* out0,
                       lg_poly* out1, lg_stream_t s) {
    LG_REQUIRE(e, "Relinearize: null evaluator");
    LG_ON_DEVICE(e->Q->device);
    const lg_ring* Q = e->Q;
    const u64 N = Q->N;
    const int nQ = Q->nl;
    LG_TRY(check_p(c0, N, nQ, -1, "Relinearize"));
    const int B = c0->batch;
    LG_TRY(check_p(c1, N, nQ, B, "Relinearize"));
    LG_TRY(check_p(c2, N, nQ, B, "Relinearize"));
    LG_TRY(check_p(out0, N, nQ, B, "Relinearize"));
    LG_TRY(check_p(out1, N, nQ, B, "Relinearize"));
    LG_REQUIRE(c2->d != out0->d && c2->d != out1->d, "Relinearize: value[2] must not alias the receiver");
    LG_TRY(copy_if_needed(Q, B, c0, out0, cs(s)));  // :484-487
    LG_TRY(copy_if_needed(Q, B, c1, out1, cs(s)));
    // :493-496
    return bfv_switch_keys(e, B, c2->d, c2->bstride, rlk, out0->d, out0->bstride, true, out1->d, out1->bstride, true, cs(s));
}

// SwitchKeys, bfv/evaluator.go:540-558
int lg_bfv_switch_keys(lg_bfv_eval* e, const lg_poly* c0, const lg_poly* c1, const lg_swk* k, lg_poly* out0, lg_poly* out1,
                       lg_stream_t s) {
    LG_REQUIRE(e, "SwitchKeys: null evaluator");
    LG_ON_DEVICE(e->Q->device);
    const lg_ring* Q = e->Q;
    const u64 N = Q->N;
    const int nQ = Q->nl;
    LG_TRY(check_p(c0, N, nQ, -1, "SwitchKeys"));
    const int B = c0->batch;
    LG_TRY(check_p(c1, N, nQ, B, "SwitchKeys"));
    LG_TRY(check_p(out0, N, nQ, B, "SwitchKeys"));
    LG_TRY(check_p(out1, N, nQ, B, "SwitchKeys"));
    LG_REQUIRE(out0->d != c1->d, "SwitchKeys: receiver value[0] must not alias input value[1]");
    LG_TRY(copy_if_needed(Q, B, c0, out0, cs(s)));
    return bfv_switch_keys(e, B, c1->d, c1->bstride, k, out0->d, out0->bstride, true, out1->d, out1->bstride, false, cs(s));
}

// permute, bfv/evaluator.go:711-733: RotateColumns with a direct key (:595, gen = galElRotColLeft[k])
// and RotateRows (:669, gen = 2N-1)
int lg_bfv_permute(lg_bfv_eval* e, const lg_poly* c0, const lg_poly* c1, uint64_t gen, const lg_swk* k, lg_poly* out0,
                   lg_poly* out1, lg_stream_t s) {
    LG_REQUIRE(e, "permute: null evaluator");
    LG_ON_DEVICE(e->Q->device);
    const lg_ring* Q = e->Q;
    const u64 N = Q->N;
    const int nQ = Q->nl;
    LG_TRY(check_p(c0, N, nQ, -1, "permute"));
    const int B = c0->batch;
    LG_TRY(check_p(c1, N, nQ, B, "permute"));
    LG_TRY(check_p(out0, N, nQ, B, "permute"));
    LG_TRY(check_p(out1, N, nQ, B, "permute"));
    cudaStream_t st = cs(s);
    const size_t bs = (size_t)nQ * N;
    Scratch el(st);
    LG_TRY(el.alloc((size_t)2 * B * bs));
    u64* el0 = el.d;
    u64* el1 = el.d + (size_t)B * bs;
    PermArgs a;
    a.T = Q->T;
    a.map = limb_map_identity();
    a.index = nullptr;
    a.gen = gen;
    a.in = c0->d;  // :723-724
    a.in_bs = c0->bstride;
    a.out = el0;
    a.out_bs = bs;
    lg_launch_permute_coeff(a, nQ, B, st);
    a.in = c1->d;
    a.in_bs = c1->bstride;
    a.out = el1;
    lg_launch_permute_coeff(a, nQ, B, st);
    LG_LAUNCH_CHECK();
    // :729-732
    LG_TRY(lgi_ew(EW_COPY, Q, limb_map_identity(), nQ, B, el0, bs, nullptr, 0, out0->d, out0->bstride, nullptr, 0, st));
    return bfv_switch_keys(e, B, el1, bs, k, out0->d, out0->bstride, true, out1->d, out1->bstride, false, st);
}

}  // extern "C"
