"""GPU parity for the dckks / dbfv protocol steps on the ring hot path (config 5): CKG and PCKS
GenShare / AggregateShares / KeySwitch, with 3 in-process parties as dckks/dckks_test.go:63,146-168
does.  The expected values are the same op sequences composed from the oracle's ring ops; samples
are supplied by the test (the reference samples on the host)."""
import numpy as np
import pytest

from oracle import ring_oracle as orc

pytestmark = pytest.mark.gpu
PARTIES = 3


@pytest.fixture(scope="module")
def lg():
    import lattigpu
    from lattigpu import ring

    ring.set_device(0)
    return lattigpu


def uni(rng, moduli, N):
    return np.ascontiguousarray(np.stack([rng.integers(0, q, size=(N,), dtype=np.uint64) for q in moduli]))


def small(rng, moduli, N, bound):
    """a small signed sample reduced mod each prime (what the host samplers produce)"""
    v = rng.integers(-bound, bound + 1, size=(N,))
    return np.ascontiguousarray(np.stack([np.where(v < 0, q + v, v).astype(np.uint64) for q in moduli]))


def ternary_mont(rng, oK, moduli, N):
    v = small(rng, moduli, N, 1)
    return oK.op2("mform_poly", v)


@pytest.mark.parametrize("scheme", ["dckks", "dbfv"])
def test_ckg_pcks(lg, scheme):
    if scheme == "dckks":
        p = lg.ckks.DefaultParams[lg.ckks.PN13QP218]
        Q, P = lg.ckks.GenModuli(p)
    else:
        p = lg.bfv.DefaultParams[lg.bfv.PN13QP218]
        Q, P, _ = lg.bfv.GenModuli(p)
    N = 1 << p["LogN"]
    QP = Q + P
    nQ = len(Q)
    rng = np.random.default_rng(61)
    oQ, oP, oK = orc.Context(N, Q), orc.Context(N, P), orc.Context(N, QP)
    oext = orc.Extender(oQ, oP)
    cQ, cP, cK = (lg.ring.NewContextWithParams(N, m) for m in (Q, P, QP))
    mod = lg.dckks if scheme == "dckks" else lg.dbfv
    F = lg.ring.Poly.from_numpy

    # ---- CKG: share_i = NTT(e_i) - sk_i * crs ; pk0 = sum_i share_i ---------------------------
    ckg = mod.CKGProtocol(cK)
    crs = uni(rng, QP, N)
    sks = [oK.op2("mform_poly", oK.ntt(small(rng, QP, N, 1))) for _ in range(PARTIES)]  # NTT + Montgomery secrets
    es = [small(rng, QP, N, 19) for _ in range(PARTIES)]
    want = None
    agg = None
    for sk, e in zip(sks, es):
        w = oK.ntt(e)
        oK.op3("mulcoeffs_montgomery_and_sub", sk, crs, w)
        want = w if want is None else oK.op3("add", want, w)
        share = ckg.AllocateShares()
        ckg.GenShare(F(sk), F(crs), share, F(e))
        assert np.array_equal(share.numpy(), w)
        if agg is None:
            agg = share
        else:
            ckg.AggregateShares(agg, share, agg)
    assert np.array_equal(agg.numpy(), want)

    # ---- PCKS ----------------------------------------------------------------------------------
    pk = (want, crs)
    level = nQ - 1 if scheme == "dbfv" else nQ - 2
    nl = level + 1
    ct = [uni(rng, Q, N), uni(rng, Q, N)]
    pcks = mod.PCKSProtocol(cQ, cP, cK)
    comb_w, comb = None, None
    for i in range(PARTIES):
        u = ternary_mont(rng, oK, QP, N)
        e0, e1 = small(rng, QP, N, 40), small(rng, QP, N, 19)
        skq = np.ascontiguousarray(sks[i][:nQ])
        t = oK.ntt(u)
        s0 = oK.op3("mulcoeffs_montgomery", t, pk[0])
        s1 = oK.op3("mulcoeffs_montgomery", t, pk[1])
        if scheme == "dckks":
            s0 = oK.op3("add", s0, oK.ntt(e0))
            s1 = oK.op3("add", s1, oK.ntt(e1))
            w0 = oext.moddown_ntt_pq(level, s0)
            w1 = oext.moddown_ntt_pq(level, s1)
            oQ.op3("mulcoeffs_montgomery_and_add", np.ascontiguousarray(ct[1][:nl]), np.ascontiguousarray(skq[:nl]), w0, nl=nl)
            share = pcks.AllocateShares(level)
            pcks.GenShare(level, F(skq), (F(pk[0]), F(pk[1])), F(ct[1]), share, F(u), F(e0), F(e1))
        else:
            s0 = oK.op3("add", oK.invntt(s0), e0)
            s1 = oK.op3("add", oK.invntt(s1), e1)
            w0 = oext.moddown_pq(level, s0)
            w1 = oext.moddown_pq(level, s1)
            tt = oQ.invntt(oQ.op3("mulcoeffs_montgomery", oQ.ntt(ct[1]), skq))
            w0 = oQ.op3("add", w0, tt)
            share = pcks.AllocateShares()
            pcks.GenShare(F(skq), (F(pk[0]), F(pk[1])), F(ct[1]), share, F(u), F(e0), F(e1))
        assert np.array_equal(share[0].numpy(nl=nl), w0) and np.array_equal(share[1].numpy(nl=nl), w1), i
        if comb is None:
            comb, comb_w = share, [w0, w1]
        else:
            if scheme == "dckks":
                pcks.AggregateShares(comb, share, comb, level)
            else:
                pcks.AggregateShares(comb, share, comb)
            comb_w = [oQ.op3("add", comb_w[0], w0, nl=nl), oQ.op3("add", comb_w[1], w1, nl=nl)]
    out = (cQ.NewPoly(), cQ.NewPoly())
    if scheme == "dckks":
        pcks.KeySwitch(comb, (F(ct[0]), F(ct[1])), out, level)
    else:
        pcks.KeySwitch(comb, (F(ct[0]), F(ct[1])), out)
    assert np.array_equal(out[0].numpy(nl=nl), oQ.op3("add", np.ascontiguousarray(ct[0][:nl]), comb_w[0], nl=nl))
    assert np.array_equal(out[1].numpy(nl=nl), comb_w[1])


def test_dckks_cks_rtg_rkg(lg):
    """CKS (dckks/keyswitching.go:55-108), RTG (rotkey_gen.go:75-174) and the three rounds of RKG
    (relinkey_gen.go:60-223) with 3 parties, every share and the final keys bit-exact against the oracle."""
    p = lg.ckks.DefaultParams[lg.ckks.PN13QP218]
    Q, P = lg.ckks.GenModuli(p)
    N = 1 << p["LogN"]
    QP = Q + P
    nQ = len(Q)
    rng = np.random.default_rng(67)
    S = orc.CkksScheme(Q, P, N)
    D = orc.DckksProtocols(S)
    cQ, cP, cK = (lg.ring.NewContextWithParams(N, m) for m in (Q, P, QP))
    F = lg.ring.Poly.from_numpy
    from lattigpu.ckks_scheme import signed_to_poly

    tern = lambda: rng.integers(-1, 2, size=N)
    gauss = lambda: np.rint(rng.normal(0, 3.2, size=N)).astype(np.int64)
    gl = lambda: [gauss() for _ in range(S.beta)]
    dev = lambda es: [signed_to_poly(cK, e) for e in es]
    sks = [S.gen_secret_key(tern()) for _ in range(PARTIES)]
    us = [S.gen_secret_key(tern()) for _ in range(PARTIES)]
    crp = [uni(rng, QP, N) for _ in range(S.beta)]
    dcrp = [F(c) for c in crp]

    def same_list(got, want):
        return all(np.array_equal(g.numpy(), w) for g, w in zip(got, want))

    # ---- RKG ----
    rkg = lg.dckks.RKGProtocol(cQ, cP, cK)
    r1 = r1w = None
    for u, s_i in zip(us, sks):
        e = gl()
        sh, w = rkg.GenShareRoundOne(F(u), F(s_i), dcrp, dev(e)), D.rkg_round1(u, s_i, crp, e)
        assert same_list(sh, w)
        if r1 is None:
            r1, r1w = sh, w
        else:
            rkg.AggregateShareRoundOne(r1, sh, r1)
            r1w = D.add_lists(r1w, w)
    r2 = r2w = None
    for s_i in sks:
        e1, e2 = gl(), gl()
        sh, w = rkg.GenShareRoundTwo(r1, F(s_i), dcrp, dev(e1), dev(e2)), D.rkg_round2(r1w, s_i, crp, e1, e2)
        assert all(np.array_equal(a[0].numpy(), b[0]) and np.array_equal(a[1].numpy(), b[1]) for a, b in zip(sh, w))
        if r2 is None:
            r2, r2w = sh, w
        else:
            rkg.AggregateShareRoundTwo(r2, sh, r2)
            r2w = D.add_pairs(r2w, w)
    r3 = r3w = None
    for u, s_i in zip(us, sks):
        e = gl()
        sh, w = rkg.GenShareRoundThree(r2, F(u), F(s_i), dev(e)), D.rkg_round3(r2w, u, s_i, e)
        assert same_list(sh, w)
        if r3 is None:
            r3, r3w = sh, w
        else:
            rkg.AggregateShareRoundThree(r3, sh, r3)
            r3w = D.add_lists(r3w, w)
    key = lg.dckks.evakey_to_numpy(rkg.GenRelinearizationKey(r2, r3))
    assert np.array_equal(key, D.rkg_key(r2w, r3w))

    # ---- RTG, rotation by 3 (Galois element 5^3) ----
    rtg = lg.dckks.RTGProtocol(cQ, cP, cK)
    gal = pow(5, 3, 2 * N)
    agg = aggw = None
    for s_i in sks:
        e = gl()
        sh, w = rtg.genShare(F(s_i), gal, dcrp, dev(e)), D.rtg_gen_share(s_i, gal, crp, e)
        assert same_list(sh, w)
        if agg is None:
            agg, aggw = sh, w
        else:
            rtg.Aggregate(agg, sh, agg)
            aggw = D.add_lists(aggw, w)
    assert np.array_equal(lg.dckks.evakey_to_numpy(rtg.Finalize(agg, dcrp)), D.rtg_finalize(aggw, crp))

    # ---- CKS at the top level and one below ----
    cks = lg.dckks.CKSProtocol(cQ, cP, cK)
    sks_out = [S.gen_secret_key(tern()) for _ in range(PARTIES)]
    ct = [uni(rng, Q, N), uni(rng, Q, N)]
    for level in (nQ - 1, nQ - 2):
        nl = level + 1
        comb = combw = None
        for a, b in zip(sks, sks_out):
            e = np.rint(rng.normal(0, 40.0, size=N)).astype(np.int64)
            share = cks.AllocateShare(level)
            cks.GenShare(level, F(np.ascontiguousarray(a[:nQ])), F(np.ascontiguousarray(b[:nQ])), F(ct[1]), share,
                         signed_to_poly(cK, e))
            w = D.cks_gen_share(level, a, b, ct[1], e)
            assert np.array_equal(share.numpy(nl=nl), w), level
            if comb is None:
                comb, combw = share, w
            else:
                cks.AggregateShares(level, comb, share, comb)
                combw = S.Q.op3("add", combw, w, nl=nl)
        out = (cQ.NewPoly(), cQ.NewPoly())
        cks.KeySwitch(level, comb, (F(ct[0]), F(ct[1])), out)
        assert np.array_equal(out[0].numpy(nl=nl), S.Q.op3("add", np.ascontiguousarray(ct[0][:nl]), combw, nl=nl))
        assert np.array_equal(out[1].numpy(nl=nl), ct[1][:nl])


def test_dbfv_cks_rtg_rkg(lg):
    """dbfv CKS (dbfv/keyswitching.go:66-122) with 3 parties against the oracle; the BFV RTG / RKG mirrors are the
    dckks ring sequences (checked here on the BFV moduli for one share each)."""
    p = lg.bfv.DefaultParams[lg.bfv.PN13QP218]
    Q, P, _ = lg.bfv.GenModuli(p)
    N = 1 << p["LogN"]
    QP = Q + P
    nQ = len(Q)
    rng = np.random.default_rng(69)
    S = orc.CkksScheme(Q, P, N)
    D = orc.DckksProtocols(S)
    cQ, cP, cK = (lg.ring.NewContextWithParams(N, m) for m in (Q, P, QP))
    F = lg.ring.Poly.from_numpy
    from lattigpu.ckks_scheme import signed_to_poly

    tern = lambda: rng.integers(-1, 2, size=N)
    gauss = lambda: np.rint(rng.normal(0, 3.2, size=N)).astype(np.int64)
    sks = [S.gen_secret_key(tern()) for _ in range(PARTIES)]
    sks_out = [S.gen_secret_key(tern()) for _ in range(PARTIES)]
    ct = [uni(rng, Q, N), uni(rng, Q, N)]
    cks = lg.dbfv.CKSProtocol(cQ, cP, cK)
    comb = combw = None
    for a, b in zip(sks, sks_out):
        e = np.rint(rng.normal(0, 40.0, size=N)).astype(np.int64)
        share = cks.AllocateShare()
        cks.GenShare(F(np.ascontiguousarray(a[:nQ])), F(np.ascontiguousarray(b[:nQ])), F(ct[1]), share, signed_to_poly(cK, e))
        w = D.bfv_cks_gen_share(a, b, ct[1], e)
        assert np.array_equal(share.numpy(), w)
        if comb is None:
            comb, combw = share, w
        else:
            cks.AggregateShares(comb, share, comb)
            combw = S.Q.op3("add", combw, w)
    out = (cQ.NewPoly(), cQ.NewPoly())
    cks.KeySwitch(comb, (F(ct[0]), F(ct[1])), out)
    assert np.array_equal(out[0].numpy(), S.Q.op3("add", ct[0], combw)) and np.array_equal(out[1].numpy(), ct[1])

    crp = [uni(rng, QP, N) for _ in range(S.beta)]
    dcrp = [F(c) for c in crp]
    e = [gauss() for _ in range(S.beta)]
    rtg = lg.dbfv.RTGProtocol(cQ, cP, cK)
    got = rtg.genShare(F(sks[0]), pow(5, 2, 2 * N), dcrp, [signed_to_poly(cK, x) for x in e])
    assert all(np.array_equal(g.numpy(), w) for g, w in zip(got, D.rtg_gen_share(sks[0], pow(5, 2, 2 * N), crp, e)))
    rkg = lg.dbfv.RKGProtocol(cQ, cP, cK)
    u = S.gen_secret_key(tern())
    got = rkg.GenShareRoundOne(F(u), F(sks[0]), dcrp, [signed_to_poly(cK, x) for x in e])
    assert all(np.array_equal(g.numpy(), w) for g, w in zip(got, D.rkg_round1(u, sks[0], crp, e)))


def test_dckks_refresh(lg):
    """dckks Refresh (public_refresh.go:43-147) with 3 parties from two start levels: every share, the masked
    decryption, the recoded polynomial and the refreshed ciphertext bit-exact against the oracle."""
    p = lg.ckks.DefaultParams[lg.ckks.PN13QP218]
    Q, P = lg.ckks.GenModuli(p)
    N = 1 << p["LogN"]
    nQ = len(Q)
    rng = np.random.default_rng(71)
    S = orc.CkksScheme(Q, P, N)
    R = orc.DckksRefresh(S)
    cQ = lg.ring.NewContextWithParams(N, Q)
    F = lg.ring.Poly.from_numpy
    from lattigpu.ckks_scheme import signed_to_poly

    tern = lambda: rng.integers(-1, 2, size=N)
    gauss = lambda: np.rint(rng.normal(0, 3.2, size=N)).astype(np.int64)
    sks = [np.ascontiguousarray(S.gen_secret_key(tern())[:nQ]) for _ in range(PARTIES)]
    proto = lg.dckks.RefreshProtocol(cQ)
    for level in (0, 2):
        nl = level + 1
        ct = [uni(rng, Q[:nl], N), uni(rng, Q[:nl], N)]
        crs = uni(rng, Q, N)
        bound = 1
        for q in Q[:nl]:
            bound *= int(q)
        bound //= 2 * PARTIES
        nbytes = (bound.bit_length() + 7) // 8 + 8
        h0 = h1 = h0w = h1w = None
        for s_i in sks:
            raw = [int.from_bytes(rng.bytes(nbytes), "big") % bound for _ in range(N)]
            e0, e1 = gauss(), gauss()
            d, r = proto.AllocateShares(level)
            proto.GenShares(F(s_i), level, PARTIES, F(ct[1]), F(crs), d, r, raw, signed_to_poly(cQ, e0), signed_to_poly(cQ, e1))
            dw, rw = R.gen_shares(s_i, level, PARTIES, ct[1], crs, raw, e0, e1)
            assert np.array_equal(d.numpy(squeeze=False)[0], dw) and np.array_equal(r.numpy(squeeze=False)[0], rw), level
            if h0 is None:
                h0, h1, h0w, h1w = d, r, dw, rw
            else:
                proto.Aggregate(h0, d, h0)
                proto.Aggregate(h1, r, h1)
                h0w, h1w = R.aggregate(h0w, dw), R.aggregate(h1w, rw)
        c0 = F(ct[0])
        proto.Decrypt(level, c0, h0)
        md = R.decrypt(ct[0], h0w)
        assert np.array_equal(c0.numpy(squeeze=False)[0], md)
        rec = proto.Recode(level, c0)
        recw = R.recode(md)
        assert rec.nlimbs == nQ and np.array_equal(rec.numpy(), recw)
        o0, o1 = proto.Recrypt(rec, F(crs), h1)
        want = R.recrypt(recw, crs, h1w)
        assert np.array_equal(o0.numpy(), want[0]) and np.array_equal(o1.numpy(), want[1])


def test_dbfv_refresh(lg):
    """dbfv Refresh (public_refresh.go:105-205) with 3 parties: shares (including the second call on the same protocol
    object, which sees the hP words the first call left), aggregation and Finalize bit-exact against the oracle;
    the refreshed ciphertext decrypts to the encoded slots."""
    p = lg.bfv.DefaultParams[lg.bfv.PN13QP218]
    Q, P, _ = lg.bfv.GenModuli(p)
    N, t = 1 << p["LogN"], p["T"]
    QP = Q + P
    nQ = len(Q)
    rng = np.random.default_rng(73)
    S = orc.BfvScheme(Q, P, N, t)
    cQ, cP, cK = (lg.ring.NewContextWithParams(N, m) for m in (Q, P, QP))
    F = lg.ring.Poly.from_numpy
    from lattigpu.ckks_scheme import signed_to_poly

    tern = lambda: rng.integers(-1, 2, size=N)
    gauss = lambda: np.rint(rng.normal(0, 3.2, size=N)).astype(np.int64)
    sks = [S.gen_secret_key(tern()) for _ in range(PARTIES)]
    sk = sks[0]
    for x in sks[1:]:
        sk = S.QP.op3("add", sk, x)
    slots = rng.integers(0, t, size=N, dtype=np.uint64)
    ct = S.encrypt_sk(S.encode_uint(slots), sk, uni(rng, QP, N), gauss())
    crs = uni(rng, QP, N)
    share = sharew = None
    protos, oracles = [], []
    for s_i in sks:
        pr, ow = lg.dbfv.RefreshProtocol(cQ, cP, cK, t), orc.DbfvRefresh(S)
        protos.append(pr)
        oracles.append(ow)
        e, e2, mask = gauss(), gauss(), rng.integers(0, t, size=N, dtype=np.uint64)
        sh = pr.AllocateShares()
        pr.GenShares(F(s_i), F(ct[1]), F(crs), sh, signed_to_poly(cK, e), signed_to_poly(cK, e2), mask)
        w = ow.gen_shares(s_i, ct[1], crs, e, e2, mask)
        assert np.array_equal(sh[0].numpy(), w[0]) and np.array_equal(sh[1].numpy(), w[1])
        if share is None:
            share, sharew = sh, w
        else:
            pr.Aggregate(share, sh, share)
            sharew = ow.aggregate(sharew, w)
    out = (cQ.NewPoly(), cQ.NewPoly())
    protos[0].Finalize((F(ct[0]), F(ct[1])), F(crs), share, out)
    want = oracles[0].finalize(ct, crs, sharew)
    assert np.array_equal(out[0].numpy(), want[0]) and np.array_equal(out[1].numpy(), want[1])
    assert np.array_equal(S.decode_uint(S.decrypt(want, sk)), slots)
    # second GenShares on the same object: hP still holds the first sample's P limbs
    e, e2, mask = gauss(), gauss(), rng.integers(0, t, size=N, dtype=np.uint64)
    sh = protos[0].AllocateShares()
    protos[0].GenShares(F(sks[0]), F(ct[1]), F(crs), sh, signed_to_poly(cK, e), signed_to_poly(cK, e2), mask)
    w = oracles[0].gen_shares(sks[0], ct[1], crs, e, e2, mask)
    assert np.array_equal(sh[0].numpy(), w[0]) and np.array_equal(sh[1].numpy(), w[1])


@pytest.mark.parametrize("scheme", ["dckks", "dbfv"])
def test_rkg_naive(lg, scheme):
    """relinkey_gen_naive.go, two rounds with 3 parties, shares and key bit-exact against the oracle; the dckks file
    writes both round-one samples into shareOut[i][0] (:73,:75), the dbfv file into [0] and [1] (:74,:76)."""
    if scheme == "dckks":
        Q, P = lg.ckks.GenModuli(lg.ckks.DefaultParams[lg.ckks.PN13QP218])
        N, second, mod = 1 << 13, 0, lg.dckks
    else:
        Q, P, _ = lg.bfv.GenModuli(lg.bfv.DefaultParams[lg.bfv.PN13QP218])
        N, second, mod = 1 << 13, 1, lg.dbfv
    QP = Q + P
    rng = np.random.default_rng(75)
    S = orc.CkksScheme(Q, P, N)
    D = orc.DckksProtocols(S)
    cQ, cP, cK = (lg.ring.NewContextWithParams(N, m) for m in (Q, P, QP))
    F = lg.ring.Poly.from_numpy
    from lattigpu.ckks_scheme import signed_to_poly

    tern = lambda: rng.integers(-1, 2, size=N)
    gauss = lambda: np.rint(rng.normal(0, 3.2, size=N)).astype(np.int64)
    pairs = lambda: [(gauss(), gauss()) for _ in range(S.beta)]
    dev_pairs = lambda es: [(signed_to_poly(cK, a), signed_to_poly(cK, b)) for a, b in es]
    terns = lambda: [ternary_mont(rng, S.QP, QP, N) for _ in range(S.beta)]
    sks = [S.gen_secret_key(tern()) for _ in range(PARTIES)]
    sk = sks[0]
    for x in sks[1:]:
        sk = S.QP.op3("add", sk, x)
    pk = S.gen_public_key(sk, gauss(), uni(rng, QP, N))
    dpk = (F(pk[0]), F(pk[1]))
    rkg = mod.RKGProtocolNaive(cQ, cP, cK)

    def same(got, want):
        return all(np.array_equal(g[0].numpy(), w[0]) and np.array_equal(g[1].numpy(), w[1]) for g, w in zip(got, want))

    r1 = r1w = None
    for s_i in sks:
        e, u = pairs(), terns()
        sh, _ = rkg.AllocateShares()
        rkg.GenShareRoundOne(F(s_i), dpk, sh, dev_pairs(e), [F(x) for x in u])
        w = D.rkg_naive_round1(s_i, pk, e, u, second)
        assert same(sh, w)
        if r1 is None:
            r1, r1w = sh, w
        else:
            rkg.AggregateShareRoundOne(r1, sh, r1)
            r1w = D.add_pairs(r1w, w)
    r2 = r2w = None
    for s_i in sks:
        e, u = pairs(), terns()
        _, sh = rkg.AllocateShares()
        rkg.GenShareRoundTwo(r1, F(s_i), dpk, sh, [F(x) for x in u], dev_pairs(e))
        w = D.rkg_naive_round2(r1w, s_i, pk, u, e)
        assert same(sh, w)
        if r2 is None:
            r2, r2w = sh, w
        else:
            rkg.AggregateShareRoundTwo(r2, sh, r2)
            r2w = D.add_pairs(r2w, w)
    assert np.array_equal(lg.dckks.evakey_to_numpy(rkg.GenRelinearizationKey(r2)), D.rkg_naive_key(r2w))
