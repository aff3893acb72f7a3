"""The reference's big-integer property tests (ring/ring_test.go) restated
against the oracle, so that the parts of the path that have no golden vectors
(basis extension, rescaling, Galois permutations, decomposition, key switching)
are anchored the same way the reference anchors them: against math/big.
"""
import random

import numpy as np
import pytest

from oracle import ring_oracle as orc

QI60 = [1152921504066306049, 1152921504057917441, 1152921504053723137, 1152921504050839553]  # ring/params.go:50-69 tail
PI60 = [576460752568975361, 576460752573431809, 576460752580902913, 576460752585490433]  # ring/params.go:28-47 tail
M64 = (1 << 64) - 1


def prod(xs):
    r = 1
    for x in xs:
        r *= x
    return r


def crt_poly(values, moduli):
    return np.array([[v % q for v in values] for q in moduli], dtype=np.uint64)


def crt_reconstruct(poly, moduli):
    Q = prod(moduli)
    out = []
    for j in range(poly.shape[1]):
        x = 0
        for i, q in enumerate(moduli):
            Qi = Q // q
            x += int(poly[i, j]) * Qi * pow(Qi, -1, q)
        out.append(x % Q)
    return out


def div_round(a, b):
    """ring/int.go:40-52 for a >= 0, b > 0"""
    q, r = divmod(a, b)
    return q + 1 if 2 * r >= b else q


def test_bred_mred_vs_bigint():
    # ring/ring_test.go:352-420
    rng = random.Random(1)
    L = orc.lib()
    for q in QI60 + PI60 + [0x200000008001, 0x400018001]:
        u = (orc.u64 * 2)(*orc.bred_params(q))
        assert (int(u[0]) << 64) + int(u[1]) == (1 << 128) // q
        qinv = L.orc_mred_params(q)
        assert (qinv * q) & M64 == 1
        for _ in range(500):
            x, y = rng.randrange(q), rng.randrange(q)
            assert L.orc_bred(x, y, q, u) == x * y % q
            # MRed(x, MForm(y)) == x*y mod q
            ym = L.orc_mform(y, q, u)
            assert ym == (y << 64) % q
            assert L.orc_mred(x, ym, q, qinv) == x * y % q
            assert L.orc_invmform(ym, q, qinv) == y
            w = rng.getrandbits(64)
            assert L.orc_bred_add(w, q, u) == w % q
            assert L.orc_modexp(x, y, q) == pow(x, y, q)


def test_generate_ntt_primes_and_ckks_moduli():
    # SURVEY.md Appendix B values, ckks/params.go:59-66,79-86
    Q, P, _ = orc.gen_moduli(14, [45] + [34] * 9, [43, 43])
    assert Q[:3] == [0x200000008001, 0x400018001, 0x400060001] and len(Q) == 10
    assert P == [0x80000050001, 0x800000B8001]
    assert sum(q.bit_length() for q in Q + P) - 0 >= 438
    Q, P, _ = orc.gen_moduli(16, [55] + [45] * 33, [55] * 4)
    assert Q[:3] == [0x80000000080001, 0x2000000A0001, 0x2000000E0001]
    assert P[:3] == [0x80000000440001, 0x80000000500001, 0x800000005E0001]
    for q in Q + P:
        assert q % (2 << 16) == 1


@pytest.mark.parametrize("N", [16, 64])
def test_mulpoly_vs_naive(N):
    # ring/ring_test.go:503-548: NTT-based product == schoolbook negacyclic product
    rng = random.Random(2)
    moduli = QI60[:2]
    ctx = orc.Context(N, moduli)
    a = np.array([[rng.randrange(q) for _ in range(N)] for q in moduli], dtype=np.uint64)
    b = np.array([[rng.randrange(q) for _ in range(N)] for q in moduli], dtype=np.uint64)
    na, nb = ctx.ntt(a), ctx.ntt(b)
    c = ctx.invntt(ctx.op3("mulcoeffs", na, nb))
    nam = ctx.op2("mform_poly", na)
    c2 = ctx.invntt(ctx.op3("mulcoeffs_montgomery", nam, nb))
    for i, q in enumerate(moduli):
        want = [0] * N
        for x in range(N):
            for y in range(N):
                k = x + y
                v = int(a[i, x]) * int(b[i, y])
                if k >= N:
                    want[k - N] = (want[k - N] - v) % q
                else:
                    want[k] = (want[k] + v) % q
        assert [int(v) for v in c[i]] == want
        assert [int(v) for v in c2[i]] == want


def test_ring_method_set_extras():
    """the oracle's MulPoly*, Shift, Rotate, Exp, Equal restatements against the reference's own property tests
    (ring/ring_test.go:422-450 GaloisShift, :503-548 MulPoly) and plain integer arithmetic"""
    N, moduli = 64, QI60[:2]
    rng = random.Random(5)
    ctx = orc.Context(N, moduli)
    a = np.array([[rng.randrange(q) for _ in range(N)] for q in moduli], dtype=np.uint64)
    b = np.array([[rng.randrange(q) for _ in range(N)] for q in moduli], dtype=np.uint64)
    want = orc.mul_poly_naive(ctx, a, b)
    assert np.array_equal(orc.mul_poly(ctx, a, b), want)
    am, bm = ctx.op2("mform_poly", a), ctx.op2("mform_poly", b)
    assert np.array_equal(ctx.op2("invmform_poly", orc.mul_poly(ctx, am, bm, montgomery=True)), want)
    assert np.array_equal(orc.mul_poly_naive(ctx, am, b, montgomery=True), want)
    for i, q in enumerate(moduli):  # and the schoolbook negacyclic product on integers
        ref = [0] * N
        for x in range(N):
            for y in range(N):
                v = int(a[i, x]) * int(b[i, y])
                k = x + y
                ref[k % N] = (ref[k % N] + (v if k < N else -v)) % q
        assert [int(v) for v in want[i]] == ref
    # GaloisShift: BitReverse, InvNTT, Rotate(1), NTT, BitReverse, Reduce == Shift(1)
    br = [int(format(j, "0%db" % (N.bit_length() - 1))[::-1], 2) for j in range(N)]
    t = ctx.invntt(np.ascontiguousarray(a[:, br]))
    t = orc.ring_rotate(ctx, t, 1)
    t = ctx.ntt(t)
    t = ctx.op2("reduce", np.ascontiguousarray(t[:, br]))
    assert np.array_equal(t, orc.ring_shift(ctx, a, 1))
    assert np.array_equal(orc.ring_shift(ctx, a, N), a)
    with pytest.raises(IndexError):
        orc.ring_shift(ctx, a, N + 1)
    # Exp ends with InvNTT(NTT(p1)) in p2 (ring.go:463)
    p1, p2 = orc.ring_exp(ctx, a, 3)
    assert np.array_equal(p1, ctx.ntt(a)) and np.array_equal(p2, a)
    # Equal compares residues
    qcol = np.array(moduli, dtype=np.uint64)[:, None]
    eq, ra, rb = orc.ring_equal(ctx, a, a + qcol)
    assert eq and np.array_equal(ra, a) and np.array_equal(rb, a)
    c = a.copy()
    c[1, 0] ^= np.uint64(1)
    assert not orc.ring_equal(ctx, a, c)[0] and orc.ring_equal(ctx, a, c, level=0)[0]


@pytest.mark.parametrize("srcdst", [(QI60, PI60), (QI60[:2], QI60[:2]), (PI60[:3], QI60)])
def test_extend_basis_vs_crt(srcdst):
    # ring/ring_test.go:550-585 (ModUpSplitQP equals reduction of the big integer)
    src, dst = srcdst
    N = 32
    rng = random.Random(3)
    ctxQ, ctxP = orc.Context(N, src), orc.Context(N, dst)
    ext = orc.Extender(ctxQ, ctxP)
    vals = [rng.randrange(prod(src)) for _ in range(N)]
    vals[0] = 0  # (x = Q-1 is outside the float64 correction's exact range: the reference's v rounds up there)
    pol = crt_poly(vals, src)
    got = ext.modup_split_qp(len(src) - 1, pol)
    assert np.array_equal(got, crt_poly(vals, dst))


def test_div_round_floor_vs_bigint():
    # ring/ring_test.go:134-220
    N = 32
    rng = random.Random(4)
    moduli = QI60
    ctx = orc.Context(N, moduli)
    vals = [rng.randrange(prod(moduli)) // 10 for _ in range(N)]
    pol = crt_poly(vals, moduli)
    nb = len(moduli) - 1
    want_r, want_f = list(vals), list(vals)
    for j in range(nb):
        want_r = [div_round(v, moduli[-1 - j]) for v in want_r]
        want_f = [v // moduli[-1 - j] for v in want_f]
    p = pol.copy()
    orc.lib().orc_div_round_by_last_modulus_many(ctx.h, len(moduli), orc.ptr(p), nb)
    assert np.array_equal(p[:1], crt_poly(want_r, moduli[:1]))
    p = pol.copy()
    orc.lib().orc_div_floor_by_last_modulus_many(ctx.h, len(moduli), orc.ptr(p), nb)
    assert np.array_equal(p[:1], crt_poly(want_f, moduli[:1]))
    # NTT-domain variants agree with the coefficient-domain ones
    for name in ("div_round_by_last_modulus", "div_floor_by_last_modulus"):
        p = pol.copy()
        getattr(orc.lib(), "orc_" + name)(ctx.h, len(moduli), orc.ptr(p))
        pn = ctx.ntt(pol)
        getattr(orc.lib(), "orc_" + name + "_ntt")(ctx.h, len(moduli), orc.ptr(pn))
        back = ctx.invntt(np.ascontiguousarray(pn[:-1]))
        assert np.array_equal(back, p[:-1])


def test_galois_shift():
    # ring/ring_test.go:422-460: NTT(Permute(a, g)) == PermuteNTT(NTT(a), g), and
    # the automorphism is X -> X^g on big-int coefficients
    N = 64
    rng = random.Random(5)
    moduli = QI60[:2]
    ctx = orc.Context(N, moduli)
    a = np.array([[rng.randrange(q) for _ in range(N)] for q in moduli], dtype=np.uint64)
    for g in (5, 25, 2 * N - 1, pow(5, 7, 2 * N)):
        perm = ctx.permute(a, g)
        out = np.zeros_like(a)
        orc.lib().orc_permute_ntt(N, 2, orc.ptr(ctx.ntt(a)), g, orc.ptr(out))
        assert np.array_equal(ctx.ntt(perm), out)
        for i, q in enumerate(moduli):
            want = [0] * N
            for k in range(N):
                e = (k * g) % (2 * N)
                want[e % N] = (q - int(a[i, k])) if e >= N else int(a[i, k])
            assert [int(v) for v in perm[i]] == want
    idx = orc.permute_ntt_index(5, 3, N)
    out = np.zeros_like(a)
    orc.lib().orc_permute_ntt(N, 2, orc.ptr(a), pow(5, 3, 2 * N), orc.ptr(out))
    assert np.array_equal(orc.permute_ntt_with_index(a, idx), out)


def _ckks_small(N=32, nQ=5, nP=2, logq=40, logp=50):
    logn = N.bit_length() - 1
    Q, P, _ = orc.gen_moduli(logn, [logq + 5] + [logq] * (nQ - 1), [logp] * nP)
    return Q, P


@pytest.mark.parametrize("level", [4, 3, 2, 0])
def test_decompose_digits_reconstruct(level):
    """Each digit of DecomposeAndSplit is the integer D_i = [x]_{digit basis}
    (in [0, prod digit)) reduced modulo every target prime: the exactness claim
    of modUpExact (ring_basis_extension.go:352-393)."""
    N = 32
    Q, P = _ckks_small(N)
    alpha = len(P)
    rng = random.Random(6)
    dec = orc.Decomposer(Q, P, N)
    nl = level + 1
    vals = [rng.randrange(prod(Q[:nl])) for _ in range(N)]
    pol = crt_poly(vals, Q[:nl])
    full = np.zeros((len(Q), N), dtype=np.uint64)
    full[:nl] = pol
    beta = -(-nl // alpha)
    for crt in range(beta):
        lo, hi = crt * alpha, min(crt * alpha + alpha, nl)
        digit_mod = Q[lo:hi]
        dvals = [v % prod(digit_mod) for v in vals]
        gotQ, gotP = dec.decompose_and_split(level, crt, full)
        assert np.array_equal(gotQ, crt_poly(dvals, Q[:nl])), (level, crt)
        assert np.array_equal(gotP, crt_poly(dvals, P)), (level, crt)
        got = dec.decompose(level, crt, full)
        assert np.array_equal(got, np.concatenate([gotQ, gotP]))


def _keygen(rng, ctxQP, Q, P, N, sk_in, sk_out):
    """ckks/keygen.go:282-340 newSwitchingKey with a toy error (|e| <= 1) --
    returns evk [beta][2][nQP][N] in NTT + Montgomery form."""
    nQ, nP = len(Q), len(P)
    alpha = nP
    beta = -(-nQ // alpha)
    QP = Q + P
    Pprod = prod(P)
    # P * skIn (coefficient ints -> NTT)
    s_in_ntt = ctxQP.op2("mform_poly", ctxQP.ntt(crt_poly([Pprod * s for s in sk_in], QP)))
    s_out_ntt = ctxQP.op2("mform_poly", ctxQP.ntt(crt_poly(sk_out, QP)))
    evk = np.zeros((beta, 2, nQ + nP, N), dtype=np.uint64)
    for i in range(beta):
        e = [rng.choice([-1, 0, 1]) for _ in range(N)]
        k0 = ctxQP.op2("mform_poly", ctxQP.ntt(crt_poly(e, QP)))
        a = np.array([[rng.randrange(q) for _ in range(N)] for q in QP], dtype=np.uint64)
        for j in range(alpha):
            index = i * alpha + j
            qi = QP[index]
            k0[index] = (k0[index].astype(object) + s_in_ntt[index].astype(object)) % qi  # CRed(p1+p0) on reduced inputs
            if index >= nQ - 1:
                break
        k0 = np.ascontiguousarray(k0.astype(np.uint64))
        ctxQP.op3("mulcoeffs_montgomery_and_sub", a, s_out_ntt, k0)
        evk[i, 0], evk[i, 1] = k0, a
    return evk


@pytest.mark.parametrize("level", [4, 3, 1])
def test_ckks_keyswitch_semantics(level):
    """switchKeysInPlace (ckks/evaluator.go:1475-1558) with a key built as
    newSwitchingKey does: p0 + p1*s_out must equal cx*s_in up to a small error
    relative to Q (the decryption check of ckks_test.go, at ring level)."""
    N = 32
    Q, P = _ckks_small(N)
    rng = random.Random(7 + level)
    ctxQ, ctxP, ctxQP = orc.Context(N, Q), orc.Context(N, P), orc.Context(N, Q + P)
    ev = orc.CkksEvaluator(ctxQ, ctxP)
    sk_in = [rng.choice([-1, 0, 1]) for _ in range(N)]
    sk_out = [rng.choice([-1, 0, 1]) for _ in range(N)]
    evk = _keygen(rng, ctxQP, Q, P, N, sk_in, sk_out)
    nl = level + 1
    Ql = Q[:nl]
    cx_vals = [rng.randrange(prod(Ql)) for _ in range(N)]
    cx = np.zeros((len(Q), N), dtype=np.uint64)
    cx[:nl] = ctxQ.ntt(crt_poly(cx_vals, Ql), nl=nl)
    p0, p1 = ev.switch_keys_in_place(level, cx, evk)

    def negacyclic(a, b, mod):
        out = [0] * N
        for x in range(N):
            for y in range(N):
                k = x + y
                if k >= N:
                    out[k - N] = (out[k - N] - a[x] * b[y]) % mod
                else:
                    out[k] = (out[k] + a[x] * b[y]) % mod
        return out

    Qp = prod(Ql)
    ctxl = orc.Context(N, Ql)
    p0c = crt_reconstruct(ctxl.invntt(p0), Ql)
    p1c = crt_reconstruct(ctxl.invntt(p1), Ql)
    lhs = [(u + v) % Qp for u, v in zip(p0c, negacyclic(p1c, sk_out, Qp))]
    rhs = negacyclic(cx_vals, sk_in, Qp)
    err = max(min((l - r) % Qp, (r - l) % Qp) for l, r in zip(lhs, rhs))
    # noise ~ beta * N * q_digit^alpha / P plus rounding: tiny next to Q
    assert err.bit_length() < 30, err.bit_length()


@pytest.mark.parametrize("level", [4, 2])
def test_ckks_rotate_hoisted_semantics(level):
    """RotateHoisted / switchKeyHoisted (ckks/evaluator.go:1252-1392): for every requested rotation
    out0 + out1*s must equal sigma_g(c0 + c1*s) up to key-switch noise, with ONE decomposition of c1
    shared by all rotations; the rotation key switches sigma_g(s) -> s."""
    N = 32
    Q, P = _ckks_small(N)
    rng = random.Random(40 + level)
    ctxQ, ctxP, ctxQP = orc.Context(N, Q), orc.Context(N, P), orc.Context(N, Q + P)
    ev = orc.CkksEvaluator(ctxQ, ctxP)
    sk = [rng.choice([-1, 0, 1]) for _ in range(N)]
    nl = level + 1
    Ql = Q[:nl]
    Qp = prod(Ql)
    ctxl = orc.Context(N, Ql)

    def sigma(vals, g, mod):  # X -> X^g on coefficient vectors (ring_galois.go:106-127)
        out = [0] * N
        for i, v in enumerate(vals):
            k = (i * g) % (2 * N)
            if k >= N:
                out[k - N] = (-v) % mod
            else:
                out[k] = v % mod
        return out

    def negacyclic(a, b, mod):
        out = [0] * N
        for x in range(N):
            for y in range(N):
                k = x + y
                if k >= N:
                    out[k - N] = (out[k - N] - a[x] * b[y]) % mod
                else:
                    out[k] = (out[k] + a[x] * b[y]) % mod
        return out

    c0v = [rng.randrange(Qp) for _ in range(N)]
    c1v = [rng.randrange(Qp) for _ in range(N)]
    ct = np.ascontiguousarray(np.stack([ctxQ.ntt(crt_poly(c0v, Ql), nl=nl), ctxQ.ntt(crt_poly(c1v, Ql), nl=nl)]))
    gens = [pow(5, k, 2 * N) for k in (1, 3)]
    indexes = [orc.permute_ntt_index(5, k, N) for k in (1, 3)]
    evks = [_keygen(rng, ctxQP, Q, P, N, [(x if x <= 1 else x - 3) for x in sigma(sk, g, 3)], sk) for g in gens]
    outs = ev.rotate_hoisted(level, ct, indexes, evks)
    msg = [(u + v) % Qp for u, v in zip(c0v, negacyclic(c1v, sk, Qp))]
    for g, out in zip(gens, outs):
        o0 = crt_reconstruct(ctxl.invntt(np.ascontiguousarray(out[0])), Ql)
        o1 = crt_reconstruct(ctxl.invntt(np.ascontiguousarray(out[1])), Ql)
        lhs = [(u + v) % Qp for u, v in zip(o0, negacyclic(o1, sk, Qp))]
        rhs = sigma(msg, g, Qp)
        err = max(min((l - r) % Qp, (r - l) % Qp) for l, r in zip(lhs, rhs))
        assert err.bit_length() < 30, (g, err.bit_length())


def test_ckks_keygen_encrypt_decrypt_semantics():
    """Key generator / encryptor / decryptor ring sequences (ckks/keygen.go, encryptor.go, decryptor.go):
    pk and sk encryption (fast and ModDown paths) decrypt to the message, and a relinearisation key built
    by newSwitchingKey makes MulRelin + Rescale decrypt to the product (what ckks_test.go checks through
    the encoder, here at ring level with integer messages)."""
    N = 32
    Q, P = _ckks_small(N)
    rng = random.Random(77)
    S = orc.CkksScheme(Q, P, N)
    nQP = len(Q) + len(P)
    tern = lambda: [rng.choice([-1, 0, 1]) for _ in range(N)]
    gauss = lambda: [rng.choice([-2, -1, 0, 0, 1, 2]) for _ in range(N)]
    unif = lambda mods: np.array([[rng.randrange(q) for _ in range(N)] for q in mods], dtype=np.uint64)
    sk_c = tern()
    sk = S.gen_secret_key(sk_c)
    pk = S.gen_public_key(sk, gauss(), unif(Q + P))
    rlk = S.gen_relin_key(sk, [gauss() for _ in range(S.beta)], [unif(Q + P) for _ in range(S.beta)])
    level = len(Q) - 1
    Qp = prod(Q)
    scale = 1 << 20

    def negacyclic(a, b):
        out = [0] * N
        for x in range(N):
            for y in range(N):
                k = x + y
                if k >= N:
                    out[k - N] -= a[x] * b[y]
                else:
                    out[k] += a[x] * b[y]
        return out

    def centered(vals, mod):
        return [v if v < mod // 2 else v - mod for v in vals]

    def plaintext(m):
        return S.Q.ntt(crt_poly([x * scale for x in m], Q))

    m0 = [rng.randrange(-500, 500) for _ in range(N)]
    m1 = [rng.randrange(-500, 500) for _ in range(N)]
    cts = {
        "pk": S.encrypt_pk(level, plaintext(m0), pk, tern(), gauss(), gauss()),
        "pk_fast": S.encrypt_pk(level, plaintext(m0), pk, tern(), gauss(), gauss(), fast=True),
        "sk": S.encrypt_sk(level, plaintext(m0), sk, unif(Q + P), gauss()),
        "sk_fast": S.encrypt_sk(level, plaintext(m0), sk, unif(Q), gauss(), fast=True),
    }
    for name, ct in cts.items():
        dec = centered(crt_reconstruct(S.Q.invntt(S.decrypt(level, ct, sk)), Q), Qp)
        err = max(abs(d - x * scale) for d, x in zip(dec, m0))
        assert err < (1 << 12), (name, err)
    ct1 = S.encrypt_pk(level, plaintext(m1), pk, tern(), gauss(), gauss())
    ev = orc.CkksEvaluator(S.Q, S.P)
    prod_ct = ev.rescale(ev.mul_relin(level, np.ascontiguousarray(cts["pk"]), np.ascontiguousarray(ct1), rlk))
    Ql = Q[:-1]
    ctxl = orc.Context(N, Ql)
    dec = centered(crt_reconstruct(ctxl.invntt(S.decrypt(level - 1, prod_ct, sk)), Ql), prod(Ql))
    want = negacyclic(m0, m1)
    err = max(abs(d - div_round(w * scale * scale, Q[-1])) for d, w in zip(dec, want))
    assert err < (1 << 12), err
    # a rotation key switches sigma_5(sk) back to sk: permuteNTT then decrypts to the rotated message
    rot = S.gen_rot_key(sk, 5, [gauss() for _ in range(S.beta)], [unif(Q + P) for _ in range(S.beta)])
    idx = orc.permute_ntt_index(5, 1, N)
    r = ev.permute_ntt(level, np.ascontiguousarray(cts["pk"]), idx, rot)
    dec = centered(crt_reconstruct(S.Q.invntt(S.decrypt(level, r, sk)), Q), Qp)
    want = [0] * N
    for i, v in enumerate(m0):
        k = (i * 5) % (2 * N)
        if k >= N:
            want[k - N] = -v
        else:
            want[k] = v
    err = max(abs(d - w * scale) for d, w in zip(dec, want))
    assert err < (1 << 12), err


def test_dckks_cks_rtg_rkg_semantics():
    """dckks CKS / RTG / RKG (keyswitching.go, rotkey_gen.go, relinkey_gen.go) with 3 parties, as
    dckks_test.go does: the collective relinearisation key relinearises, the collective rotation key rotates,
    and CKS re-encrypts from the sum of the input shares to the sum of the output shares."""
    N, parties = 32, 3
    Q, P = _ckks_small(N)
    rng = random.Random(123)
    S = orc.CkksScheme(Q, P, N)
    D = orc.DckksProtocols(S)
    K = S.QP
    tern = lambda: [rng.choice([-1, 0, 1]) for _ in range(N)]
    gauss = lambda: [rng.choice([-2, -1, 0, 0, 1, 2]) for _ in range(N)]
    gl = lambda: [gauss() for _ in range(S.beta)]
    unif = lambda mods: np.array([[rng.randrange(q) for _ in range(N)] for q in mods], dtype=np.uint64)
    level = len(Q) - 1
    Qp = prod(Q)
    scale = 1 << 20

    def centered(vals, mod):
        return [v if v < mod // 2 else v - mod for v in vals]

    def negacyclic(a, b):
        out = [0] * N
        for x in range(N):
            for y in range(N):
                k = x + y
                if k >= N:
                    out[k - N] -= a[x] * b[y]
                else:
                    out[k] += a[x] * b[y]
        return out

    sks = [S.gen_secret_key(tern()) for _ in range(parties)]
    sk = sks[0]
    for x in sks[1:]:
        sk = K.op3("add", sk, x)
    plaintext = lambda m: S.Q.ntt(crt_poly([x * scale for x in m], Q))
    m0 = [rng.randrange(-500, 500) for _ in range(N)]
    m1 = [rng.randrange(-500, 500) for _ in range(N)]
    ct0 = S.encrypt_sk(level, plaintext(m0), sk, unif(Q + P), gauss())
    ct1 = S.encrypt_sk(level, plaintext(m1), sk, unif(Q + P), gauss())
    ev = orc.CkksEvaluator(S.Q, S.P)

    # RKG, three rounds
    crp = [unif(Q + P) for _ in range(S.beta)]
    us = [S.gen_secret_key(tern()) for _ in range(parties)]
    r1 = None
    for u, s_i in zip(us, sks):
        sh = D.rkg_round1(u, s_i, crp, gl())
        r1 = sh if r1 is None else D.add_lists(r1, sh)
    r2 = None
    for s_i in sks:
        sh = D.rkg_round2(r1, s_i, crp, gl(), gl())
        r2 = sh if r2 is None else D.add_pairs(r2, sh)
    r3 = None
    for u, s_i in zip(us, sks):
        sh = D.rkg_round3(r2, u, s_i, gl())
        r3 = sh if r3 is None else D.add_lists(r3, sh)
    rlk = D.rkg_key(r2, r3)
    prod_ct = ev.rescale(ev.mul_relin(level, np.ascontiguousarray(ct0), np.ascontiguousarray(ct1), rlk))
    Ql = Q[:-1]
    dec = centered(crt_reconstruct(orc.Context(N, Ql).invntt(S.decrypt(level - 1, prod_ct, sk)), Ql), prod(Ql))
    want = negacyclic(m0, m1)
    err = max(abs(d - div_round(w * scale * scale, Q[-1])) for d, w in zip(dec, want))
    assert err < (1 << 14), err

    # RTG for the Galois element 5
    crp = [unif(Q + P) for _ in range(S.beta)]
    agg = None
    for s_i in sks:
        sh = D.rtg_gen_share(s_i, 5, crp, gl())
        agg = sh if agg is None else D.add_lists(agg, sh)
    rot = D.rtg_finalize(agg, crp)
    r = ev.permute_ntt(level, np.ascontiguousarray(ct0), orc.permute_ntt_index(5, 1, N), rot)
    dec = centered(crt_reconstruct(S.Q.invntt(S.decrypt(level, r, sk)), Q), Qp)
    wantr = [0] * N
    for i, v in enumerate(m0):
        k = (i * 5) % (2 * N)
        if k >= N:
            wantr[k - N] = -v
        else:
            wantr[k] = v
    err = max(abs(d - w * scale) for d, w in zip(dec, wantr))
    assert err < (1 << 14), err

    # CKS from sum(sks) to sum(sks_out)
    sks_out = [S.gen_secret_key(tern()) for _ in range(parties)]
    sk_out = sks_out[0]
    for x in sks_out[1:]:
        sk_out = K.op3("add", sk_out, x)
    comb = None
    for a, b in zip(sks, sks_out):
        sh = D.cks_gen_share(level, a, b, ct0[1], gauss())
        comb = sh if comb is None else S.Q.op3("add", comb, sh)
    switched = np.stack([S.Q.op3("add", np.ascontiguousarray(ct0[0]), comb), ct0[1]])
    dec = centered(crt_reconstruct(S.Q.invntt(S.decrypt(level, switched, sk_out)), Q), Qp)
    err = max(abs(d - x * scale) for d, x in zip(dec, m0))
    assert err < (1 << 14), err


def test_dbfv_cks_semantics():
    """dbfv CKS (dbfv/keyswitching.go:66-122) on a coefficient-domain ciphertext: c0 + sum(shares) + c1*s_out
    equals c0 + c1*s_in up to the smudging noise."""
    N, parties = 32, 3
    Q, P, _ = _bfv_small(N)
    rng = random.Random(321)
    S = orc.CkksScheme(Q, P, N)  # contexts Q, P, QP + extender; nothing CKKS-specific is used
    D = orc.DckksProtocols(S)
    Qp = prod(Q)
    tern = lambda: [rng.choice([-1, 0, 1]) for _ in range(N)]
    sin_c, sout_c = [tern() for _ in range(parties)], [tern() for _ in range(parties)]
    sks_in = [S.gen_secret_key(c) for c in sin_c]
    sks_out = [S.gen_secret_key(c) for c in sout_c]
    c0v = [rng.randrange(Qp) for _ in range(N)]
    c1v = [rng.randrange(Qp) for _ in range(N)]
    c1 = crt_poly(c1v, Q)
    comb = None
    for a, b in zip(sks_in, sks_out):
        sh = D.bfv_cks_gen_share(a, b, c1, [rng.randrange(-40, 41) for _ in range(N)])
        comb = sh if comb is None else S.Q.op3("add", comb, sh)
    combv = crt_reconstruct(comb, Q)

    def negacyclic(a, b):
        out = [0] * N
        for x in range(N):
            for y in range(N):
                k = x + y
                if k >= N:
                    out[k - N] = (out[k - N] - a[x] * b[y]) % Qp
                else:
                    out[k] = (out[k] + a[x] * b[y]) % Qp
        return out

    s_in = [sum(c[i] for c in sin_c) for i in range(N)]
    s_out = [sum(c[i] for c in sout_c) for i in range(N)]
    lhs = [(u + v + w) % Qp for u, v, w in zip(c0v, combv, negacyclic(c1v, s_out))]
    rhs = [(u + w) % Qp for u, w in zip(c0v, negacyclic(c1v, s_in))]
    err = max(min((l - r) % Qp, (r - l) % Qp) for l, r in zip(lhs, rhs))
    assert err < (1 << 12), err


@pytest.mark.parametrize("second", [0, 1], ids=["dckks", "dbfv"])
def test_rkg_naive_semantics(second):
    """relinkey_gen_naive.go with 3 parties: the two-round key relinearises a product (the dckks file loses round
    one's second error sample, the dbfv file keeps it; both are valid keys)."""
    N, parties = 32, 3
    Q, P = _ckks_small(N)
    rng = random.Random(91 + second)
    S = orc.CkksScheme(Q, P, N)
    D = orc.DckksProtocols(S)
    K = S.QP
    tern = lambda: [rng.choice([-1, 0, 1]) for _ in range(N)]
    gauss = lambda: [rng.choice([-2, -1, 0, 0, 1, 2]) for _ in range(N)]
    unif = lambda mods: np.array([[rng.randrange(q) for _ in range(N)] for q in mods], dtype=np.uint64)
    tern_mont = lambda: K.op2("mform_poly", orc.signed_residues(Q + P, tern()))
    level = len(Q) - 1
    scale = 1 << 20
    sks = [S.gen_secret_key(tern()) for _ in range(parties)]
    sk = sks[0]
    for x in sks[1:]:
        sk = K.op3("add", sk, x)
    pk = S.gen_public_key(sk, gauss(), unif(Q + P))
    r1 = None
    for s_i in sks:
        sh = D.rkg_naive_round1(s_i, pk, [(gauss(), gauss()) for _ in range(S.beta)], [tern_mont() for _ in range(S.beta)], second)
        r1 = sh if r1 is None else D.add_pairs(r1, sh)
    r2 = None
    for s_i in sks:
        sh = D.rkg_naive_round2(r1, s_i, pk, [tern_mont() for _ in range(S.beta)], [(gauss(), gauss()) for _ in range(S.beta)])
        r2 = sh if r2 is None else D.add_pairs(r2, sh)
    rlk = D.rkg_naive_key(r2)
    plaintext = lambda m: S.Q.ntt(crt_poly([x * scale for x in m], Q))
    m0 = [rng.randrange(-500, 500) for _ in range(N)]
    m1 = [rng.randrange(-500, 500) for _ in range(N)]
    ct0 = S.encrypt_sk(level, plaintext(m0), sk, unif(Q + P), gauss())
    ct1 = S.encrypt_sk(level, plaintext(m1), sk, unif(Q + P), gauss())
    ev = orc.CkksEvaluator(S.Q, S.P)
    prod_ct = ev.rescale(ev.mul_relin(level, np.ascontiguousarray(ct0), np.ascontiguousarray(ct1), rlk))
    Ql = Q[:-1]
    dec = crt_reconstruct(orc.Context(N, Ql).invntt(S.decrypt(level - 1, prod_ct, sk)), Ql)
    dec = [v if v < prod(Ql) // 2 else v - prod(Ql) for v in dec]
    want = [0] * N
    for x in range(N):
        for y in range(N):
            k = x + y
            if k >= N:
                want[k - N] -= m0[x] * m1[y]
            else:
                want[k] += m0[x] * m1[y]
    err = max(abs(d - div_round(w * scale * scale, Q[-1])) if w >= 0 else abs(d + div_round(-w * scale * scale, Q[-1]))
              for d, w in zip(dec, want))
    assert err < (1 << 16), err


def test_dckks_refresh_semantics():
    """dckks Refresh (public_refresh.go:43-147) with 3 parties, as dckks_test.go's testRefresh: a ciphertext at a low
    level is masked-decrypted, recoded at the top level and re-encrypted under the common reference polynomial; it
    must decrypt (collective key) to the same message at the top level."""
    N, parties = 32, 3
    Q, P = _ckks_small(N)
    rng = random.Random(77)
    S = orc.CkksScheme(Q, P, N)
    R = orc.DckksRefresh(S)
    K = S.QP
    nQ = len(Q)
    tern = lambda: [rng.choice([-1, 0, 1]) for _ in range(N)]
    gauss = lambda: [rng.choice([-2, -1, 0, 0, 1, 2]) for _ in range(N)]
    unif = lambda mods: np.array([[rng.randrange(q) for _ in range(N)] for q in mods], dtype=np.uint64)
    sks = [S.gen_secret_key(tern()) for _ in range(parties)]
    sk = sks[0]
    for x in sks[1:]:
        sk = K.op3("add", sk, x)
    scale = 1 << 20
    m = [rng.randrange(-500, 500) for _ in range(N)]
    for level_start in (0, 1):
        nl = level_start + 1
        # encrypted at the top level, then the upper limbs dropped (what a ciphertext that went down the levels is)
        pt = S.Q.ntt(crt_poly([x * scale for x in m], Q))
        ct = np.ascontiguousarray(S.encrypt_sk(nQ - 1, pt, sk, unif(Q + P), gauss())[:, :nl])
        crs = unif(Q)
        bound = prod(Q[:nl]) // (2 * parties)
        h0 = h1 = None
        for s_i in sks:
            a, b = R.gen_shares(s_i, level_start, parties, ct[1], crs, [rng.randrange(bound) for _ in range(N)], gauss(), gauss())
            assert a.shape[0] == nl and b.shape[0] == nQ
            h0, h1 = (a, b) if h0 is None else (R.aggregate(h0, a), R.aggregate(h1, b))
        fresh = R.recrypt(R.recode(R.decrypt(np.ascontiguousarray(ct[0]), h0)), crs, h1)
        assert fresh.shape == (2, nQ, N)
        dec = crt_reconstruct(S.Q.invntt(S.decrypt(nQ - 1, fresh, sk)), Q)
        Qp = prod(Q)
        dec = [v if v < Qp // 2 else v - Qp for v in dec]
        err = max(abs(d - x * scale) for d, x in zip(dec, m))
        assert err < (1 << 12), (level_start, err)
    # the CRT helpers are the big-integer maps of ring_context.go:343-421
    vals = [rng.randrange(-prod(Q) // 2, prod(Q) // 2) for _ in range(N)]
    assert orc.poly_to_bigint(orc.set_coefficients_bigint(Q, vals), Q) == [v % prod(Q) for v in vals]


def test_dbfv_refresh_semantics():
    """dbfv Refresh (public_refresh.go:105-205) with 3 parties, as dbfv_test.go's testRefresh: after Finalize the
    ciphertext decrypts (collective key) to the same plaintext slots, with fresh noise."""
    N, parties, t = 32, 3, 65537
    Q, P, _ = _bfv_small(N)
    rng = random.Random(78)
    S = orc.BfvScheme(Q, P, N, t)
    K = S.QP
    tern = lambda: [rng.choice([-1, 0, 1]) for _ in range(N)]
    gauss = lambda: [rng.choice([-2, -1, 0, 0, 1, 2]) for _ in range(N)]
    unif = lambda mods: np.array([[rng.randrange(q) for _ in range(N)] for q in mods], dtype=np.uint64)
    sks = [S.gen_secret_key(tern()) for _ in range(parties)]
    sk = sks[0]
    for x in sks[1:]:
        sk = K.op3("add", sk, x)
    slots = [rng.randrange(t) for _ in range(N)]
    ct = S.encrypt_sk(S.encode_uint(slots), sk, unif(Q + P), gauss())
    # some noise growth a refresh is meant to remove: add an encryption of zero scaled up
    crs = unif(Q + P)
    protos = [orc.DbfvRefresh(S) for _ in range(parties)]
    share = None
    for pr, s_i in zip(protos, sks):
        sh = pr.gen_shares(s_i, ct[1], crs, gauss(), gauss(), [rng.randrange(t) for _ in range(N)])
        share = sh if share is None else pr.aggregate(share, sh)
    fresh = protos[0].finalize(ct, crs, share)
    assert np.array_equal(S.decode_uint(S.decrypt(fresh, sk)), np.array(slots, dtype=np.uint64))
    # hP is state: a second GenShares on the same object differs from a fresh object's by the first error's P limbs
    e1, e2, e3, mask = gauss(), gauss(), gauss(), [rng.randrange(t) for _ in range(N)]
    again = protos[0].gen_shares(sks[0], ct[1], crs, e2, e3, mask)
    clean = orc.DbfvRefresh(S).gen_shares(sks[0], ct[1], crs, e2, e3, mask)
    assert np.array_equal(again[1], clean[1]) and not np.array_equal(again[0], clean[0])


def test_ckks_const_ops_semantics():
    """Constant ops (ckks/evaluator.go:373-833) in the coefficient domain: AddConst(a+bi) adds round(a*scale)
    to coefficient 0 and round(b*scale) to coefficient N/2; MultByConst multiplies the polynomial by
    round(a*scale) + round(b*scale)*X^(N/2); MultByi / DivByi multiply by X^(N/2) / X^(3N/2)."""
    N = 32
    Q, _ = _ckks_small(N)
    rng = random.Random(91)
    ctx = orc.Context(N, Q)
    Qp = prod(Q)
    vals = [rng.randrange(Qp) for _ in range(N)]
    p = ctx.ntt(crt_poly(vals, Q))
    level = len(Q) - 1
    scale = float(1 << 30)

    def coeffs(poly):
        return crt_reconstruct(ctx.invntt(np.ascontiguousarray(poly)), Q)

    def mul_monomial(v, k, c=1):  # c * X^k * v in Z_Q[X]/(X^N+1)
        out = [0] * N
        for i, x in enumerate(v):
            d = (i + k) % (2 * N)
            if d >= N:
                out[d - N] = (out[d - N] - c * x) % Qp
            else:
                out[d] = (out[d] + c * x) % Qp
        return out

    a, b = 3.25, -1.5
    A, B = int(a * scale + 0.5), -int(-b * scale + 0.5)
    got = coeffs(orc.ckks_const_op(ctx, "add", level, [p], [p], a, b, scale)[0])
    want = list(vals)
    want[0] = (want[0] + A) % Qp
    want[N // 2] = (want[N // 2] + B) % Qp
    assert got == want
    got = coeffs(orc.ckks_const_op(ctx, "mul", level, [p], [p], a, b, scale)[0])
    want = [(x + y) % Qp for x, y in zip(mul_monomial(vals, 0, A), mul_monomial(vals, N // 2, B))]
    assert got == want
    acc = ctx.ntt(crt_poly([7] * N, Q))
    got = coeffs(orc.ckks_const_op(ctx, "mul_add", level, [p], [acc], 5.0, 0.0, 1.0)[0])
    assert got == [(7 + 5 * x) % Qp for x in vals]
    assert coeffs(orc.ckks_const_op(ctx, "mul_i", level, [p], [p])[0]) == mul_monomial(vals, N // 2)
    assert coeffs(orc.ckks_const_op(ctx, "div_i", level, [p], [p])[0]) == mul_monomial(vals, 3 * N // 2)


def test_unreduced_inputs_are_defined():
    """NewPolyUniform feeds full 64-bit words (ring_object.go:26-46); the oracle
    must be total on them (used later as a formula-exactness probe for CUDA)."""
    N = 64
    rng = np.random.default_rng(8)
    ctx = orc.Context(N, QI60)
    a = rng.integers(0, 1 << 64, size=(4, N), dtype=np.uint64)
    out = ctx.ntt(a)
    assert all((out[i] < QI60[i]).all() for i in range(4))
    out = ctx.invntt(a)
    assert all((out[i] < QI60[i]).all() for i in range(4))


# ---------------------------------------------------------------------------
# BFV drivers (bfv/evaluator.go:278-813), anchored like bfv/bfv_test.go: the ring-level
# plaintext must come out right
# ---------------------------------------------------------------------------
def _bfv_small(N=32):
    logn = N.bit_length() - 1
    Q, P, QMul = orc.gen_moduli(logn, [39, 39, 38], [40, 40], [60, 60, 60])
    return Q, P, QMul


def test_bfv_tensor_and_rescale_semantics():
    """Mul of noiseless encryptions (Delta*m, 0) x (Delta*m', 0): value[0] must be Delta*(m*m' mod t) up to
    the rounding error of the t/Q scaling, value[1] = value[2] = 0-ish (bfv/evaluator.go:278-464)."""
    N, t = 32, 65537
    Q, P, QMul = _bfv_small(N)
    rng = random.Random(31)
    ev = orc.BfvEvaluator(orc.Context(N, Q), orc.Context(N, QMul), orc.Context(N, P), t)
    Qp = prod(Q)
    delta = Qp // t
    m0 = [rng.randrange(t) for _ in range(N)]
    m1 = [rng.randrange(t) for _ in range(N)]
    ct0 = np.stack([crt_poly([delta * m for m in m0], Q), np.zeros((len(Q), N), np.uint64)])
    ct1 = np.stack([crt_poly([delta * m for m in m1], Q), np.zeros((len(Q), N), np.uint64)])
    out = ev.tensor_and_rescale(np.ascontiguousarray(ct0), np.ascontiguousarray(ct1))
    mm = [0] * N
    for x in range(N):
        for y in range(N):
            k = x + y
            if k >= N:
                mm[k - N] -= m0[x] * m1[y]
            else:
                mm[k] += m0[x] * m1[y]
    got = crt_reconstruct(out[0], Q)
    for k in range(N):
        v = got[k] if got[k] < Qp // 2 else got[k] - Qp
        # decrypt: round(t * v / Q) mod t
        dec = ((2 * t * v + Qp) // (2 * Qp)) % t
        assert dec == mm[k] % t, k
    for i in (1, 2):
        v = crt_reconstruct(out[i], Q)
        assert all(min(x, Qp - x) < (1 << 20) for x in v)
    sq = ev.tensor_and_rescale(np.ascontiguousarray(ct0), np.ascontiguousarray(ct0))
    assert sq.shape == out.shape


def test_bfv_keyswitch_semantics():
    """bfv switchKeys (bfv/evaluator.go:736-813) with a key built as bfv/keygen.go newSwitchingKey does
    (same structure as the ckks one): p0 + p1*s_out = cx*s_in + small, in the coefficient domain."""
    N, t = 32, 65537
    Q, P, QMul = _bfv_small(N)
    rng = random.Random(32)
    ctxQ, ctxP, ctxQP = orc.Context(N, Q), orc.Context(N, P), orc.Context(N, Q + P)
    ev = orc.BfvEvaluator(ctxQ, orc.Context(N, QMul), ctxP, t)
    sk_in = [rng.choice([-1, 0, 1]) for _ in range(N)]
    sk_out = [rng.choice([-1, 0, 1]) for _ in range(N)]
    evk = _keygen(rng, ctxQP, Q, P, N, sk_in, sk_out)
    Qp = prod(Q)
    cx_vals = [rng.randrange(Qp) for _ in range(N)]
    cx = crt_poly(cx_vals, Q)
    p0, p1 = ev.switch_keys_core(cx, evk)

    def negacyclic(a, b, mod):
        out = [0] * N
        for x in range(N):
            for y in range(N):
                k = x + y
                if k >= N:
                    out[k - N] = (out[k - N] - a[x] * b[y]) % mod
                else:
                    out[k] = (out[k] + a[x] * b[y]) % mod
        return out

    p0c, p1c = crt_reconstruct(p0, Q), crt_reconstruct(p1, Q)
    lhs = [(u + v) % Qp for u, v in zip(p0c, negacyclic(p1c, sk_out, Qp))]
    rhs = negacyclic(cx_vals, sk_in, Qp)
    err = max(min((l - r) % Qp, (r - l) % Qp) for l, r in zip(lhs, rhs))
    assert err.bit_length() < 30, err.bit_length()


def test_poly_wire_format():
    """ring.Poly.MarshalBinary / UnmarshalBinary (ring_object.go:146-289): header bytes, big-endian words,
    limb-major order, round trip, and the length check of UnmarshalBinary."""
    N = 16
    p = np.arange(3 * N, dtype=np.uint64).reshape(3, N) * np.uint64(0x0102030405060708)
    data = orc.poly_marshal(p)
    assert len(data) == 2 + 3 * N * 8 and data[0] == 4 and data[1] == 3
    assert data[2 + 8:2 + 16] == int(p[0, 1]).to_bytes(8, "big")
    assert data[2 + N * 8:2 + N * 8 + 8] == int(p[1, 0]).to_bytes(8, "big")
    assert np.array_equal(orc.poly_unmarshal(data), p)
    with pytest.raises(ValueError, match="invalid polynomial encoding"):
        orc.poly_unmarshal(data[:-8])
