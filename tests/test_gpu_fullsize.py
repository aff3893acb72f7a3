"""Parity at BASELINE.json's full sizes.  The oracle is fast enough for ONE ciphertext at the
largest parameter sets, so each test compares one batch entry bit-for-bit against the oracle and
pins the rest of the batch with size-independent properties: every entry of a batched call equals
the single-entry call on the same input (batch independence), NTT -> InvNTT is the identity, and
duplicated inputs give duplicated outputs.

config 3: BFV PN15QP880  (bfv/params.go:80-87)   Mul + Relinearize + RotateColumns
config 4: CKKS PN16QP1761 (ckks/params.go:79-86) MulRelin + Rescale + RotateColumns at level 33
"""
import numpy as np
import pytest

from oracle import ring_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lg():
    import lattigpu
    from lattigpu import ring

    ring.set_device(0)
    return lattigpu


def uniform_ct(rng, Q, N, batch, deg=1):
    return np.ascontiguousarray(np.stack([rng.integers(0, q, size=(batch, deg + 1, N), dtype=np.uint64) for q in Q], axis=2))


def polys(lg, ct):
    return tuple(lg.ring.Poly.from_numpy(np.ascontiguousarray(ct[:, i])) for i in range(ct.shape[1]))


def host(ct, nl=None):
    return np.stack([p.numpy(nl=nl, squeeze=False) for p in ct], axis=1)


def test_ckks_pn16_mulrelin_rescale_rotate(lg):
    p = lg.ckks.DefaultParams[lg.ckks.PN16QP1761]
    N = 1 << p["LogN"]
    Q, P = lg.ckks.GenModuli(p)
    nQ, nP = len(Q), len(P)
    assert (N, nQ, nP) == (65536, 34, 4)
    beta = -(-nQ // nP)
    rng = np.random.default_rng(0x1A771C0 + 4)
    evk = np.ascontiguousarray(np.stack([rng.integers(0, q, size=(beta, 2, N), dtype=np.uint64) for q in Q + P], axis=2))
    batch = 3
    a, b = uniform_ct(rng, Q, N, batch), uniform_ct(rng, Q, N, batch)
    a[2], b[2] = a[0], b[0]  # duplicated input -> duplicated output
    cQ, cP = lg.ring.NewContextWithParams(N, Q), lg.ring.NewContextWithParams(N, P)
    ev = lg.ckks.NewEvaluator(cQ, cP)
    key = lg.ckks.SwitchingKey(evk)
    level = nQ - 1
    pa, pb = polys(lg, a), polys(lg, b)
    out = (lg.ring.Poly(N, nQ, batch), lg.ring.Poly(N, nQ, batch))
    ev.MulRelin(level, pa, pb, key, out)
    ev.Rescale(nQ, out)
    rot = (lg.ring.Poly(N, nQ, batch), lg.ring.Poly(N, nQ, batch))
    idx = lg.ring.PermuteNTTIndex(5, 1, N)
    ev.permuteNTT(level - 1, out, idx, key, rot)
    got_mr, got_rot = host(out, nQ - 1), host(rot, nQ - 1)
    # (1) one entry against the oracle, bit-exact
    oev = orc.CkksEvaluator(orc.Context(N, Q), orc.Context(N, P))
    w = oev.rescale(oev.mul_relin(level, np.ascontiguousarray(a[1]), np.ascontiguousarray(b[1]), evk))
    assert np.array_equal(got_mr[1], w)
    w = oev.permute_ntt(level - 1, w, orc.permute_ntt_index(5, 1, N), evk)
    assert np.array_equal(got_rot[1], w)
    # (2) duplicated inputs
    assert np.array_equal(got_mr[0], got_mr[2]) and np.array_equal(got_rot[0], got_rot[2])
    # (3) batch independence: entry 0 recomputed alone
    pa1, pb1 = polys(lg, a[:1]), polys(lg, b[:1])
    o1 = (lg.ring.Poly(N, nQ, 1), lg.ring.Poly(N, nQ, 1))
    ev.MulRelin(level, pa1, pb1, key, o1)
    ev.Rescale(nQ, o1)
    assert np.array_equal(host(o1, nQ - 1)[0], got_mr[0])
    # (4) NTT round trip over the whole batch at full size
    t = lg.ring.Poly(N, nQ, batch)
    cQ.InvNTT(pa[0], t)
    cQ.NTT(t, t)
    assert np.array_equal(t.numpy(squeeze=False), a[:, 0])
    # (5) RotateHoisted at full size: two rotations off one decomposition, one entry against the oracle,
    # duplicated inputs elsewhere (the same key stands in for both rotation keys)
    idxs = [lg.ring.PermuteNTTIndex(5, k, N) for k in (1, 7)]
    outs = [(lg.ring.Poly(N, nQ, batch), lg.ring.Poly(N, nQ, batch)) for _ in idxs]
    ev.RotateHoisted(level, pa, [(i, key) for i in idxs], outs)
    want = oev.rotate_hoisted(level, np.ascontiguousarray(a[1]), [orc.permute_ntt_index(5, k, N) for k in (1, 7)], [evk, evk])
    for o, w in zip(outs, want):
        g = host(o, nQ)
        assert np.array_equal(g[1], w)
        assert np.array_equal(g[0], g[2])


@pytest.mark.parametrize("level", [32, 30])
def test_ckks_pn16_partial_digit_levels(lg, level):
    """PN16QP1761 at full size below the top level: level 32 (digit 8 is a broadcast copy of limb 32) and level 30
    (beta = 8, digit 7 has three active limbs and takes modUpParams[7][1]) -- SURVEY.md Appendix A's worked cases.
    One entry against the oracle, duplicated inputs elsewhere."""
    p = lg.ckks.DefaultParams[lg.ckks.PN16QP1761]
    N = 1 << p["LogN"]
    Q, P = lg.ckks.GenModuli(p)
    nQ, nP = len(Q), len(P)
    beta = -(-nQ // nP)
    nl = level + 1
    rng = np.random.default_rng(0x1A771C0 + 40 + level)
    evk = np.ascontiguousarray(np.stack([rng.integers(0, q, size=(beta, 2, N), dtype=np.uint64) for q in Q + P], axis=2))
    batch = 2
    a, b = uniform_ct(rng, Q[:nl], N, batch), uniform_ct(rng, Q[:nl], N, batch)
    a[1], b[1] = a[0], b[0]
    cQ, cP = lg.ring.NewContextWithParams(N, Q), lg.ring.NewContextWithParams(N, P)
    ev = lg.ckks.NewEvaluator(cQ, cP)
    key = lg.ckks.SwitchingKey(evk)
    pa, pb = polys(lg, a), polys(lg, b)
    out = (lg.ring.Poly(N, nl, batch), lg.ring.Poly(N, nl, batch))
    ev.MulRelin(level, pa, pb, key, out)
    ev.Rescale(nl, out)
    got = host(out, nl - 1)
    oev = orc.CkksEvaluator(orc.Context(N, Q), orc.Context(N, P))
    w = oev.rescale(oev.mul_relin(level, np.ascontiguousarray(a[0]), np.ascontiguousarray(b[0]), evk))
    assert np.array_equal(got[0], w)
    assert np.array_equal(got[0], got[1])


def test_bfv_pn15_mul_relin_rotate(lg):
    p = lg.bfv.DefaultParams[lg.bfv.PN15QP880]
    N = 1 << p["LogN"]
    Q, P, QMul = lg.bfv.GenModuli(p)
    nQ, nP = len(Q), len(P)
    assert (N, nQ, nP, len(QMul)) == (32768, 12, 3, 12)
    beta = -(-nQ // nP)
    rng = np.random.default_rng(0x1A771C0 + 3)
    evk = np.ascontiguousarray(np.stack([rng.integers(0, q, size=(beta, 2, N), dtype=np.uint64) for q in Q + P], axis=2))
    batch = 3
    a, b = uniform_ct(rng, Q, N, batch), uniform_ct(rng, Q, N, batch)
    a[2], b[2] = a[0], b[0]
    cQ, cM, cP = (lg.ring.NewContextWithParams(N, m) for m in (Q, QMul, P))
    ev = lg.bfv.NewEvaluator(cQ, cM, cP, p["T"])
    key = lg.ckks.SwitchingKey(evk)
    pa, pb = polys(lg, a), polys(lg, b)
    d2 = tuple(lg.ring.Poly(N, nQ, batch) for _ in range(3))
    ev.Mul(pa, pb, d2)
    d1 = (lg.ring.Poly(N, nQ, batch), lg.ring.Poly(N, nQ, batch))
    ev.Relinearize(d2, key, d1)
    rot = (lg.ring.Poly(N, nQ, batch), lg.ring.Poly(N, nQ, batch))
    gen = pow(lg.bfv.GaloisGen, 1, 2 * N)
    ev.permute(d1, gen, key, rot)
    got = host(rot)
    oev = orc.BfvEvaluator(orc.Context(N, Q), orc.Context(N, QMul), orc.Context(N, P), p["T"])
    w = oev.tensor_and_rescale(np.ascontiguousarray(a[1]), np.ascontiguousarray(b[1]))
    assert np.array_equal(host(d2)[1], w)
    w = oev.relinearize(w, evk)
    w = oev.permute(w, gen, evk)
    assert np.array_equal(got[1], w)
    assert np.array_equal(got[0], got[2])


def test_bfv_pn15_scheme_pipeline(lg):
    """BASELINE config 3 at full size (BFV PN15QP880), whole scheme pipeline on the device: keygen -> encode ->
    encrypt -> Mul -> Relinearize -> RotateColumns -> decrypt -> decode.  Size-independent property: the decoded
    slots are the rotated slot-wise product mod t; one batch entry is also compared with the oracle bit for bit."""
    p = lg.bfv.DefaultParams[lg.bfv.PN15QP880]
    N, t = 1 << p["LogN"], p["T"]
    Q, P, QMul = lg.bfv.GenModuli(p)
    nQ = len(Q)
    R = lg.ring
    rng = np.random.default_rng(0x1A771C0 + 33)
    cQ, cM, cP = (R.NewContextWithParams(N, m) for m in (Q, QMul, P))
    kg = lg.bfv_scheme.KeyGenerator(cQ, cP)
    enc = lg.bfv_scheme.Encoder(cQ, t)
    ev = lg.bfv.NewEvaluator(cQ, cM, cP, t)
    tern = lambda *s: rng.integers(-1, 2, size=s + (N,))
    gauss = lambda *s: np.rint(rng.normal(0, 3.2, size=s + (N,))).astype(np.int64)
    unif = lambda: np.ascontiguousarray(np.stack([rng.integers(0, q, size=N, dtype=np.uint64) for q in Q + P]))
    sk_c = tern()
    sk = kg.GenSecretKey(sk_c)
    e_pk, a_pk = gauss(), unif()
    pk = kg.GenPublicKey(sk, e_pk, a_pk)
    rl_e, rl_u = [gauss() for _ in range(kg.beta)], [unif() for _ in range(kg.beta)]
    rlk, rlk_host = kg.GenRelinKey(sk, rl_e, rl_u)
    k = 5
    gen = pow(lg.bfv.GaloisGen, k, 2 * N)
    ro_e, ro_u = [gauss() for _ in range(kg.beta)], [unif() for _ in range(kg.beta)]
    rot, rot_host = kg.genrotkey(sk, gen, ro_e, ro_u)
    E = lg.bfv_scheme.Encryptor(cQ, cP, kg.contextQP, pk=pk, sk=sk)
    D = lg.bfv_scheme.Decryptor(cQ, sk)
    batch = 4
    m0 = rng.integers(0, t, size=(batch, N), dtype=np.uint64)
    m1 = rng.integers(0, t, size=(batch, N), dtype=np.uint64)
    new = lambda n: tuple(R.Poly(N, nQ, batch) for _ in range(n))
    (pt0, pt1), ct0, ct1 = new(2), new(2), new(2)
    enc.EncodeUint(m0, pt0)
    enc.EncodeUint(m1, pt1)
    u, e0, e1 = tern(batch), gauss(batch), gauss(batch)
    E.EncryptPk(pt0, ct0, u, e0, e1)
    E.EncryptPk(pt1, ct1, tern(batch), gauss(batch), gauss(batch))
    ct2, ctr, cto = new(3), new(2), new(2)
    ev.Mul(ct0, ct1, ct2)
    ev.Relinearize(ct2, rlk, ctr)
    ev.permute(ctr, gen, rot, cto)
    (dec,) = new(1)
    D.Decrypt(cto, dec)
    slots = enc.DecodeUint(dec)
    row = N // 2
    prod = (m0.astype(np.uint64) * m1.astype(np.uint64)) % np.uint64(t)  # t^2 < 2^64
    want = np.concatenate([np.roll(prod[:, :row], -k, axis=1), np.roll(prod[:, row:], -k, axis=1)], axis=1)
    assert np.array_equal(slots, want)
    # oracle, entry 0: keys, first ciphertext and the decode of the final plaintext
    S = orc.BfvScheme(Q, P, N, t)
    osk = S.gen_secret_key(sk_c)
    opk = S.gen_public_key(osk, e_pk, a_pk)
    assert np.array_equal(rlk_host, S.gen_relin_key(osk, rl_e, rl_u))
    o0 = S.encrypt_pk(S.encode_uint(m0[0]), opk, u[0], e0[0], e1[0])
    assert np.array_equal(host(ct0)[0], o0)
    final = host(cto)[0]
    odec = S.decrypt(final, osk)
    assert np.array_equal(dec.numpy(squeeze=False)[0], odec)
    assert np.array_equal(S.decode_uint(odec), slots[0])
