"""GPU parity for ring.SimpleScaler (ring/ring_scaling.go:166-300, Float128 accumulation on the device) and the BFV
key generator / encryptor / decryptor / batch encoder sequences (bfv/keygen.go, encryptor.go, decryptor.go,
encoder.go), bit-exact against the oracle restatement with the same sampled values, plus the BASELINE config-3
pipeline encode -> encrypt -> Mul -> Relinearize -> RotateColumns -> decrypt -> decode, device-resident."""
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


@pytest.mark.parametrize("t", [65537, 0x3EE0001, 1 << 16, 1 << 40], ids=["65537", "T_ref", "pow2_16", "pow2_40"])
@pytest.mark.parametrize("shape", [(10, [60, 60]), (12, [39, 39]), (13, [54, 54, 54]), (12, [59, 59, 59] + [58] * 9)],
                         ids=["2x60", "PN12", "PN13", "PN15moduli"])
@pytest.mark.parametrize("kind", ["reduced", "words"])
def test_simple_scaler(lg, t, shape, kind):
    logN, logq = shape
    N = 1 << logN
    Q, _, _ = orc.gen_moduli(logN, logq, [])
    ctx = lg.ring.NewContextWithParams(N, Q)
    octx = orc.Context(N, Q)
    sc, osc = lg.ring.NewSimpleScaler(t, ctx), orc.Scaler(t, octx)
    wi, ti = sc.params()
    owi, oti = osc.params()
    assert np.array_equal(wi, owi) and np.array_equal(ti, oti)
    rng = np.random.default_rng(t % 1000 + len(Q))
    batch = 3
    if kind == "words":
        p = rng.integers(0, 1 << 64, size=(batch, len(Q), N), dtype=np.uint64)
    else:
        p = np.ascontiguousarray(np.stack([rng.integers(0, q, size=(batch, N), dtype=np.uint64) for q in Q], axis=1))
        p[0, :, :4] = 0  # the ends of every residue range
        for i, q in enumerate(Q):
            p[0, i, 4:8] = q - 1
    p1 = lg.ring.Poly.from_numpy(p)
    for nl_out in (1, len(Q)):
        p2 = lg.ring.Poly(N, nl_out, batch)
        sc.Scale(p1, p2)
        got = p2.numpy(squeeze=False)
        for b in range(batch):
            assert np.array_equal(got[b], osc.scale(p[b], nl_out)), (b, nl_out)
    sc.Scale(p1, p1)  # in place (ring_test.go:614)
    got = p1.numpy(squeeze=False)
    for b in range(batch):
        assert np.array_equal(got[b], osc.scale(p[b], len(Q)))


def test_simple_scaler_errors(lg):
    N = 1 << 10
    Q, _, _ = orc.gen_moduli(10, [55, 55], [])
    ctx = lg.ring.NewContextWithParams(N, Q)
    sc = lg.ring.NewSimpleScaler(65537, ctx)
    with pytest.raises(lg.LattigpuError):
        sc.Scale(lg.ring.Poly(N, 1, 1), lg.ring.Poly(N, 2, 1))  # input lacks a limb of the context
    with pytest.raises(lg.LattigpuError):
        sc.Scale(lg.ring.Poly(N, 2, 2), lg.ring.Poly(N, 2, 1))  # batch mismatch
    with pytest.raises(lg.LattigpuError):
        lg.ring.NewSimpleScaler(1, ctx)


PN12 = (12, [39, 39], [30], [60, 60])
PN13 = (13, [54, 54, 54], [55], [60, 60, 60])
ALPHA2 = (12, [50, 45, 45, 45, 45], [50, 50], [60] * 5)  # 5 limbs, alpha 2: the last digit is short


class Rig:
    def __init__(self, lg, params, t=65537, seed=9):
        logN, lq, lp, lm = params
        self.N, self.t = 1 << logN, t
        self.Q, self.P, self.QMul = orc.gen_moduli(logN, lq, lp, lm)
        self.nQ = len(self.Q)
        N = self.N
        self.S = orc.BfvScheme(self.Q, self.P, N, t)
        self.oev = orc.BfvEvaluator(self.S.Q, orc.Context(N, self.QMul), self.S.P, t)
        R = lg.ring
        self.cQ, self.cP, self.cM = R.NewContextWithParams(N, self.Q), R.NewContextWithParams(N, self.P), R.NewContextWithParams(N, self.QMul)
        self.kg = lg.bfv_scheme.KeyGenerator(self.cQ, self.cP)
        self.enc = lg.bfv_scheme.Encoder(self.cQ, t)
        self.ev = lg.bfv.NewEvaluator(self.cQ, self.cM, self.cP, t)
        self.rng = np.random.default_rng(seed)

    def tern(self, *shape):
        return self.rng.integers(-1, 2, size=shape + (self.N,))

    def gauss(self, *shape):
        return np.rint(self.rng.normal(0, 3.2, size=shape + (self.N,))).astype(np.int64)

    def unif(self, mods, batch=None):
        if batch is None:
            return np.ascontiguousarray(np.stack([self.rng.integers(0, q, size=self.N, dtype=np.uint64) for q in mods]))
        return np.ascontiguousarray(np.stack([self.rng.integers(0, q, size=(batch, self.N), dtype=np.uint64) for q in mods], axis=1))

    def keyset(self):
        return [self.gauss() for _ in range(self.S.beta)], [self.unif(self.Q + self.P) for _ in range(self.S.beta)]


def host(ct):
    return np.stack([p.numpy(squeeze=False) for p in ct], axis=1)


@pytest.mark.parametrize("params", [PN12, PN13, ALPHA2], ids=["PN12", "PN13", "alpha2"])
def test_bfv_keygen_encrypt_decrypt_encode(lg, params):
    r = Rig(lg, params)
    S, N, nQ, t = r.S, r.N, r.nQ, r.t
    R = lg.ring
    # keys
    sk_c = r.tern()
    sk, osk = r.kg.GenSecretKey(sk_c), S.gen_secret_key(sk_c)
    assert np.array_equal(sk.numpy(), osk)
    e, a = r.gauss(), r.unif(r.Q + r.P)
    pk, opk = r.kg.GenPublicKey(sk, e, a), S.gen_public_key(osk, e, a)
    assert np.array_equal(pk[0].numpy(), opk[0]) and np.array_equal(pk[1].numpy(), opk[1])
    errs, unis = r.keyset()
    _, rlk_host = r.kg.GenRelinKey(sk, errs, unis)
    assert np.array_equal(rlk_host, S.gen_relin_key(osk, errs, unis))
    errs, unis = r.keyset()
    gen = pow(5, 2, 2 * N)
    _, rot_host = r.kg.genrotkey(sk, gen, errs, unis)
    assert np.array_equal(rot_host, S.gen_rot_key(osk, gen, errs, unis))
    sk2_c = r.tern()
    sk2, osk2 = r.kg.GenSecretKey(sk2_c), S.gen_secret_key(sk2_c)
    errs, unis = r.keyset()
    _, swk_host = r.kg.GenSwitchingKey(sk, sk2, errs, unis)
    assert np.array_equal(swk_host, S.gen_switching_key(osk, osk2, errs, unis))
    assert np.array_equal(sk.numpy(), osk)  # GenSwitchingKey leaves skIn alone (:254)

    # encoder: lift parameters, encode (unsigned, signed, short), decode
    assert [int(x) for x in r.enc.deltaMont()] == S.delta_mont
    assert np.array_equal(r.enc.indexMatrix, S.index_matrix)
    batch = 2
    m = r.rng.integers(0, t, size=(batch, N), dtype=np.uint64)
    pt = R.Poly(N, nQ, batch)
    r.enc.EncodeUint(m, pt)
    opt = np.stack([S.encode_uint(m[b]) for b in range(batch)])
    assert np.array_equal(pt.numpy(squeeze=False), opt)
    assert np.array_equal(r.enc.DecodeUint(pt), m)
    mi = r.rng.integers(-(t // 2), t // 2 + 1, size=(batch, N))
    pti = R.Poly(N, nQ, batch)
    r.enc.EncodeInt(mi, pti)
    assert np.array_equal(pti.numpy(squeeze=False), np.stack([S.encode_int(mi[b]) for b in range(batch)]))
    assert np.array_equal(r.enc.DecodeInt(pti), mi)
    pts = R.Poly(N, nQ, batch)
    r.enc.EncodeUint(m[:, :100], pts)
    assert np.array_equal(pts.numpy(squeeze=False), np.stack([S.encode_uint(m[b, :100]) for b in range(batch)]))

    # encryption, every path
    E = lg.bfv_scheme.Encryptor(r.cQ, r.cP, r.kg.contextQP, pk=pk, sk=sk)
    D = lg.bfv_scheme.Decryptor(r.cQ, sk)
    new_ct = lambda: (R.Poly(N, nQ, batch), R.Poly(N, nQ, batch))
    u, e0, e1 = r.tern(batch), r.gauss(batch), r.gauss(batch)
    for fast in (False, True):
        ct = new_ct()
        E.EncryptPk(pt, ct, u, e0, e1, fast=fast)
        got = host(ct)
        for b in range(batch):
            assert np.array_equal(got[b], S.encrypt_pk(opt[b], opk, u[b], e0[b], e1[b], fast=fast)), ("pk", fast, b)
        if not fast:
            dec = R.Poly(N, nQ, batch)
            D.Decrypt(ct, dec)
            for b in range(batch):
                assert np.array_equal(dec.numpy(squeeze=False)[b], S.decrypt(got[b], osk))
            assert np.array_equal(r.enc.DecodeUint(dec), m)
    for fast in (False, True):
        crp = r.unif(r.Q if fast else r.Q + r.P, batch)
        ct = new_ct()
        E.EncryptSk(pt, ct, crp, e0, fast=fast)
        got = host(ct)
        for b in range(batch):
            assert np.array_equal(got[b], S.encrypt_sk(opt[b], osk, crp[b], e0[b], fast=fast)), ("sk", fast, b)
        dec = R.Poly(N, nQ, batch)
        D.Decrypt(ct, dec)
        assert np.array_equal(r.enc.DecodeUint(dec), m)


@pytest.mark.parametrize("params", [PN12, PN13], ids=["PN12", "PN13"])
def test_bfv_pipeline_config3_shape(lg, params):
    """encode -> encrypt(pk / sk) -> Mul -> Relinearize -> RotateColumns(k) -> decrypt -> decode: every stage
    bit-exact against the oracle, and the decoded slots equal the rotated slot-wise product"""
    r = Rig(lg, params, seed=13)
    S, N, nQ, t = r.S, r.N, r.nQ, r.t
    R = lg.ring
    batch, k = 2, 3
    sk_c = r.tern()
    sk, osk = r.kg.GenSecretKey(sk_c), S.gen_secret_key(sk_c)
    e, a = r.gauss(), r.unif(r.Q + r.P)
    pk, opk = r.kg.GenPublicKey(sk, e, a), S.gen_public_key(osk, e, a)
    errs, unis = r.keyset()
    rlk, rlk_host = r.kg.GenRelinKey(sk, errs, unis)
    gen = pow(5, k, 2 * N)
    errs, unis = r.keyset()
    rot, rot_host = r.kg.genrotkey(sk, gen, errs, unis)
    E = lg.bfv_scheme.Encryptor(r.cQ, r.cP, r.kg.contextQP, pk=pk, sk=sk)
    D = lg.bfv_scheme.Decryptor(r.cQ, sk)
    m0 = r.rng.integers(0, t, size=(batch, N), dtype=np.uint64)
    m1 = r.rng.integers(0, t, size=(batch, N), dtype=np.uint64)
    pt0, pt1 = R.Poly(N, nQ, batch), R.Poly(N, nQ, batch)
    r.enc.EncodeUint(m0, pt0)
    r.enc.EncodeUint(m1, pt1)
    ct0, ct1 = (R.Poly(N, nQ, batch), R.Poly(N, nQ, batch)), (R.Poly(N, nQ, batch), R.Poly(N, nQ, batch))
    u, e0, e1, crp = r.tern(batch), r.gauss(batch), r.gauss(batch), r.unif(r.Q + r.P, batch)
    E.EncryptPk(pt0, ct0, u, e0, e1)
    E.EncryptSk(pt1, ct1, crp, e0)
    ct2 = tuple(R.Poly(N, nQ, batch) for _ in range(3))
    r.ev.Mul(ct0, ct1, ct2)
    ctr = (R.Poly(N, nQ, batch), R.Poly(N, nQ, batch))
    r.ev.Relinearize(ct2, rlk, ctr)
    cto = (R.Poly(N, nQ, batch), R.Poly(N, nQ, batch))
    r.ev.permute(ctr, gen, rot, cto)
    dec = R.Poly(N, nQ, batch)
    D.Decrypt(cto, dec)
    slots = r.enc.DecodeUint(dec)
    row = N // 2
    for b in range(batch):
        o0 = S.encrypt_pk(S.encode_uint(m0[b]), opk, u[b], e0[b], e1[b])
        o1 = S.encrypt_sk(S.encode_uint(m1[b]), osk, crp[b], e0[b])
        o2 = r.oev.tensor_and_rescale(np.ascontiguousarray(o0), np.ascontiguousarray(o1))
        assert np.array_equal(host(ct2)[b], o2)
        orl = r.oev.relinearize(np.ascontiguousarray(o2), rlk_host)
        oro = r.oev.permute(np.ascontiguousarray(orl), gen, rot_host)
        assert np.array_equal(host(cto)[b], oro)
        odec = S.decrypt(oro, osk)
        assert np.array_equal(dec.numpy(squeeze=False)[b], odec)
        assert np.array_equal(slots[b], S.decode_uint(odec))
        prod = (m0[b].astype(object) * m1[b].astype(object)) % t
        want = [prod[(i + k) % row] for i in range(row)] + [prod[row + (i + k) % row] for i in range(row)]
        assert [int(x) for x in slots[b]] == want
