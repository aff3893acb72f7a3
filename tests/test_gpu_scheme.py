"""GPU parity for the CKKS key generator / encryptor / decryptor ring sequences (ckks/keygen.go:96-494,
ckks/encryptor.go:179-362, ckks/decryptor.go:53-78) and for the full BASELINE config-2 pipeline
encrypt -> MulRelin -> Rescale -> decrypt, device-resident, against the oracle restatement with the same
sampled values (sampling is host-side in the reference: crypto/rand)."""
import numpy as np
import pytest

from oracle import ring_oracle as orc

pytestmark = pytest.mark.gpu

PN12 = (12, [37, 32], [38])
SMALL3 = (12, [50, 40, 40, 40, 40, 40, 40], [50, 50, 50])
PN14 = (14, [45] + [34] * 9, [43, 43])


@pytest.fixture(scope="module")
def lg():
    import lattigpu
    from lattigpu import ring

    ring.set_device(0)
    return lattigpu


@pytest.mark.parametrize("params", [PN12, SMALL3, PN14], ids=["PN12", "alpha3", "PN14"])
def test_keygen_encrypt_mulrelin_rescale_decrypt(lg, params):
    logN, lq, lp = params
    N = 1 << logN
    Q, P, _ = orc.gen_moduli(logN, lq, lp)
    nQ = len(Q)
    rng = np.random.default_rng(55)
    S = orc.CkksScheme(Q, P, N)
    cQ, cP = lg.ring.NewContextWithParams(N, Q), lg.ring.NewContextWithParams(N, P)
    kg = lg.ckks_scheme.KeyGenerator(cQ, cP)
    tern = lambda *shape: rng.integers(-1, 2, size=shape + (N,))
    gauss = lambda *shape: np.rint(rng.normal(0, 3.2, size=shape + (N,))).astype(np.int64)
    unif = lambda mods: np.ascontiguousarray(np.stack([rng.integers(0, q, size=N, dtype=np.uint64) for q in mods]))

    # keys
    sk_c = tern()
    sk = kg.GenSecretKey(sk_c)
    osk = S.gen_secret_key(sk_c)
    assert np.array_equal(sk.numpy(), osk)
    e, a = gauss(), unif(Q + P)
    pk = kg.GenPublicKey(sk, e, a)
    opk = S.gen_public_key(osk, e, a)
    assert np.array_equal(pk[0].numpy(), opk[0]) and np.array_equal(pk[1].numpy(), opk[1])
    errs, unis = [gauss() for _ in range(S.beta)], [unif(Q + P) for _ in range(S.beta)]
    rlk, rlk_host = kg.GenRelinKey(sk, errs, unis)
    orlk = S.gen_relin_key(osk, errs, unis)
    assert np.array_equal(rlk_host, orlk)
    errs, unis = [gauss() for _ in range(S.beta)], [unif(Q + P) for _ in range(S.beta)]
    rot, rot_host = kg.genrotKey(sk, 5, errs, unis)
    assert np.array_equal(rot_host, S.gen_rot_key(osk, 5, errs, unis))

    # encryption, every path, top level and one below (the ModDownPQ call is kept literal)
    batch = 2
    enc = lg.ckks_scheme.Encryptor(cQ, cP, kg.contextQP, pk=pk, sk=sk)
    pt = np.ascontiguousarray(np.stack([rng.integers(0, q, size=(batch, N), dtype=np.uint64) for q in Q], axis=1))
    for level in (nQ - 1, nQ - 2):
        nl = level + 1
        ppt = lg.ring.Poly.from_numpy(pt)
        for fast in (False, True):
            u, e0, e1 = tern(batch), gauss(batch), gauss(batch)
            ct = (lg.ring.Poly(N, nQ, batch), lg.ring.Poly(N, nQ, batch))
            enc.EncryptPk(level, ppt, ct, u, e0, e1, fast=fast)
            for b in range(batch):
                want = S.encrypt_pk(level, pt[b], opk, u[b], e0[b], e1[b], fast=fast)
                assert np.array_equal(ct[0].numpy(nl=nl, squeeze=False)[b], want[0]), (level, fast, "pk0")
                assert np.array_equal(ct[1].numpy(nl=nl, squeeze=False)[b], want[1]), (level, fast, "pk1")
            crp = np.ascontiguousarray(np.stack([unif(Q if fast else Q + P) for _ in range(batch)]))
            ee = gauss(batch)
            ct = (lg.ring.Poly(N, nQ, batch), lg.ring.Poly(N, nQ, batch))
            enc.EncryptSk(level, ppt, ct, crp, ee, fast=fast)
            for b in range(batch):
                want = S.encrypt_sk(level, pt[b], osk, crp[b], ee[b], fast=fast)
                assert np.array_equal(ct[0].numpy(nl=nl, squeeze=False)[b], want[0]), (level, fast, "sk0")
                assert np.array_equal(ct[1].numpy(nl=nl, squeeze=False)[b], want[1]), (level, fast, "sk1")

    # config 2: encrypt -> MulRelin -> Rescale -> decrypt, all on the device
    level = nQ - 1
    ev = lg.ckks.NewEvaluator(cQ, cP)
    oev = orc.CkksEvaluator(S.Q, S.P)
    dec = lg.ckks_scheme.Decryptor(cQ, sk)
    pts = [np.ascontiguousarray(np.stack([rng.integers(0, q, size=(batch, N), dtype=np.uint64) for q in Q], axis=1))
           for _ in range(2)]
    rnd = [(tern(batch), gauss(batch), gauss(batch)) for _ in range(2)]
    cts = []
    for ptv, (u, e0, e1) in zip(pts, rnd):
        ct = (lg.ring.Poly(N, nQ, batch), lg.ring.Poly(N, nQ, batch))
        enc.EncryptPk(level, lg.ring.Poly.from_numpy(ptv), ct, u, e0, e1)
        cts.append(ct)
    out = (lg.ring.Poly(N, nQ, batch), lg.ring.Poly(N, nQ, batch))
    ev.MulRelin(level, cts[0], cts[1], rlk, out)
    ev.Rescale(nQ, out)
    ptout = lg.ring.Poly(N, nQ, batch)
    dec.Decrypt(level - 1, out, ptout)
    for b in range(batch):
        oc = [S.encrypt_pk(level, pts[k][b], opk, rnd[k][0][b], rnd[k][1][b], rnd[k][2][b]) for k in range(2)]
        w = oev.rescale(oev.mul_relin(level, np.ascontiguousarray(oc[0]), np.ascontiguousarray(oc[1]), orlk))
        assert np.array_equal(ptout.numpy(nl=level, squeeze=False)[b], S.decrypt(level - 1, w, osk))
