"""GPU parity tests for the BFV evaluator hot path (bfv/evaluator.go:278-813) against the CPU
oracle: Mul (tensorAndRescale, incl. squaring), switchKeys, Relinearize, SwitchKeys and permute
(RotateColumns / RotateRows with a direct key).  Shapes follow bfv/params.go DefaultParams
(PN12, PN13, PN14 bit-exact against the oracle; PN15 = config 3 through properties in
test_gpu_fullsize.py).  BFV ciphertexts are in the coefficient domain."""
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


class Setup:
    def __init__(self, lg, pid):
        p = lg.bfv.DefaultParams[pid]
        self.N = 1 << p["LogN"]
        self.t = p["T"]
        self.Q, self.P, self.QMul = lg.bfv.GenModuli(p)
        oq = orc.gen_moduli(p["LogN"], p["LogQi"], p["LogPi"], p["LogQiMul"])
        assert [self.Q, self.P, self.QMul] == oq
        self.nQ, self.nP = len(self.Q), len(self.P)
        self.beta = -(-self.nQ // self.nP)
        self.oev = orc.BfvEvaluator(orc.Context(self.N, self.Q), orc.Context(self.N, self.QMul), orc.Context(self.N, self.P), self.t)
        self.cQ = lg.ring.NewContextWithParams(self.N, self.Q)
        self.cM = lg.ring.NewContextWithParams(self.N, self.QMul)
        self.cP = lg.ring.NewContextWithParams(self.N, self.P)
        self.ev = lg.bfv.NewEvaluator(self.cQ, self.cM, self.cP, self.t)

    def evk(self, rng):
        k = np.stack([rng.integers(0, q, size=(self.beta, 2, self.N), dtype=np.uint64) for q in self.Q + self.P], axis=2)
        return np.ascontiguousarray(k)

    def ct(self, rng, kind, batch, deg=1):
        if kind == "words":
            return rng.integers(0, 1 << 64, size=(batch, deg + 1, self.nQ, self.N), dtype=np.uint64)
        return np.ascontiguousarray(np.stack(
            [rng.integers(0, q, size=(batch, deg + 1, self.N), dtype=np.uint64) for q in self.Q], axis=2))


def polys(lg, ct):
    return tuple(lg.ring.Poly.from_numpy(np.ascontiguousarray(ct[:, i])) for i in range(ct.shape[1]))


def host(ct):
    return np.stack([p.numpy(squeeze=False) for p in ct], axis=1)


@pytest.mark.parametrize("pid", [0, 1, 2], ids=["PN12", "PN13", "PN14"])
@pytest.mark.parametrize("kind", ["reduced", "words"])
def test_bfv_mul(lg, pid, kind):
    s = Setup(lg, pid)
    rng = np.random.default_rng(41 + pid)
    batch = 2
    a, b = s.ct(rng, kind, batch), s.ct(rng, kind, batch)
    pa, pb = polys(lg, a), polys(lg, b)
    out = tuple(lg.ring.Poly(s.N, s.nQ, batch) for _ in range(3))
    s.ev.Mul(pa, pb, out)
    got = host(out)
    for i in range(batch):
        assert np.array_equal(got[i], s.oev.tensor_and_rescale(np.ascontiguousarray(a[i]), np.ascontiguousarray(b[i]))), i
    s.ev.Mul(pa, pa, out)  # squaring branch (bfv/evaluator.go:334-349)
    got = host(out)
    for i in range(batch):
        x = np.ascontiguousarray(a[i])
        assert np.array_equal(got[i], s.oev.tensor_and_rescale(x, x)), i


@pytest.mark.parametrize("pid", [0, 1, 2], ids=["PN12", "PN13", "PN14"])
@pytest.mark.parametrize("kind", ["reduced", "words"])
def test_bfv_keyswitch_relin_rotate(lg, pid, kind):
    s = Setup(lg, pid)
    rng = np.random.default_rng(51 + pid)
    evk = s.evk(rng)
    key = lg.ckks.SwitchingKey(evk)
    batch = 2
    c = s.ct(rng, kind, batch, deg=2)
    pc = polys(lg, c)
    p0, p1 = lg.ring.Poly(s.N, s.nQ, batch), lg.ring.Poly(s.N, s.nQ, batch)
    s.ev.switchKeys(pc[2], key, p0, p1)
    for i in range(batch):
        w0, w1 = s.oev.switch_keys_core(np.ascontiguousarray(c[i, 2]), evk)
        assert np.array_equal(p0.numpy(squeeze=False)[i], w0) and np.array_equal(p1.numpy(squeeze=False)[i], w1)
    out = (lg.ring.Poly(s.N, s.nQ, batch), lg.ring.Poly(s.N, s.nQ, batch))
    s.ev.Relinearize(pc, key, out)
    for i in range(batch):
        assert np.array_equal(host(out)[i], s.oev.relinearize(np.ascontiguousarray(c[i]), evk))
    s.ev.SwitchKeys(pc[:2], key, out)
    for i in range(batch):
        assert np.array_equal(host(out)[i], s.oev.switch_keys(np.ascontiguousarray(c[i, :2]), evk))
    for gen in (pow(lg.bfv.GaloisGen, 1, 2 * s.N), pow(lg.bfv.GaloisGen, 3, 2 * s.N), 2 * s.N - 1):  # columns by 1, 3; rows
        s.ev.permute(pc[:2], gen, key, out)
        for i in range(batch):
            assert np.array_equal(host(out)[i], s.oev.permute(np.ascontiguousarray(c[i, :2]), gen, evk)), gen
    # in place relinearisation (ctOut == ct0)
    s.ev.Relinearize(pc, key, pc[:2])
    for i in range(batch):
        assert np.array_equal(host(pc[:2])[i], s.oev.relinearize(np.ascontiguousarray(c[i]), evk))


@pytest.mark.parametrize("pid", [0, 1], ids=["PN12", "PN13"])
def test_bfv_general_degree_mul_relinearize(lg, pid):
    """tensorAndRescale for operands that are not both of degree 1 (bfv/evaluator.go:374-417): ciphertext x plaintext,
    degree 2 x degree 1, squaring of a degree-2 ciphertext; Relinearize of a degree-3 ciphertext with two keys
    (:480-507, evakey[deg-2])."""
    s = Setup(lg, pid)
    rng = np.random.default_rng(61 + pid)
    batch = 2
    c1 = s.ct(rng, "reduced", batch, deg=1)
    c2 = s.ct(rng, "reduced", batch, deg=2)
    pt = s.ct(rng, "reduced", batch, deg=0)
    new = lambda n: tuple(lg.ring.Poly(s.N, s.nQ, batch) for _ in range(n))
    for x, y, square in ((c1, pt, False), (c2, c1, False), (c2, c2, True), (c1, c2, False)):
        px = polys(lg, x)
        py = px if square else polys(lg, y)
        out = new(x.shape[1] + y.shape[1] - 1)
        s.ev.Mul(px, py, out)
        got = host(out)
        for i in range(batch):
            want = s.oev.tensor_and_rescale_general(np.ascontiguousarray(x[i]), np.ascontiguousarray(y[i]), square)
            assert np.array_equal(got[i], want), (x.shape, y.shape, square)
    with pytest.raises(ValueError):
        s.ev.Mul(polys(lg, c2), polys(lg, c1), new(3))
    # the general path on two degree-1 operands equals the fused one's oracle (same values, different schedule)
    g = s.oev.tensor_and_rescale_general(np.ascontiguousarray(c1[0]), np.ascontiguousarray(c1[1]))
    assert np.array_equal(g, s.oev.tensor_and_rescale(np.ascontiguousarray(c1[0]), np.ascontiguousarray(c1[1])))
    # Relinearize, degree 3 -> 1 with evakey[1] (degree 3) and evakey[0] (degree 2)
    c3 = s.ct(rng, "reduced", batch, deg=3)
    evks = [s.evk(rng), s.evk(rng)]
    keys = [lg.ckks.SwitchingKey(k) for k in evks]
    out = new(2)
    s.ev.Relinearize(polys(lg, c3), keys, out)
    for i in range(batch):
        assert np.array_equal(host(out)[i], s.oev.relinearize_general(np.ascontiguousarray(c3[i]), evks))
    with pytest.raises(ValueError):
        s.ev.Relinearize(polys(lg, c3), keys[:1], out)
    # degree 1 in: copy (:521-524)
    s.ev.Relinearize(polys(lg, c1), keys, out)
    assert np.array_equal(host(out), c1)


def test_bfv_rotate_columns_pow2_rotate_rows(lg):
    """RotateColumns with a direct key and its power-of-two fallback over the left / right keys (bfv/evaluator.go:578-666),
    RotateRows (:669-680)."""
    s = Setup(lg, 0)
    rng = np.random.default_rng(71)
    N, half = s.N, s.N >> 1
    rk = lg.bfv.RotationKeys()
    left, right = {}, {}
    n = 1
    while n < half:
        kl, kr = s.evk(rng), s.evk(rng)
        rk.evakeyRotColLeft[n], rk.evakeyRotColRight[n] = lg.ckks.SwitchingKey(kl), lg.ckks.SwitchingKey(kr)
        left[n], right[n] = kl, kr
        n <<= 1
    kd, krow = s.evk(rng), s.evk(rng)
    rk.evakeyRotColLeft[3] = lg.ckks.SwitchingKey(kd)
    left[3] = kd
    rk.evakeyRotRow = lg.ckks.SwitchingKey(krow)
    batch = 2
    c = s.ct(rng, "reduced", batch)
    pc = polys(lg, c)
    for k in (3, 5, half - 1, half - 3, 0, half + 2):
        out = (lg.ring.Poly(N, s.nQ, batch), lg.ring.Poly(N, s.nQ, batch))
        s.ev.RotateColumns(pc, k, rk, out)
        for i in range(batch):
            assert np.array_equal(host(out)[i], s.oev.rotate_columns(np.ascontiguousarray(c[i]), k, left, right)), k
    out = (lg.ring.Poly(N, s.nQ, batch), lg.ring.Poly(N, s.nQ, batch))
    s.ev.RotateRows(pc, rk, out)
    for i in range(batch):
        assert np.array_equal(host(out)[i], s.oev.permute(np.ascontiguousarray(c[i]), 2 * N - 1, krow))
    with pytest.raises(ValueError):
        s.ev.RotateColumns(pc, 5, lg.bfv.RotationKeys(), out)
    with pytest.raises(ValueError):
        s.ev.RotateRows(pc, lg.bfv.RotationKeys(), out)
