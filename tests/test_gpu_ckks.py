"""GPU parity tests for the CKKS evaluator key-switch path (ckks/evaluator.go
:933-1591) against the CPU oracle: switchKeysInPlace, MulRelin (+ squaring
branch), Relinearize, Rescale, SwitchKeys and RotateColumns/Conjugate via
permuteNTT.  Shapes follow ckks/params.go DefaultParams (PN12..PN14 bit-exact
against the oracle; the largest sets through size-independent properties in
test_gpu_fullsize.py).  Inputs are NewCiphertextRandom-style (ckks/ciphertext.go
:33-49: full 64-bit words) as in the reference's benchmarks, and in-range.
"""
import numpy as np
import pytest

from oracle import ring_oracle as orc

pytestmark = pytest.mark.gpu

# (logN, LogQi, LogPi) -- ckks/params.go:36-87
PN12 = (12, [37, 32], [38])
PN13 = (13, [33, 30, 30, 30, 30, 30], [35])
PN14 = (14, [45] + [34] * 9, [43, 43])
SMALL3 = (12, [50, 40, 40, 40, 40, 40, 40], [50, 50, 50])  # alpha=3, beta=3 with a partial last digit
# the headline set's digit shape (ckks/params.go:79-86: 34 + 4 limbs, alpha = 4, beta = 9) on a ring the oracle walks in
# seconds: every level exercises one of SURVEY.md Appendix A's digit cases (level 33: digit 8 has two limbs, index 0;
# level 32: digit 8 is a broadcast copy; level 30: digit 7 has three limbs, index 1; ...) through the fused digit loop,
# its own-limb selection and the 96-bit accumulators (beta = 9 terms of 45-bit primes)
ALPHA4 = (12, [55] + [45] * 33, [55] * 4)
WIDE = (12, [60, 59, 59, 59], [60, 60])  # q >= 2^56: the [0,8q) butterflies and their accumulator forms


@pytest.fixture(scope="module")
def lg():
    import lattigpu
    from lattigpu import ring

    ring.set_device(0)
    return lattigpu


class Setup:
    def __init__(self, lg, params):
        logN, lq, lp = params
        self.N = 1 << logN
        self.Q, self.P, _ = orc.gen_moduli(logN, lq, lp)
        self.nQ, self.nP = len(self.Q), len(self.P)
        self.oQ, self.oP = orc.Context(self.N, self.Q), orc.Context(self.N, self.P)
        self.oev = orc.CkksEvaluator(self.oQ, self.oP)
        self.cQ = lg.ring.NewContextWithParams(self.N, self.Q)
        self.cP = lg.ring.NewContextWithParams(self.N, self.P)
        self.ev = lg.ckks.NewEvaluator(self.cQ, self.cP)
        self.beta = -(-self.nQ // self.nP)

    def evk(self, rng):
        """uniform key over QP in [0,q): statistically what newSwitchingKey produces (keygen.go:299-335)"""
        k = np.stack([rng.integers(0, q, size=(self.beta, 2, self.N), dtype=np.uint64) for q in self.Q + self.P], axis=2)
        return np.ascontiguousarray(k)

    def ct(self, rng, kind, batch):
        if kind == "words":
            return rng.integers(0, 1 << 64, size=(batch, 2, self.nQ, self.N), dtype=np.uint64)
        return np.ascontiguousarray(np.stack(
            [rng.integers(0, q, size=(batch, 2, self.N), dtype=np.uint64) for q in self.Q], axis=2))


def polys(lg, ct):
    """[batch][2][nl][N] -> (value0, value1) device polys"""
    return (lg.ring.Poly.from_numpy(np.ascontiguousarray(ct[:, 0])), lg.ring.Poly.from_numpy(np.ascontiguousarray(ct[:, 1])))


def new_ct(lg, s, batch):
    return (lg.ring.Poly(s.N, s.nQ, batch), lg.ring.Poly(s.N, s.nQ, batch))


def host(ct, nl):
    return np.stack([ct[0].numpy(nl=nl, squeeze=False), ct[1].numpy(nl=nl, squeeze=False)], axis=1)


@pytest.mark.parametrize("params", [PN12, PN13, SMALL3, PN14, ALPHA4, WIDE], ids=["PN12", "PN13", "alpha3", "PN14", "alpha4", "wide"])
@pytest.mark.parametrize("kind", ["reduced", "words"])
def test_switch_keys_in_place_all_levels(lg, params, kind):
    s = Setup(lg, params)
    rng = np.random.default_rng(21)
    evk = s.evk(rng)
    dk = lg.ckks.SwitchingKey(evk)
    cx = s.ct(rng, kind, 2)[:, 0]
    pcx = lg.ring.Poly.from_numpy(np.ascontiguousarray(cx))
    levels = range(s.nQ - 1, -1, -1) if params is not PN14 else [9, 8, 5, 0]
    for level in levels:
        p0, p1 = lg.ring.Poly(s.N, s.nQ, 2), lg.ring.Poly(s.N, s.nQ, 2)
        s.ev.switchKeysInPlace(level, pcx, dk, p0, p1)
        g0, g1 = p0.numpy(nl=level + 1), p1.numpy(nl=level + 1)
        for b in range(2):
            w0, w1 = s.oev.switch_keys_in_place(level, np.ascontiguousarray(cx[b]), evk)
            assert np.array_equal(g0[b], w0) and np.array_equal(g1[b], w1), level


@pytest.mark.parametrize("params", [PN13, PN14, SMALL3, WIDE], ids=["PN13", "PN14", "alpha3", "wide"])
@pytest.mark.parametrize("keykind", ["words", "one-word"])
def test_switch_keys_unreduced_key_words(lg, params, keykind):
    """MRed is total (modular_reduction.go:70-79): a switching key holding arbitrary 64-bit words still has a defined
    result in the reference.  The lazy key-switch accumulators (96-bit and 64-bit) assume key words of at most bits(q)
    bits; a CTA that meets a wider word must repeat its tile on the exact path (canonical digit, MRed + CRed per term) and
    still match -- on every butterfly class (45-bit, 50-bit and 60-bit limbs) -- as must the "ks_acc64" switch."""
    s = Setup(lg, params)
    rng = np.random.default_rng(27)
    evk = s.evk(rng)
    if keykind == "words":
        evk = rng.integers(0, 1 << 64, size=evk.shape, dtype=np.uint64)
    else:  # a single out-of-range word in one tile of one limb of one digit
        evk[s.beta - 1, 1, 1, 2049] = np.uint64((1 << 64) - 3)
        evk[0, 0, s.nQ, 5] = np.uint64(1 << 62)
    cx = s.ct(rng, "reduced", 2)[:, 0]
    pcx = lg.ring.Poly.from_numpy(np.ascontiguousarray(cx))
    try:
        for acc64 in (0, 1):
            lg.ring.debug_set_switch("ks_acc64", acc64)
            dk = lg.ckks.SwitchingKey(evk)
            for level in (s.nQ - 1, s.nQ - 2):
                p0, p1 = lg.ring.Poly(s.N, s.nQ, 2), lg.ring.Poly(s.N, s.nQ, 2)
                s.ev.switchKeysInPlace(level, pcx, dk, p0, p1)
                g0, g1 = p0.numpy(nl=level + 1), p1.numpy(nl=level + 1)
                for b in range(2):
                    w0, w1 = s.oev.switch_keys_in_place(level, np.ascontiguousarray(cx[b]), evk)
                    assert np.array_equal(g0[b], w0) and np.array_equal(g1[b], w1), (level, acc64)
    finally:
        lg.ring.debug_set_switch("ks_acc64", 0)


@pytest.mark.parametrize("batch", [1, 2, 3, 5])
def test_digit_loop_kernels_agree(lg, batch):
    """The fused digit loop runs as two launches: ks_fused_tma_kernel (two batch entries per CTA, key tiles by TMA) on the
    FP64-class limbs whose key words are canonical, ks_fused_kernel on the rest.  Both must give the oracle's words for odd
    and even batches (an odd batch leaves half of the last CTA idle, a single entry stays on ks_fused_kernel), with the
    "no_ks_tma" switch forcing the old kernel everywhere, and with a key whose one non-canonical word moves exactly one limb
    from one kernel to the other."""
    s = Setup(lg, ALPHA4)
    rng = np.random.default_rng(31 + batch)
    evk = s.evk(rng)
    evk_bad = evk.copy()
    evk_bad[2, 1, 3, 7] = np.uint64((1 << 64) - 5)  # limb 3 (45-bit): not canonical -> integer accumulators for that limb
    cx = s.ct(rng, "reduced", batch)[:, 0]
    pcx = lg.ring.Poly.from_numpy(np.ascontiguousarray(cx))
    try:
        for key in (evk, evk_bad):
            results = []
            for off in (0, 1):
                lg.ring.debug_set_switch("no_ks_tma", off)
                dk = lg.ckks.SwitchingKey(key)
                for level in (s.nQ - 1, s.nQ - 3):
                    p0, p1 = lg.ring.Poly(s.N, s.nQ, batch), lg.ring.Poly(s.N, s.nQ, batch)
                    s.ev.switchKeysInPlace(level, pcx, dk, p0, p1)
                    results.append((p0.numpy(nl=level + 1, squeeze=False), p1.numpy(nl=level + 1, squeeze=False)))
            half = len(results) // 2
            for (a0, a1), (b0, b1) in zip(results[:half], results[half:]):
                assert np.array_equal(a0, b0) and np.array_equal(a1, b1)
            for k, level in enumerate((s.nQ - 1, s.nQ - 3)):
                for b in range(batch):
                    w0, w1 = s.oev.switch_keys_in_place(level, np.ascontiguousarray(cx[b]), key)
                    assert np.array_equal(results[k][0][b], w0) and np.array_equal(results[k][1][b], w1), (level, b)
    finally:
        lg.ring.debug_set_switch("no_ks_tma", 0)


@pytest.mark.parametrize("params", [PN12, PN13, SMALL3, PN14], ids=["PN12", "PN13", "alpha3", "PN14"])
@pytest.mark.parametrize("kind", ["reduced", "words"])
def test_mul_relin_rescale(lg, params, kind):
    """the north-star sequence: MulRelin (ckks/evaluator.go:1016) then Rescale (:933)"""
    s = Setup(lg, params)
    rng = np.random.default_rng(22)
    evk = s.evk(rng)
    rlk = lg.ckks.SwitchingKey(evk)
    batch = 3
    a, b = s.ct(rng, kind, batch), s.ct(rng, kind, batch)
    for level in sorted({s.nQ - 1, max(s.nQ - 2, 1)}, reverse=True):
        nl = level + 1
        pa, pb, out = polys(lg, a), polys(lg, b), new_ct(lg, s, batch)
        s.ev.MulRelin(level, pa, pb, rlk, out)
        got = host(out, nl)
        want = np.stack([s.oev.mul_relin(level, np.ascontiguousarray(a[i, :, :nl]), np.ascontiguousarray(b[i, :, :nl]), evk)
                         for i in range(batch)])
        assert np.array_equal(got, want), level
        if level >= 1:
            s.ev.Rescale(nl, out)
            got = host(out, nl - 1)
            wr = np.stack([s.oev.rescale(want[i]) for i in range(batch)])
            assert np.array_equal(got, wr), level
    # squaring branch (el0 == el1, :1080-1085) and output aliasing an input
    level = s.nQ - 1
    pa = polys(lg, a)
    s.ev.MulRelin(level, pa, pa, rlk, pa)
    for i in range(batch):
        x = np.ascontiguousarray(a[i])
        assert np.array_equal(host(pa, s.nQ)[i], s.oev.mul_relin(level, x, x, evk))
    if s.nQ >= 3:  # RescaleMany-style double drop
        out = polys(lg, a)
        s.ev.Rescale(s.nQ, out, nb=2)
        for i in range(batch):
            assert np.array_equal(host(out, s.nQ - 2)[i], s.oev.rescale(np.ascontiguousarray(a[i]), nb=2))


@pytest.mark.parametrize("kind", ["reduced", "words"])
def test_mul_relin_rescale_alpha4_all_levels(lg, kind):
    """MulRelin -> Rescale at EVERY level 33..1 of the headline digit shape (34 + 4 limbs, alpha 4, beta 9; N = 2^12)."""
    s = Setup(lg, ALPHA4)
    rng = np.random.default_rng(31)
    evk = s.evk(rng)
    rlk = lg.ckks.SwitchingKey(evk)
    batch = 2
    a, b = s.ct(rng, kind, batch), s.ct(rng, kind, batch)
    for level in range(s.nQ - 1, 0, -1):
        nl = level + 1
        pa, pb, out = polys(lg, a), polys(lg, b), new_ct(lg, s, batch)
        s.ev.MulRelin(level, pa, pb, rlk, out)
        got = host(out, nl)
        want = np.stack([s.oev.mul_relin(level, np.ascontiguousarray(a[i, :, :nl]), np.ascontiguousarray(b[i, :, :nl]), evk)
                         for i in range(batch)])
        assert np.array_equal(got, want), level
        s.ev.Rescale(nl, out)
        got = host(out, nl - 1)
        wr = np.stack([s.oev.rescale(want[i]) for i in range(batch)])
        assert np.array_equal(got, wr), level


def test_mul_relin_without_key_and_with_plaintext(lg):
    """The other branches of MulRelin: evakey == nil leaves a degree-2 ciphertext (:1059-1063, :1111-1117) that
    Relinearize (:1144-1162) brings to the keyed result; plaintext x ciphertext in either order (:1121-1137)."""
    s = Setup(lg, PN13)
    rng = np.random.default_rng(29)
    evk = s.evk(rng)
    rlk = lg.ckks.SwitchingKey(evk)
    batch, level = 2, s.nQ - 2
    nl = level + 1
    for kind in ("reduced", "words"):
        a, b = s.ct(rng, kind, batch), s.ct(rng, kind, batch)
        oq = s.oQ
        mf = lambda x: oq.op2("mform_poly", np.ascontiguousarray(x[:nl]), nl=nl)
        mul = lambda x, y: oq.op3("mulcoeffs_montgomery", x, np.ascontiguousarray(y[:nl]), nl=nl)
        for sq in (False, True):
            pa = polys(lg, a)
            pb = pa if sq else polys(lg, b)
            out = tuple(lg.ring.Poly(s.N, s.nQ, batch) for _ in range(3))
            s.ev.MulRelin(level, pa, pb, None, out)
            bb = a if sq else b
            for i in range(batch):
                c00, c01 = mf(a[i, 0]), mf(a[i, 1])
                w0 = mul(c00, bb[i, 0])
                if sq:
                    w1 = mul(c00, bb[i, 1])
                    w1 = oq.op3("add", w1, w1.copy(), nl=nl)
                else:
                    w1 = oq.op3("mulcoeffs_montgomery_and_add", c01, np.ascontiguousarray(bb[i, 0, :nl]), mul(c00, bb[i, 1]), nl=nl)
                w2 = mul(c01, bb[i, 1])
                for got, want in zip(out, (w0, w1, w2)):
                    assert np.array_equal(got.numpy(nl=nl, squeeze=False)[i], want), (kind, sq)
            rel = new_ct(lg, s, batch)
            s.ev.Relinearize(level, out, rlk, rel)
            for i in range(batch):
                x, y = np.ascontiguousarray(a[i, :, :nl]), np.ascontiguousarray(bb[i, :, :nl])
                assert np.array_equal(host(rel, nl)[i], s.oev.mul_relin(level, x, y, evk)), (kind, sq)
        # receiver aliasing an input: the three products go through the pool and are copied (:1111-1116)
        pa, pb = polys(lg, a), polys(lg, b)
        out = (pa[0], pa[1], lg.ring.Poly(s.N, s.nQ, batch))
        s.ev.MulRelin(level, pa, pb, None, out)
        i = 1
        assert np.array_equal(out[0].numpy(nl=nl, squeeze=False)[i], mul(mf(a[i, 0]), b[i, 0]))
        assert np.array_equal(out[2].numpy(nl=nl, squeeze=False)[i], mul(mf(a[i, 1]), b[i, 1]))
        # plaintext (degree 0) x ciphertext, both orders
        pt = a[:, 0]
        ppt = (lg.ring.Poly.from_numpy(np.ascontiguousarray(pt)),)
        for order in (0, 1):
            pb, out = polys(lg, b), new_ct(lg, s, batch)
            if order == 0:
                s.ev.MulRelin(level, ppt, pb, None, out)
            else:
                s.ev.MulRelin(level, pb, ppt, rlk, out)
            for i in range(batch):
                c = mf(pt[i])
                assert np.array_equal(out[0].numpy(nl=nl, squeeze=False)[i], mul(c, b[i, 0])), kind
                assert np.array_equal(out[1].numpy(nl=nl, squeeze=False)[i], mul(c, b[i, 1])), kind


@pytest.mark.parametrize("params", [PN13, SMALL3], ids=["PN13", "alpha3"])
def test_rotate_conjugate_switchkeys_relinearize(lg, params):
    s = Setup(lg, params)
    rng = np.random.default_rng(23)
    evk = s.evk(rng)
    key = lg.ckks.SwitchingKey(evk)
    batch = 2
    a = s.ct(rng, "reduced", batch)
    pa = polys(lg, a)
    GaloisGen = 5  # ckks/ckks.go
    for level in (s.nQ - 1, 1):
        nl = level + 1
        for gen, power in [(GaloisGen, 1), (GaloisGen, 5), (2 * s.N - 1, 1)]:  # rotations by 1, 5; conjugation
            idx = lg.ring.PermuteNTTIndex(gen, power, s.N)
            out = new_ct(lg, s, batch)
            s.ev.permuteNTT(level, pa, idx, key, out)
            widx = orc.permute_ntt_index(gen, power, s.N)
            for i in range(batch):
                want = s.oev.permute_ntt(level, np.ascontiguousarray(a[i, :, :nl]), widx, evk)
                assert np.array_equal(host(out, nl)[i], want), (level, gen, power)
        out = new_ct(lg, s, batch)
        s.ev.SwitchKeys(level, pa, key, out)
        for i in range(batch):
            assert np.array_equal(host(out, nl)[i], s.oev.switch_keys(level, np.ascontiguousarray(a[i, :, :nl]), evk))
        # Relinearize (:1144-1162) of a degree-2 ciphertext = switch keys on value[2]
        c2 = s.ct(rng, "reduced", batch)[:, 0]
        p2 = lg.ring.Poly.from_numpy(np.ascontiguousarray(c2))
        out = new_ct(lg, s, batch)
        s.ev.Relinearize(level, (pa[0], pa[1], p2), key, out)
        for i in range(batch):
            k0, k1 = s.oev.switch_keys_in_place(level, np.ascontiguousarray(c2[i]), evk)
            assert np.array_equal(out[0].numpy(nl=nl, squeeze=False)[i], s.oQ.op3("add", np.ascontiguousarray(a[i, 0, :nl]), k0))
            assert np.array_equal(out[1].numpy(nl=nl, squeeze=False)[i], s.oQ.op3("add", np.ascontiguousarray(a[i, 1, :nl]), k1))


@pytest.mark.parametrize("params", [PN12, SMALL3], ids=["PN12", "alpha3"])
def test_rotate_columns_pow2_conjugate_rescale_many(lg, params):
    """The drivers around permuteNTT: RotateColumns with a direct key, its power-of-two fallback over the left or the
    right keys (ckks/evaluator.go:1201-1248, :1402-1424), Conjugate (:1437-1450), RescaleMany (:971-1000) and the
    threshold loop of Rescale (:933-968).  One (uniform) switching key per rotation, as RotationKeys holds them."""
    s = Setup(lg, params)
    rng = np.random.default_rng(41)
    N, half = s.N, s.N >> 1
    rk = lg.ckks.RotationKeys(N)
    left, right = {}, {}
    n = 1
    while n < half:  # GenRot's power-of-two set (keygen.go:405-413)
        kl, kr = s.evk(rng), s.evk(rng)
        rk.SetPow2(n, lg.ckks.SwitchingKey(kl), lg.ckks.SwitchingKey(kr))
        left[n] = (orc.permute_ntt_index(5, n, N), kl)
        right[n] = (orc.permute_ntt_index(5, 2 * N - n, N), kr)
        n <<= 1
    kd = s.evk(rng)  # a specific rotation: direct key
    rk.SetRotKey(lg.ckks.SwitchingKey(kd), rk.RotationLeft, 3)
    left[3] = (orc.permute_ntt_index(5, 3, N), kd)
    kc = s.evk(rng)
    rk.SetRotKey(lg.ckks.SwitchingKey(kc), rk.Conjugate)
    batch = 2
    a = s.ct(rng, "reduced", batch)
    pa = polys(lg, a)
    level = s.nQ - 1
    nl = level + 1
    # k = 3: direct key; 5: two left rotations; half - 1: one right rotation; 0 and half: copy
    for k in (3, 5, half - 1, half - 3, 0, half, N + 6):
        out = new_ct(lg, s, batch)
        s.ev.RotateColumns(level, pa, k, rk, out)
        for i in range(batch):
            want = s.oev.rotate_columns(level, np.ascontiguousarray(a[i, :, :nl]), k, left, right)
            assert np.array_equal(host(out, nl)[i], want), k
    out = new_ct(lg, s, batch)
    s.ev.Conjugate(level, pa, rk, out)
    cidx = orc.permute_ntt_index(2 * N - 1, 1, N)
    for i in range(batch):
        assert np.array_equal(host(out, nl)[i], s.oev.permute_ntt(level, np.ascontiguousarray(a[i, :, :nl]), cidx, kc))
    with pytest.raises(ValueError):  # neither the rotation nor the power-of-two set
        s.ev.RotateColumns(level, pa, 5, lg.ckks.RotationKeys(N), new_ct(lg, s, batch))
    if s.nQ >= 3:
        out = polys(lg, a)
        div = s.ev.RescaleMany(s.nQ, out, 2)
        assert div == float(s.Q[-1]) * float(s.Q[-2])
        for i in range(batch):
            assert np.array_equal(host(out, s.nQ - 2)[i], s.oev.rescale_many(np.ascontiguousarray(a[i]), 2))
        with pytest.raises(ValueError):
            s.ev.RescaleMany(2, out, 2)
    # Rescale's threshold loop: a scale of about q_top * q_top-1 * 2^10 drops two levels against threshold 2^10
    out = polys(lg, a)
    scale0 = float(s.Q[-1]) * float(s.Q[-2]) * 1024.0 if s.nQ >= 3 else float(s.Q[-1]) * 1024.0
    scale, nl2 = s.ev.RescaleThreshold(s.nQ, out, scale0, 1024.0)
    drops = s.nQ - nl2
    assert drops == (2 if s.nQ >= 3 else 1)
    for i in range(batch):
        assert np.array_equal(host(out, nl2)[i], s.oev.rescale(np.ascontiguousarray(a[i]), nb=drops))


@pytest.mark.parametrize("params", [PN13, SMALL3, PN14], ids=["PN13", "alpha3", "PN14"])
@pytest.mark.parametrize("kind", ["reduced", "words"])
def test_rotate_hoisted(lg, params, kind):
    """RotateHoisted / switchKeyHoisted (ckks/evaluator.go:1252-1392): one decomposition, several rotations,
    each with its own rotation key; bit-exact against the oracle at the top level and at a partial level."""
    s = Setup(lg, params)
    rng = np.random.default_rng(29)
    batch = 2
    a = s.ct(rng, kind, batch)
    pa = polys(lg, a)
    rots = [(5, 1), (5, 3), (5, 64)]
    evks = [s.evk(rng) for _ in rots]
    keys = [lg.ckks.SwitchingKey(k) for k in evks]
    for level in (s.nQ - 1, 2 if s.nQ > 3 else 0):
        nl = level + 1
        idxs = [lg.ring.PermuteNTTIndex(g, pw, s.N) for g, pw in rots]
        widx = [orc.permute_ntt_index(g, pw, s.N) for g, pw in rots]
        outs = [new_ct(lg, s, batch) for _ in rots]
        s.ev.RotateHoisted(level, pa, list(zip(idxs, keys)), outs)
        for i in range(batch):
            want = s.oev.rotate_hoisted(level, np.ascontiguousarray(a[i, :, :nl]), widx, evks)
            for r in range(len(rots)):
                assert np.array_equal(host(outs[r], nl)[i], want[r]), (level, rots[r], i)
    with pytest.raises(lg.LattigpuError, match="not in place"):
        s.ev.RotateHoisted(s.nQ - 1, pa, [(idxs[0], keys[0])], [pa])


@pytest.mark.parametrize("params", [PN13, SMALL3], ids=["PN13", "alpha3"])
def test_const_ops(lg, params):
    """AddConst, MultByConst, MultByConstAndAdd, MultByi, DivByi (ckks/evaluator.go:373-833) bit-exact against
    the oracle, complex / fractional / negative / integer constants, full and partial level."""
    s = Setup(lg, params)
    rng = np.random.default_rng(31)
    batch = 2
    a = s.ct(rng, "reduced", batch)
    acc = s.ct(rng, "reduced", batch)
    pa = polys(lg, a)
    scale = float(1 << 30)
    for level in (s.nQ - 1, 1):
        nl = level + 1
        for const in (3.25 - 1.5j, -2.0, 7, 0.0 + 2.5j, -0.75 + 0.0j):
            c = complex(const)
            out = new_ct(lg, s, batch)
            s.ev.AddConst(level, pa, const, scale, out)
            sc = s.ev.const_scale(const, scale)
            out2 = new_ct(lg, s, batch)
            s.ev.MultByConst(level, pa, const, sc, out2)
            out3 = polys(lg, acc)
            s.ev.MultByConstAndAdd(level, pa, const, sc, out3)
            for i in range(batch):
                ins = [np.ascontiguousarray(a[i, u, :nl]) for u in range(2)]
                w = orc.ckks_const_op(s.oQ, "add", level, ins[:1], ins[:1], c.real, c.imag, scale)
                assert np.array_equal(out[0].numpy(nl=nl, squeeze=False)[i], w[0]), (level, const, "add")
                w = orc.ckks_const_op(s.oQ, "mul", level, ins, ins, c.real, c.imag, sc)
                assert np.array_equal(host(out2, nl)[i], np.stack(w)), (level, const, "mul")
                accs = [np.ascontiguousarray(acc[i, u, :nl]) for u in range(2)]
                w = orc.ckks_const_op(s.oQ, "mul_add", level, ins, accs, c.real, c.imag, sc)
                assert np.array_equal(host(out3, nl)[i], np.stack(w)), (level, const, "mul_add")
        oi, od = new_ct(lg, s, batch), new_ct(lg, s, batch)
        s.ev.MultByi(level, pa, oi)
        s.ev.DivByi(level, pa, od)
        for i in range(batch):
            ins = [np.ascontiguousarray(a[i, u, :nl]) for u in range(2)]
            assert np.array_equal(host(oi, nl)[i], np.stack(orc.ckks_const_op(s.oQ, "mul_i", level, ins, ins)))
            assert np.array_equal(host(od, nl)[i], np.stack(orc.ckks_const_op(s.oQ, "div_i", level, ins, ins)))


def test_error_paths(lg):
    s = Setup(lg, PN12)
    rng = np.random.default_rng(24)
    key = lg.ckks.SwitchingKey(s.evk(rng))
    a = polys(lg, s.ct(rng, "reduced", 1))
    with pytest.raises(lg.LattigpuError, match="level 0"):  # ckks/evaluator.go:938
        s.ev.Rescale(1, a)
    with pytest.raises(lg.LattigpuError, match="out of range"):
        s.ev.MulRelin(5, a, a, key, a)
    short = lg.ckks.SwitchingKey(s.evk(rng)[:1])
    with pytest.raises(lg.LattigpuError, match="digits"):
        s.ev.MulRelin(1, a, a, short, a)


def test_keyswitch_batch_chunking():
    """The digit scratch of a key switch is bounded: large batches are processed in chunks.  Forced here with a
    tiny budget (LATTIGPU_KS_SCRATCH_WORDS, read once per process -> subprocess) so that a batch of 5 takes three
    chunks; results must equal the unchunked ones bit for bit."""
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = r"""
import sys, os
sys.path.insert(0, %r); sys.path.insert(0, os.path.join(%r, "lattigo-fhe-by-go_b200"))
import numpy as np
import lattigpu
from lattigpu import ckks, ring
ring.set_device(0)
p = dict(LogN=12, LogQi=[50, 40, 40, 40, 40, 40, 40], LogPi=[50, 50, 50])
N = 1 << p["LogN"]; Q, P = ckks.GenModuli(p); nQ, nP = len(Q), len(P); beta = -(-nQ // nP)
rng = np.random.default_rng(3)
evk = np.ascontiguousarray(np.stack([rng.integers(0, q, size=(beta, 2, N), dtype=np.uint64) for q in Q + P], axis=2))
B = 5
a = np.ascontiguousarray(np.stack([rng.integers(0, q, size=(B, 2, N), dtype=np.uint64) for q in Q], axis=2))
b = np.ascontiguousarray(np.stack([rng.integers(0, q, size=(B, 2, N), dtype=np.uint64) for q in Q], axis=2))
ev = ckks.NewEvaluator(ring.NewContextWithParams(N, Q), ring.NewContextWithParams(N, P))
key = ckks.SwitchingKey(evk)
F = lambda x: (ring.Poly.from_numpy(np.ascontiguousarray(x[:, 0])), ring.Poly.from_numpy(np.ascontiguousarray(x[:, 1])))
out = (ring.Poly(N, nQ, B), ring.Poly(N, nQ, B))
ev.MulRelin(nQ - 1, F(a), F(b), key, out)
np.save(sys.argv[1], np.stack([out[0].numpy(squeeze=False), out[1].numpy(squeeze=False)]))
""" % (root, root)
    import tempfile

    outs = []
    for budget in (None, str(2 * 3 * 10 * 4096)):  # second run: room for two batch entries per chunk
        env = dict(os.environ)
        if budget:
            env["LATTIGPU_KS_SCRATCH_WORDS"] = budget
        with tempfile.NamedTemporaryFile(suffix=".npy") as f:
            subprocess.run([sys.executable, "-c", code, f.name], check=True, env=env, timeout=300)
            outs.append(np.load(f.name))
    assert np.array_equal(outs[0], outs[1])
