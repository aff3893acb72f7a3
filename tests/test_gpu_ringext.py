"""GPU parity of the ring.Context methods outside the evaluator path (csrc/ringext.cu): MulPoly*, Exp, Shift, Rotate,
Equal -- against the oracle at sizes its Python loops finish in seconds, and through the reference's own property tests
(ring/ring_test.go:422-548: GaloisShift, MulPoly) at its test size and at N = 2^16.  All calls go through the C ABI."""
import numpy as np
import pytest

from oracle import ring_oracle as orc

pytestmark = pytest.mark.gpu

QI60 = [1152921504606584833, 1152921504598720513, 1152921504592429057, 1152921504581419009]  # ring/params.go:50-69
Q45 = [35184372744193, 35184373006337, 35184376545281]  # 45-bit, = 1 mod 2^17: the FP64-butterfly class


@pytest.fixture(scope="module")
def lg():
    import lattigpu
    from lattigpu import ring

    ring.set_device(0)
    return lattigpu


def uniform(rng, moduli, N):
    return np.ascontiguousarray(np.stack([rng.integers(0, q, size=N, dtype=np.uint64) for q in moduli]))


@pytest.mark.parametrize("moduli", [QI60[:2], Q45[:2]])
@pytest.mark.parametrize("N", [16, 64])
def test_mul_poly_family_vs_oracle(lg, N, moduli):
    rng = np.random.default_rng(N)
    o = orc.Context(N, moduli)
    ctx = lg.ring.NewContextWithParams(N, moduli)
    a, b = uniform(rng, moduli, N), uniform(rng, moduli, N)
    pa, pb = lg.ring.Poly.from_numpy(a), lg.ring.Poly.from_numpy(b)
    for name, fn, mont in (("MulPoly", orc.mul_poly, False), ("MulPolyMontgomery", orc.mul_poly, True),
                           ("MulPolyNaive", orc.mul_poly_naive, False), ("MulPolyNaiveMontgomery", orc.mul_poly_naive, True)):
        pc = lg.ring.Poly.from_numpy(rng.integers(0, 1 << 64, size=a.shape, dtype=np.uint64))  # stale words in the receiver
        getattr(ctx, name)(pa, pb, pc)
        assert np.array_equal(pc.numpy(), fn(o, a, b, montgomery=mont)), name
        assert np.array_equal(pa.numpy(), a) and np.array_equal(pb.numpy(), b), name  # operands untouched
    # receiver aliasing an operand (the reference copies p1 and p2 first)
    pc = lg.ring.Poly.from_numpy(a)
    ctx.MulPolyNaive(pc, pb, pc)
    assert np.array_equal(pc.numpy(), orc.mul_poly_naive(o, a, b))


@pytest.mark.parametrize("N", [1 << 13, 1 << 16])
def test_mul_poly_property(lg, N):
    """ring/ring_test.go:503-548 at full size: MulPoly == MulPolyNaive, and InvMForm(MulPolyMontgomery(MForm, MForm)) too"""
    moduli = Q45[:2] if N == 1 << 16 else QI60[:2]
    rng = np.random.default_rng(7)
    ctx = lg.ring.NewContextWithParams(N, moduli)
    a, b = uniform(rng, moduli, N), uniform(rng, moduli, N)
    p1, p2 = lg.ring.Poly.from_numpy(a), lg.ring.Poly.from_numpy(b)
    want, test = ctx.NewPoly(), ctx.NewPoly()
    ctx.MulPolyNaive(p1, p2, want)
    ctx.MulPoly(p1, p2, test)
    assert ctx.Equal(want, test)
    ctx.MForm(p1, p1)
    ctx.MForm(p2, p2)
    ctx.MulPolyMontgomery(p1, p2, test)
    ctx.InvMForm(test, test)
    assert np.array_equal(want.numpy(), test.numpy())


@pytest.mark.parametrize("N", [16, 256])
def test_shift_rotate_exp_equal_vs_oracle(lg, N):
    moduli = QI60[:3]
    rng = np.random.default_rng(N + 1)
    o = orc.Context(N, moduli)
    ctx = lg.ring.NewContextWithParams(N, moduli)
    a = uniform(rng, moduli, N)
    w = rng.integers(0, 1 << 64, size=a.shape, dtype=np.uint64)  # unreduced words
    for n in (0, 1, 5, N - 1, N):
        pa, pc = lg.ring.Poly.from_numpy(w), ctx.NewPoly()
        ctx.Shift(pa, n, pc)
        assert np.array_equal(pc.numpy(), orc.ring_shift(o, w, n)), n
        ctx.Shift(pa, n, pa)  # in place (append builds a new slice)
        assert np.array_equal(pa.numpy(), orc.ring_shift(o, w, n)), n
    with pytest.raises(lg.LattigpuError):
        ctx.Shift(lg.ring.Poly.from_numpy(a), N + 1, ctx.NewPoly())  # Go: slice bounds out of range
    for n in (0, 1, 3, 2 * N + 1):
        for src in (a, w):
            pa, pb = lg.ring.Poly.from_numpy(src), lg.ring.Poly.from_numpy(w)
            ctx.Rotate(pa, n, pb)
            assert np.array_equal(pa.numpy(), orc.ring_rotate(o, src, n)), n
            assert np.array_equal(pb.numpy(), w)  # never written (ring.go:791)
    for e in (0, 1, 5):
        pa, pb = lg.ring.Poly.from_numpy(a), ctx.NewPoly()
        ctx.Exp(pa, e, pb)
        w1, w2 = orc.ring_exp(o, a, e)
        assert np.array_equal(pa.numpy(), w1) and np.array_equal(pb.numpy(), w2), e
    # Equal reduces both operands in place and compares
    qcol = np.array(moduli, dtype=np.uint64)[:, None]
    shifted = a + qcol  # same residues, other words
    pa, pb = lg.ring.Poly.from_numpy(a), lg.ring.Poly.from_numpy(shifted)
    eq, ra, rb = orc.ring_equal(o, a, shifted)
    assert eq and ctx.Equal(pa, pb)
    assert np.array_equal(pb.numpy(), rb)
    other = a.copy()
    other[2, N - 1] ^= np.uint64(1)
    assert not ctx.Equal(lg.ring.Poly.from_numpy(a), lg.ring.Poly.from_numpy(other))
    assert ctx.EqualLvl(1, lg.ring.Poly.from_numpy(a), lg.ring.Poly.from_numpy(other))  # the last limb is not looked at


@pytest.mark.parametrize("N", [1 << 13, 1 << 16])
def test_galois_shift_property(lg, N):
    """ring/ring_test.go:422-450: BitReverse, InvNTT, Rotate(1), NTT, BitReverse, Reduce == Shift(1)"""
    moduli = Q45 if N == 1 << 16 else QI60
    rng = np.random.default_rng(3)
    ctx = lg.ring.NewContextWithParams(N, moduli)
    a = uniform(rng, moduli, N)
    want, test, tmp = lg.ring.Poly.from_numpy(a), lg.ring.Poly.from_numpy(a), ctx.NewPoly()
    ctx.BitReverse(test, tmp)  # not in place here (lattigpu.h: p1 != p2)
    ctx.InvNTT(tmp, tmp)
    ctx.Rotate(tmp, 1, tmp)
    ctx.NTT(tmp, tmp)
    ctx.BitReverse(tmp, test)
    ctx.Reduce(test, test)
    ctx.Shift(want, 1, want)
    assert np.array_equal(test.numpy(), want.numpy())
