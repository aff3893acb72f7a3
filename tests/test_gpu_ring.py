"""GPU parity tests for the ring layer: every C-ABI op against the CPU oracle on
the same seeded inputs, bit-exact.  Structured after the reference's
ring/ring_test.go and ring/ntt_test.go (known-answer NTT vectors first, then
op-by-op checks), with in-contract inputs (uniform in [0,q)) and the
out-of-contract 64-bit words that ring.NewPolyUniform produces
(ring/ring_object.go:26-46) as a formula-exactness probe.
"""
import os

import numpy as np
import pytest

from oracle import ring_oracle as orc

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(__file__), "golden", "ring_test_data")
QI60 = [1152921504606584833, 1152921504598720513, 1152921504592429057, 1152921504581419009, 1152921504580894721,
        1152921504578273281, 1152921504577748993, 1152921504577486849, 1152921504066306049, 1152921504057917441,
        1152921504053723137, 1152921504050839553]  # ring/params.go:50-69 (head and tail)
PI60 = [576460752308273153, 576460752315482113, 576460752319021057, 576460752319414273, 576460752321642497,
        576460752325705729, 576460752328327169, 576460752329113601, 576460752568975361, 576460752573431809,
        576460752580902913, 576460752585490433]  # ring/params.go:28-47 (head and tail)


@pytest.fixture(scope="module")
def lg():
    import lattigpu
    from lattigpu import ring

    ring.set_device(0)
    return lattigpu


def uniform(rng, moduli, N, batch=None):
    shape = (N,) if batch is None else (batch, N)
    a = np.stack([rng.integers(0, q, size=shape, dtype=np.uint64) for q in moduli], axis=-2)
    return np.ascontiguousarray(a)


def words(rng, nl, N, batch=None):
    shape = (nl, N) if batch is None else (batch, nl, N)
    return rng.integers(0, 1 << 64, size=shape, dtype=np.uint64)


def load_golden(name):
    with open(os.path.join(GOLD, name)) as f:
        lines = [l for l in f.read().split("\n") if l.strip()]
    N = int(lines[0])
    moduli = [int(x) for x in lines[1].split()]
    return N, moduli, np.array([[int(x) for x in lines[2 + i].split()] for i in range(len(moduli))], dtype=np.uint64)


# ---------------------------------------------------------------------------
# known-answer vectors (ring/ntt_test.go:101-142), all N coefficients
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("n", [8, 16, 32, 64, 128, 256, 512])
def test_ntt_golden_vectors(lg, n):
    w = str(n).rjust(4, "_")
    N, moduli, x = load_golden("test_pol_60_%s_2" % w)
    _, _, want = load_golden("test_pol_NTT_60_%s_2" % w)
    ctx = lg.ring.NewContextWithParams(N, moduli)
    p = lg.ring.Poly.from_numpy(x)
    out = ctx.NewPoly()
    ctx.NTT(p, out)
    assert np.array_equal(out.numpy(), want)
    ctx.InvNTT(out, out)
    assert np.array_equal(out.numpy(), x)


# ---------------------------------------------------------------------------
# context tables: native GenNTTParams == oracle (ring_context.go:129-209)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("N,moduli", [
    (512, [576460752303439873, 576460752303702017]),
    (1 << 13, QI60[-4:]),
    (1 << 14, [0x200000008001, 0x400018001, 0x400060001, 0x80000050001, 0x800000B8001]),
    (1 << 16, [0x80000000080001, 0x2000000A0001, 0x80000000440001]),
])
def test_context_tables(lg, N, moduli):
    o = orc.Context(N, moduli)
    ctx = lg.ring.NewContextWithParams(N, moduli)
    t = ctx.tables()
    psi, psi_inv = o.all_tables()
    assert np.array_equal(t["psi"], psi) and np.array_equal(t["psi_inv"], psi_inv)
    assert np.array_equal(t["bred"], o.bred) and np.array_equal(t["mred"], o.mred) and np.array_equal(t["ninv"], o.ninv)
    flat = [v for row in o.rescale_params() for v in row]
    assert [int(v) for v in t["rescale"][: len(flat)]] == flat


def test_context_rejects_bad_moduli(lg):
    # GenNTTParams returns an error for non-NTT-friendly moduli (ring_context.go:142-145)
    with pytest.raises(lg.LattigpuError, match="does not allow NTT"):
        lg.ring.NewContextWithParams(1 << 13, [QI60[-1], 1152921504606846975])
    with pytest.raises(lg.LattigpuError, match="power of 2"):  # :72
        lg.ring.NewContextWithParams(1000, [QI60[-1]])


# ---------------------------------------------------------------------------
# NTT / InvNTT over every kernel schedule
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("logN", [1, 2, 3, 5, 8, 10, 11, 12, 13, 14, 15, 16])
@pytest.mark.parametrize("kind", ["reduced", "words"])
def test_ntt_parity(lg, logN, kind):
    N = 1 << logN
    moduli = QI60[-3:] + PI60[-2:]
    rng = np.random.default_rng(100 + logN)
    o = orc.Context(N, moduli)
    ctx = lg.ring.NewContextWithParams(N, moduli)
    batch = 3 if logN <= 14 else 2
    x = uniform(rng, moduli, N, batch) if kind == "reduced" else words(rng, len(moduli), N, batch)
    p = lg.ring.Poly.from_numpy(x)
    out = ctx.NewPoly(batch)
    ctx.NTT(p, out)
    want = np.stack([o.ntt(x[b]) for b in range(batch)])
    assert np.array_equal(out.numpy(), want)
    ctx.InvNTT(p, out)
    want = np.stack([o.invntt(x[b]) for b in range(batch)])
    assert np.array_equal(out.numpy(), want)
    # in place, partial level (NTTLvl / InvNTTLvl, ntt.go:10-15, :24-29)
    q = lg.ring.Poly.from_numpy(x)
    ctx.NTTLvl(2, q, q)
    got = q.numpy()
    assert np.array_equal(got[:, :3], np.stack([o.ntt(x[b], nl=3) for b in range(batch)]))
    assert np.array_equal(got[:, 3:], x[:, 3:])
    ctx.InvNTTLvl(2, q, q)
    if kind == "reduced":
        assert np.array_equal(q.numpy(), x)


@pytest.mark.parametrize("logN", [12, 16])
def test_invntt_mixed_batch(lg, logN):
    """A batch whose entries differ in kind: in-range words in some entries, arbitrary 64-bit words in others, for moduli
    of every butterfly class.  The inverse transform's two phases exchange raw doubles (FP64-only butterflies) or integers
    (literal butterflies, taken when a word exceeds 2q) through HBM, so they must take the same decision per limb whatever
    their CTAs' grouping of the batch entries."""
    N = 1 << logN
    moduli = orc.generate_ntt_primes(45, logN, 2) + orc.generate_ntt_primes(55, logN, 1) + orc.generate_ntt_primes(60, logN, 1)
    rng = np.random.default_rng(41 + logN)
    o = orc.Context(N, moduli)
    ctx = lg.ring.NewContextWithParams(N, moduli)
    batch = 5
    a = uniform(rng, moduli, N, batch)
    a[1] = words(rng, len(moduli), N)          # every limb of entry 1 out of range
    a[3, 0] = words(rng, 1, N)[0]              # one limb of entry 3
    a[4, 2, 7] = np.uint64((1 << 64) - 1)      # a single word
    pa, pb = lg.ring.Poly.from_numpy(a), lg.ring.Poly(N, len(moduli), batch)
    ctx.InvNTT(pa, pb)
    got = pb.numpy(squeeze=False)
    for b in range(batch):
        assert np.array_equal(got[b], o.invntt(np.ascontiguousarray(a[b]))), b
    ctx.NTT(pa, pb)
    got = pb.numpy(squeeze=False)
    for b in range(batch):
        assert np.array_equal(got[b], o.ntt(np.ascontiguousarray(a[b]))), b


def test_ntt_from_go_tables_and_single_limb(lg):
    """tables supplied by the host language (lg_ring_create_from_tables) and the free
    functions ring.NTT / ring.InvNTT on one limb (ntt.go:53, :89)"""
    N, moduli = 1 << 12, QI60[-2:] + PI60[-1:]
    rng = np.random.default_rng(7)
    o = orc.Context(N, moduli)
    psi, psi_inv = o.all_tables()
    flat = [v for row in o.rescale_params() for v in row]
    ctx = lg.ring.Context.from_tables(N, moduli, o.bred, o.mred, psi, psi_inv, o.ninv, flat)
    x = uniform(rng, moduli, N)
    p = lg.ring.Poly.from_numpy(x)
    out = ctx.NewPoly()
    ctx.NTT(p, out)
    assert np.array_equal(out.numpy(), o.ntt(x))
    # limb 0 of p transformed with the tables of prime 2 into limb 1 of out
    lg.ring.NTT(ctx, 2, p, 0, out, 1)
    want = np.zeros(N, np.uint64)
    orc.lib().orc_ntt_one(o.h, 2, orc.ptr(x[0]), orc.ptr(want))
    assert np.array_equal(out.numpy()[1], want)
    lg.ring.InvNTT(ctx, 2, p, 0, out, 1)
    orc.lib().orc_invntt_one(o.h, 2, orc.ptr(x[0]), orc.ptr(want))
    assert np.array_equal(out.numpy()[1], want)


# ---------------------------------------------------------------------------
# coefficient-wise ops (ring/ring.go)
# ---------------------------------------------------------------------------
OPS3 = [
    ("Add", "add"), ("AddNoMod", "add_nomod"), ("Sub", "sub"), ("SubNoMod", "sub_nomod"),
    ("MulCoeffs", "mulcoeffs"), ("MulCoeffsAndAdd", "mulcoeffs_and_add"),
    ("MulCoeffsAndAddNoMod", "mulcoeffs_and_add_nomod"), ("MulCoeffsConstant", "mulcoeffs_constant"),
    ("MulCoeffsMontgomery", "mulcoeffs_montgomery"), ("MulCoeffsMontgomeryAndAdd", "mulcoeffs_montgomery_and_add"),
    ("MulCoeffsMontgomeryAndAddNoMod", "mulcoeffs_montgomery_and_add_nomod"),
    ("MulCoeffsMontgomeryAndSub", "mulcoeffs_montgomery_and_sub"),
    ("MulCoeffsMontgomeryAndSubNoMod", "mulcoeffs_montgomery_and_sub_nomod"),
    ("MulCoeffsMontgomeryConstant", "mulcoeffs_montgomery_constant"),
]
OPS2 = [("Neg", "neg"), ("Reduce", "reduce"), ("MForm", "mform_poly"), ("InvMForm", "invmform_poly"),
        ("BitReverse", "bitreverse_poly")]


@pytest.mark.parametrize("kind", ["reduced", "words"])
@pytest.mark.parametrize("N", [8, 1 << 13])
def test_coefficientwise_ops(lg, N, kind):
    moduli = QI60[-4:]  # the reference's N=2^13 test shape (ring/params.go:12)
    rng = np.random.default_rng(N + len(kind))
    o = orc.Context(N, moduli)
    ctx = lg.ring.NewContextWithParams(N, moduli)
    gen = (lambda: uniform(rng, moduli, N)) if kind == "reduced" else (lambda: words(rng, 4, N))
    a, b, c = gen(), gen(), gen()
    pa, pb = lg.ring.Poly.from_numpy(a), lg.ring.Poly.from_numpy(b)
    for go, oname in OPS3:
        pc = lg.ring.Poly.from_numpy(c)
        getattr(ctx, go)(pa, pb, pc)
        want = o.op3(oname, a, b, c.copy())
        assert np.array_equal(pc.numpy(), want), go
    for go, oname in OPS2:
        pc = lg.ring.Poly.from_numpy(c)
        getattr(ctx, go)(pa, pc)
        assert np.array_equal(pc.numpy(), o.op2(oname, a)), go
    # Lvl variants touch only level+1 limbs
    pc = lg.ring.Poly.from_numpy(c)
    ctx.MulCoeffsMontgomeryAndAddLvl(1, pa, pb, pc)
    want = c.copy()
    want[:2] = o.op3("mulcoeffs_montgomery_and_add", a, b, c.copy(), nl=2)[:2]
    assert np.array_equal(pc.numpy(), want)
    pc = lg.ring.Poly.from_numpy(c)
    ctx.MulCoeffsMontgomeryConstantAndAddNoModLvl(2, pa, pb, pc)
    want = c.copy()
    want[:3] = o.op3("mulcoeffs_montgomery_constant_and_add_nomod", a, b, c.copy(), nl=3)[:3]
    assert np.array_equal(pc.numpy(), want)
    # aliasing: out == in (ring.go ops allow it)
    pc = lg.ring.Poly.from_numpy(a)
    ctx.Add(pc, pc, pc)
    assert np.array_equal(pc.numpy(), o.op3("add", a, a))


def test_scalar_and_word_ops(lg):
    N, moduli = 1 << 12, QI60[-3:]
    rng = np.random.default_rng(11)
    o = orc.Context(N, moduli)
    L = orc.lib()
    ctx = lg.ring.NewContextWithParams(N, moduli)
    a = uniform(rng, moduli, N)
    big = (1 << 200) + 12345678901234567890123
    for scalar, bigint in [(3, False), ((1 << 64) - 5, False), (big, True)]:
        sc = orc.arr([scalar % q for q in moduli] if bigint else [scalar] * 3)
        pa, out = lg.ring.Poly.from_numpy(a), ctx.NewPoly()
        (ctx.MulScalarBigint if bigint else ctx.MulScalar)(pa, scalar, out)
        assert np.array_equal(out.numpy(), o.mul_scalar(a, sc))
        want = a.copy()
        L.orc_add_scalar(o.h, 3, orc.ptr(want), orc.ptr(sc))
        (ctx.AddScalarBigint if bigint else ctx.AddScalar)(pa, scalar, pa)
        assert np.array_equal(pa.numpy(), want)
        L.orc_sub_scalar(o.h, 3, orc.ptr(want), orc.ptr(sc))
        (ctx.SubScalarBigint if bigint else ctx.SubScalar)(pa, scalar, pa)
        assert np.array_equal(pa.numpy(), want)
    pa, out = lg.ring.Poly.from_numpy(a), ctx.NewPoly()
    for pow2 in (0, 1, 17, 63):
        want = np.zeros_like(a)
        L.orc_mul_by_pow2(o.h, 3, orc.ptr(a), pow2, orc.ptr(want))
        ctx.MulByPow2(pa, pow2, out)
        assert np.array_equal(out.numpy(), want), pow2
    inpl = lg.ring.Poly.from_numpy(a)
    want = a.copy()
    L.orc_mul_by_pow2(o.h, 3, orc.ptr(want), 9, orc.ptr(want))
    ctx.MulByPow2(inpl, 9, inpl)
    assert np.array_equal(inpl.numpy(), want)
    for deg in (0, 1, N - 1, N, N + 5, 2 * N - 1, 5 * N + 3):
        want = np.zeros_like(a)
        L.orc_mult_by_monomial(o.h, 3, orc.ptr(a), deg, orc.ptr(want))
        ctx.MultByMonomial(pa, deg, out)
        assert np.array_equal(out.numpy(), want), deg
    m = 0xFFFF0000FFFF
    ctx.AND(pa, m, out)
    assert np.array_equal(out.numpy(), a & np.uint64(m))
    ctx.OR(pa, m, out)
    assert np.array_equal(out.numpy(), a | np.uint64(m))
    ctx.XOR(pa, m, out)
    assert np.array_equal(out.numpy(), a ^ np.uint64(m))
    ctx.Mod(pa, 65537, out)
    assert np.array_equal(out.numpy(), a % np.uint64(65537))
    vec = rng.integers(0, moduli[0], size=(1, N), dtype=np.uint64)
    pv = lg.ring.Poly.from_numpy(vec)
    want = np.zeros_like(a)
    L.orc_mul_by_vector_montgomery(o.h, 3, orc.ptr(a), orc.ptr(vec[0]), orc.ptr(want))
    ctx.MulByVectorMontgomery(pa, pv, out)
    assert np.array_equal(out.numpy(), want)
    L.orc_mul_by_vector_montgomery_and_add_nomod(o.h, 3, orc.ptr(a), orc.ptr(vec[0]), orc.ptr(want))
    ctx.MulByVectorMontgomeryAndAddNoMod(pa, pv, out)
    assert np.array_equal(out.numpy(), want)


# ---------------------------------------------------------------------------
# Galois automorphisms (ring/ring_galois.go), cf. testGaloisShift ring_test.go:422
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("N", [16, 1 << 12, 1 << 15])
def test_galois(lg, N):
    moduli = QI60[-2:]
    rng = np.random.default_rng(N)
    o = orc.Context(N, moduli)
    ctx = lg.ring.NewContextWithParams(N, moduli)
    a = uniform(rng, moduli, N)
    a[0, 0] = 0  # Permute maps 0 to q on sign flips (ring_galois.go:124)
    pa, out = lg.ring.Poly.from_numpy(a), ctx.NewPoly()
    for gen, power in [(5, 1), (5, 7), (5, N // 2 - 1), (2 * N - 1, 1)]:
        idx = lg.ring.PermuteNTTIndex(gen, power, N)
        want_idx = orc.permute_ntt_index(gen, power, N)
        assert np.array_equal(idx.numpy(), want_idx)
        lg.ring.PermuteNTTWithIndex(pa, idx, out)
        assert np.array_equal(out.numpy(), orc.permute_ntt_with_index(a, want_idx))
        g = pow(gen, power, 2 * N)
        lg.ring.PermuteNTT(pa, g, out)
        assert np.array_equal(out.numpy(), orc.permute_ntt_with_index(a, want_idx))
        ctx.Permute(pa, g, out)
        assert np.array_equal(out.numpy(), o.permute(a, g))
    # index computed by the host language
    idx2 = lg.ring.GaloisIndex(index=orc.permute_ntt_index(5, 3, N))
    lg.ring.PermuteNTTWithIndex(pa, idx2, out)
    assert np.array_equal(out.numpy(), orc.permute_ntt_with_index(a, orc.permute_ntt_index(5, 3, N)))
    with pytest.raises(lg.LattigpuError, match="not in place"):
        ctx.Permute(pa, 5, pa)


# ---------------------------------------------------------------------------
# RNS rescaling (ring/ring_scaling.go), cf. ring_test.go:134-220
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("N", [64, 1 << 13])
@pytest.mark.parametrize("kind", ["reduced", "words"])
def test_div_by_last_modulus(lg, N, kind):
    moduli = QI60[-4:]
    rng = np.random.default_rng(N + 3)
    o = orc.Context(N, moduli)
    ctx = lg.ring.NewContextWithParams(N, moduli)
    a = uniform(rng, moduli, N, 2) if kind == "reduced" else words(rng, 4, N, 2)
    L = orc.lib()
    for name in ["div_floor_by_last_modulus_ntt", "div_floor_by_last_modulus", "div_round_by_last_modulus_ntt",
                 "div_round_by_last_modulus"]:
        go = "".join(w.upper() if w == "ntt" else w.capitalize() for w in name.split("_"))
        for nl in (4, 3, 2):
            p = lg.ring.Poly.from_numpy(a)
            getattr(ctx, go)(p, nl=nl)
            got = p.numpy()
            for b in range(2):
                want = a[b].copy()
                getattr(L, "orc_" + name)(o.h, nl, orc.ptr(want))
                assert np.array_equal(got[b, : nl - 1], want[: nl - 1]), (name, nl)
                assert np.array_equal(got[b, nl:], a[b, nl:])
    for name in ["div_floor_by_last_modulus_many", "div_floor_by_last_modulus_many_ntt",
                 "div_round_by_last_modulus_many", "div_round_by_last_modulus_many_ntt"]:
        go = "".join(w.upper() if w == "ntt" else w.capitalize() for w in name.split("_"))
        for nb in (1, 2, 3):
            p = lg.ring.Poly.from_numpy(a)
            getattr(ctx, go)(p, nb)
            got = p.numpy()
            for b in range(2):
                want = a[b].copy()
                getattr(L, "orc_" + name)(o.h, 4, orc.ptr(want), nb)
                assert np.array_equal(got[b, : 4 - nb], want[: 4 - nb]), (name, nb)


# ---------------------------------------------------------------------------
# FastBasisExtender (ring/ring_basis_extension.go:9-393), cf. ring_test.go:550
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("N,nQ,nP", [(32, 4, 4), (1 << 12, 2, 2), (1 << 13, 8, 3), (1 << 12, 12, 12), (1 << 12, 16, 4)])
@pytest.mark.parametrize("kind", ["reduced", "words"])
def test_basis_extender(lg, N, nQ, nP, kind):
    # 5..16 source limbs take the wide column-form kernel (8 / 12 / 16 instantiations), 1..4 the fast ones
    Qm, Pm = (QI60 + PI60[: nQ - 12] if nQ > 12 else QI60[-nQ:]), PI60[-nP:]
    rng = np.random.default_rng(N + nQ)
    oQ, oP = orc.Context(N, Qm), orc.Context(N, Pm)
    oe = orc.Extender(oQ, oP)
    cQ, cP = lg.ring.NewContextWithParams(N, Qm), lg.ring.NewContextWithParams(N, Pm)
    be = lg.ring.NewFastBasisExtender(cQ, cP)
    gen = (lambda m: uniform(rng, m, N)) if kind == "reduced" else (lambda m: words(rng, len(m), N))
    aQ, aP = gen(Qm), gen(Pm)
    aQP = np.concatenate([aQ, aP])
    pQ, pP = lg.ring.Poly.from_numpy(aQ), lg.ring.Poly.from_numpy(aP)
    for level in sorted({nQ - 1, nQ // 2, 0}):
        out = cP.NewPoly()
        be.ModUpSplitQP(level, pQ, out)
        assert np.array_equal(out.numpy(), oe.modup_split_qp(level, aQ)), ("ModUpSplitQP", level)
        out = cQ.NewPoly()
        pqp = lg.ring.Poly.from_numpy(aQP)
        be.ModDownNTTPQ(level, pqp, out)
        assert np.array_equal(out.numpy()[: level + 1], oe.moddown_ntt_pq(level, aQP)), ("ModDownNTTPQ", level)
        out = cQ.NewPoly()
        pPc = lg.ring.Poly.from_numpy(aP)
        be.ModDownSplitedNTTPQ(level, pQ, pPc, out)
        assert np.array_equal(out.numpy()[: level + 1], oe.moddown_splited_ntt_pq(level, aQ, aP))
        # in place on the Q part, as switchKeysInPlace does (ckks/evaluator.go:1556)
        pQc, pPc = lg.ring.Poly.from_numpy(aQ), lg.ring.Poly.from_numpy(aP)
        be.ModDownSplitedNTTPQ(level, pQc, pPc, pQc)
        assert np.array_equal(pQc.numpy()[: level + 1], oe.moddown_splited_ntt_pq(level, aQ, aP))
        packed = np.concatenate([aQ[: level + 1], aP])
        out = cQ.NewPoly()
        be.ModDownPQ(level, lg.ring.Poly.from_numpy(packed), out)
        assert np.array_equal(out.numpy()[: level + 1], oe.moddown_pq(level, packed)), ("ModDownPQ", level)
        out = cQ.NewPoly()
        be.ModDownSplitedPQ(level, pQ, pP, out)
        assert np.array_equal(out.numpy()[: level + 1], oe.moddown_splited_pq(level, aQ, aP))
    for levelP in sorted({nP - 1, 0}):
        out = cQ.NewPoly()
        be.ModUpSplitPQ(levelP, pP, out)
        assert np.array_equal(out.numpy(), oe.modup_split_pq(levelP, aP)), ("ModUpSplitPQ", levelP)
        out = cP.NewPoly()
        be.ModDownSplitedQP(nQ - 1, levelP, pQ, pP, out)
        assert np.array_equal(out.numpy()[: levelP + 1], oe.moddown_splited_qp(nQ - 1, levelP, aQ, aP))


# ---------------------------------------------------------------------------
# Decomposer (ring/ring_basis_extension.go:398-713)
# ---------------------------------------------------------------------------
@pytest.mark.parametrize("N,nQ,nP", [(64, 5, 2), (1 << 12, 10, 2), (1 << 12, 7, 3), (1 << 12, 9, 4), (1 << 12, 4, 1)])
@pytest.mark.parametrize("kind", ["reduced", "words"])
def test_decomposer(lg, N, nQ, nP, kind):
    Qm, Pm = QI60[-nQ:], PI60[-nP:]
    rng = np.random.default_rng(N + nQ + nP)
    od = orc.Decomposer(Qm, Pm, N)
    d = lg.ring.NewDecomposer(N, Qm, Pm)
    assert d.beta == od.beta
    assert d.Xalpha() == [orc.lib().orc_decomposer_xalpha(od.h, i) for i in range(od.beta)]
    a = uniform(rng, Qm, N, 2) if kind == "reduced" else words(rng, nQ, N, 2)
    p0 = lg.ring.Poly.from_numpy(a)
    for level in range(nQ - 1, -1, -1):
        beta = -(-(level + 1) // nP)
        for crt in range(beta):
            p1Q, p1P = lg.ring.Poly(N, level + 1, 2), lg.ring.Poly(N, nP, 2)
            d.DecomposeAndSplit(level, crt, p0, p1Q, p1P)
            p1 = lg.ring.Poly(N, level + 1 + nP, 2)
            d.Decompose(level, crt, p0, p1)
            gq, gp, g1 = p1Q.numpy(), p1P.numpy(), p1.numpy()
            for b in range(2):
                wq, wp = od.decompose_and_split(level, crt, a[b])
                assert np.array_equal(gq[b], wq) and np.array_equal(gp[b], wp), (level, crt)
                assert np.array_equal(g1[b], od.decompose(level, crt, a[b])), (level, crt)


def test_poly_wire_format(lg):
    """MarshalBinary / WriteCoeffs / UnmarshalBinary (ring/ring_object.go:146-289) against the oracle's bytes:
    device-side byte swap, any batch entry, partial limb count, error on a truncated encoding."""
    N, nl, batch = 4096, 3, 2
    rng = np.random.default_rng(71)
    a = rng.integers(0, 1 << 64, size=(batch, nl, N), dtype=np.uint64)
    p = lg.ring.Poly.from_numpy(a)
    for b in range(batch):
        assert p.MarshalBinary(batch_index=b) == orc.poly_marshal(a[b])
        assert p.MarshalBinary(batch_index=b, nl=2) == orc.poly_marshal(a[b, :2])
        assert p.MarshalBinary(batch_index=b, WithMetadata=False) == orc.poly_marshal(a[b], with_metadata=False)
    assert p.GetDataLen() == 2 + nl * N * 8 and p.GetDataLen(False) == nl * N * 8
    q = lg.ring.Poly(N, nl, batch)
    q.Zero()
    q.UnmarshalBinary(orc.poly_marshal(a[1]), batch_index=0)
    q.UnmarshalBinary(orc.poly_marshal(a[0]), batch_index=1)
    got = q.numpy(squeeze=False)
    assert np.array_equal(got[0], a[1]) and np.array_equal(got[1], a[0])
    assert np.array_equal(orc.poly_unmarshal(p.MarshalBinary(1)), a[1])
    with pytest.raises(lg.LattigpuError, match="invalid polynomial encoding"):
        q.UnmarshalBinary(orc.poly_marshal(a[0])[:-8])
    with pytest.raises(lg.LattigpuError, match="moduli encoded"):
        lg.ring.Poly(N, 2, 1).UnmarshalBinary(orc.poly_marshal(a[0]))


@pytest.mark.parametrize("logN", [12, 14, 16])
def test_ntt_zero_and_sparse_inputs(lg, logN):
    """Zero and sparse polynomials through every butterfly flavour (45-bit: FP64 quotient, where y = 0 makes the
    quotient estimate -1; 55-bit: Shoup; 60-bit: [0,8q) / [0,4q)), forward and inverse, against the oracle."""
    N = 1 << logN
    moduli = orc.generate_ntt_primes(45, logN, 2) + orc.generate_ntt_primes(55, logN, 1) + orc.generate_ntt_primes(60, logN, 1)
    octx = orc.Context(N, moduli)
    ctx = lg.ring.NewContextWithParams(N, moduli)
    rng = np.random.default_rng(83 + logN)
    a = np.zeros((3, len(moduli), N), dtype=np.uint64)
    pos = rng.integers(0, N, size=37)
    for i, q in enumerate(moduli):
        a[1, i, pos] = rng.integers(0, q, size=37, dtype=np.uint64)  # sparse
        a[2, i] = rng.integers(0, q, size=N, dtype=np.uint64)
        a[2, i, ::2] = 0  # every other coefficient zero
        a[2, i, 1] = q - 1
    p = lg.ring.Poly.from_numpy(a)
    out = lg.ring.Poly(N, len(moduli), 3)
    ctx.NTT(p, out)
    got = out.numpy(squeeze=False)
    for b in range(3):
        assert np.array_equal(got[b], octx.ntt(np.ascontiguousarray(a[b]))), ("fwd", b)
    ctx.InvNTT(p, out)
    got = out.numpy(squeeze=False)
    for b in range(3):
        assert np.array_equal(got[b], octx.invntt(np.ascontiguousarray(a[b]))), ("inv", b)


def _prime_below(bound, two_n):
    """largest NTT-friendly prime below `bound` (p = 1 mod 2N)"""
    def is_prime(n):
        if n < 2:
            return False
        for sp in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
            if n % sp == 0:
                return n == sp
        d, r = n - 1, 0
        while d % 2 == 0:
            d //= 2
            r += 1
        for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
            x = pow(a, d, n)
            if x in (1, n - 1):
                continue
            for _ in range(r - 1):
                x = x * x % n
                if x == n - 1:
                    break
            else:
                return False
        return True

    p = (bound - 1) // two_n * two_n + 1
    while not is_prime(p):
        p -= two_n
    return p


@pytest.mark.parametrize("logN", [12, 13, 14, 15, 16])
@pytest.mark.parametrize("no_d64", [0, 1], ids=["d64", "int"])
def test_ntt_parity_fp64_butterfly_moduli(lg, logN, no_d64):
    """Moduli below 3*2^44 take the FP64-only butterflies (values kept as doubles, modarith.cuh) in both directions, or the
    integer ones behind the "no_d64_ntt" switch.  Both must match the oracle on: uniform residues, arbitrary 64-bit words,
    and the inputs that drive the lazy values to their extremes (all q-1; alternating 0 / q-1; all 2q for the inverse).
    The largest admissible prime (just below 3*2^44) and its neighbour above (integer path) are included."""
    N = 1 << logN
    edge = _prime_below(3 << 44, 2 * N)
    above = _prime_below((3 << 44) + (1 << 40), 2 * N)
    assert edge < (3 << 44) <= above
    moduli = orc.generate_ntt_primes(30, logN, 1) + orc.generate_ntt_primes(45, logN, 2) + [edge, above]
    octx = orc.Context(N, moduli)
    ctx = lg.ring.NewContextWithParams(N, moduli)
    rng = np.random.default_rng(300 + logN)
    nl = len(moduli)
    a = np.zeros((5, nl, N), dtype=np.uint64)
    for i, q in enumerate(moduli):
        a[0, i] = rng.integers(0, q, size=N, dtype=np.uint64)
        a[2, i] = q - 1
        a[3, i, ::2] = q - 1
        a[4, i] = 2 * q  # the inverse's in-range limit (values <= 2q)
    a[1] = rng.integers(0, 1 << 64, size=(nl, N), dtype=np.uint64)
    try:
        lg.ring.debug_set_switch("no_d64_ntt", no_d64)
        p = lg.ring.Poly.from_numpy(a)
        out = lg.ring.Poly(N, nl, 5)
        ctx.NTT(p, out)
        got = out.numpy(squeeze=False)
        for b in range(5):
            assert np.array_equal(got[b], octx.ntt(np.ascontiguousarray(a[b]))), ("fwd", b)
        ctx.InvNTT(p, out)
        got = out.numpy(squeeze=False)
        for b in range(5):
            want = octx.invntt(np.ascontiguousarray(a[b]))
            bad = [i for i in range(nl) if not np.array_equal(got[b, i], want[i])]
            assert not bad, ("inv", b, bad, [int(v) for v in got[b, bad[0]][:4]], [int(v) for v in want[bad[0]][:4]], moduli[bad[0]])
        ctx.InvNTT(out, out)  # in place on canonical data
        ctx.NTT(out, out)
        assert np.array_equal(out.numpy(squeeze=False), got)
    finally:
        lg.ring.debug_set_switch("no_d64_ntt", 0)


@pytest.mark.parametrize("src", [[45] * 4, [45] * 3, [34, 34], [45], [46] * 3, [55] * 4, [55, 45, 45, 45], [55, 55], [49] * 3, [56]],
                         ids=["4x45", "3x45", "2x34", "1x45", "3x46", "4x55", "55+3x45", "2x55", "3x49", "1x56"])
@pytest.mark.parametrize("dst_bits", [[60, 60, 55, 45, 34], [36, 59], [55, 45, 45, 34]], ids=["wide", "narrow", "ckks"])
def test_modup_fp64_quotient_path(lg, src, dst_bits):
    """modup_fp_kernel / modup_fp2_kernel (csrc/basisext.cu): sources summing below 2^48 take the FP64-quotient basis
    extension, wider ones its two-step variant when the first remainder fits 64 bits for every target (e.g. the 55-bit
    special primes of a ModDown onto 45/55-bit targets; the 60-bit targets of the "wide" set send them back to the integer
    kernel).  Besides random residues, the inputs are built so that every y_i = MRed(a_i, qibMont_i) sits at the ends of
    its range (0, 1, q_i - 1, q_i - 2, q_i / 2 and mixtures), where the quotient estimates and the correction index v are
    extreme; bit-exact against the oracle's modUpExact (ring_basis_extension.go:352-393)."""
    logN = 10
    N = 1 << logN
    nsrc = len(src)
    Qm = []
    for b in src:
        Qm.append([p for p in orc.generate_ntt_primes(b, logN, nsrc + 1) if p not in Qm][0])
    Pm = []
    for b in dst_bits:
        Pm.append([p for p in orc.generate_ntt_primes(b, logN, len(dst_bits) + nsrc + 1) if p not in Qm and p not in Pm][0])
    oe = orc.Extender(orc.Context(N, Qm), orc.Context(N, Pm))
    cQ, cP = lg.ring.NewContextWithParams(N, Qm), lg.ring.NewContextWithParams(N, Pm)
    be = lg.ring.NewFastBasisExtender(cQ, cP)
    rng = np.random.default_rng(sum(src) * 10 + nsrc)
    a = np.stack([rng.integers(0, q, size=N, dtype=np.uint64) for q in Qm])
    Qbig = 1
    for q in Qm:
        Qbig *= q
    ends = lambda q: [0, 1, q - 1, q - 2, q // 2]
    for col in range(min(N, 400)):
        for i, q in enumerate(Qm):
            star = (Qbig // q) % q  # y = a * (Q/q)^-1 mod q  =>  a = y * (Q/q) mod q
            y = ends(q)[(col // (5 ** i)) % 5] if col < 5 ** min(nsrc, 3) else ends(q)[rng.integers(0, 5)]
            a[i, col] = (y * star) % q
    pQ = lg.ring.Poly.from_numpy(a)
    for level in range(nsrc):
        out = cP.NewPoly()
        be.ModUpSplitQP(level, pQ, out)
        assert np.array_equal(out.numpy(), oe.modup_split_qp(level, a)), level
    w = rng.integers(0, 1 << 64, size=(nsrc, N), dtype=np.uint64)  # unreduced words
    out = cP.NewPoly()
    be.ModUpSplitQP(nsrc - 1, lg.ring.Poly.from_numpy(w), out)
    assert np.array_equal(out.numpy(), oe.modup_split_qp(nsrc - 1, w))
