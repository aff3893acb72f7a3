"""CPU tests of the oracle's SimpleScaler / Float128 restatement (ring/ring_scaling.go:166-300,
ring/float128.go) and of the BFV key generator / encryptor / decryptor / encoder sequences
(bfv/keygen.go, encryptor.go, decryptor.go, encoder.go).

Anchors: (1) the reference's own property, ring/ring_test.go:587-624 testSimpleScaling -- Scale(x) must equal
round(t*x/Q) mod t for random x below Q, with the reference's T = 0x3ee0001 and its 60-bit moduli;
(2) an independent pure-Python restatement of the double-double operations (CPython floats are IEEE binary64
with one rounding per operation), compared value for value with the C oracle; (3) scheme-level semantics:
encode -> encrypt -> Mul -> Relinearize -> RotateColumns -> decrypt -> decode gives the rotated slot-wise
product, as bfv/bfv_test.go checks."""
import math
import random

import numpy as np
import pytest

from oracle import ring_oracle as orc

from test_oracle_properties import QI60, PI60, crt_poly, div_round, prod


# ---- independent literal restatement of ring/float128.go in Python floats -------------------------------
def two_sum(a, b):
    s = a + b
    bb = s - a
    return s, (a - (s - bb)) + (b - bb)


def quick_two_sum(a, b):
    s = a + b
    return s, b - (s - a)


def two_diff(a, b):
    s = a - b
    bb = s - a
    return s, (a - (s - bb)) - (b + bb)


def split(a):
    temp = 134217729.0 * a
    hi = temp - (temp - a)
    return hi, a - hi


def two_prod(a, b):
    p = a * b
    ah, al = split(a)
    bh, bl = split(b)
    return p, ((ah * bh - p) + ah * bl + al * bh) + al * bl


def f_add(a, b):
    s1, s2 = two_sum(a[0], b[0])
    t1, t2 = two_sum(a[1], b[1])
    s2 += t1
    s1, s2 = quick_two_sum(s1, s2)
    s2 += t2
    return quick_two_sum(s1, s2)


def f_mul(a, b):
    p1, p2 = two_prod(a[0], b[0])
    p2 += a[0] * b[1] + a[1] * b[0]
    return quick_two_sum(p1, p2)


def f_div(a, b):
    q1 = a[0] / b[0]
    p1, p2 = two_prod(q1, b[0])
    p2 += q1 * b[1]
    t0 = p1 + p2
    t1 = p2 - (t0 - p1)
    p3, p4 = two_diff(a[0], t0)
    v1, v2 = two_diff(a[1], t1)
    p4 += v1
    p3, p4 = quick_two_sum(p3, p4)
    p4 += v2
    r = (p3 + p4) / b[0]
    f0 = q1 + r
    return f0, r - (f0 - q1)


def set_u64(i):
    return float(i >> 12), float(i & 0xFFF) / 4096.0


def go_round(x):  # math.Round: half away from zero
    return math.copysign(math.floor(abs(x) + 0.5), x) if abs(x) < 2.0**52 else x


def to_u64(f):
    a = f[0] * 4096.0
    ai = int(a)
    return (ai + int(go_round((a - float(ai)) + f[1] * 4096.0))) & ((1 << 64) - 1)


def test_float128_ops_match_python_restatement():
    rng = random.Random(7)
    for _ in range(3000):
        a = set_u64(rng.getrandbits(rng.randrange(13, 62)))
        b = set_u64(rng.getrandbits(rng.randrange(13, 62)) | (1 << 12))
        q = f_div(a, b)
        assert orc.f128_op(2, a, b) == q
        m = f_mul(q, b)
        assert orc.f128_op(1, q, b) == m
        s = f_add(m, a)
        assert orc.f128_op(0, m, a) == s
        assert orc.f128_to_u64(a) == to_u64(a)
        assert orc.f128_to_u64(s) == to_u64(s)


def py_scaler_params(t, Q):
    """NewSimpleScaler ring_scaling.go:188-262 in Python integers + the float restatement above"""
    Qp = prod(Q)
    wi, ti = [], []
    for qi in Q:
        barre = pow(Qp // qi, -1, qi)
        tmp = f_mul(f_div((float(t), 0.0), set_u64(qi)), set_u64(barre))
        w = int(tmp[0])
        if t & (t - 1):
            w = (w << 64) % t  # MForm
        wi.append(w)
        ti.append(f_div(set_u64(barre * t % qi), set_u64(qi)))
    return wi, ti


def py_scale(t, Q, wi, ti, column):
    a, b = 0, (0.0, 0.0)
    rinv = pow(1 << 64, -1, t) if t & (t - 1) else None
    for j, x in enumerate(column):
        a += (wi[j] * x) & (t - 1) if rinv is None else wi[j] * x * rinv % t
        b = f_add(b, f_mul(ti[j], set_u64(x)))
    a += to_u64(b)
    return a & (t - 1) if rinv is None else a % t


@pytest.mark.parametrize("t", [0x3EE0001, 65537, 1 << 16], ids=["T_ref", "65537", "pow2"])
@pytest.mark.parametrize("nq", [2, 4])
def test_simple_scaler_vs_bigint_and_python(t, nq):
    N = 256
    Q = QI60[:nq]
    ctx = orc.Context(N, Q)
    sc = orc.Scaler(t, ctx)
    wi, ti = sc.params()
    pwi, pti = py_scaler_params(t, Q)
    assert [int(w) for w in wi] == pwi
    assert [tuple(x) for x in ti.tolist()] == [tuple(x) for x in pti]
    rng = random.Random(nq * 1000 + t % 97)
    Qp = prod(Q)
    # random values as in the reference's test plus the ends of the range; Q//2 sits within t/(2Q) of a rounding
    # tie, beyond the 106-bit double-double: it is only compared with the literal restatement below
    vals = [rng.randrange(Qp) for _ in range(N - 4)] + [0, 1, Qp - 1, Qp // 2]
    p = crt_poly(vals, Q)
    out = sc.scale(p, nl_out=nq)
    want = [div_round(v * t, Qp) % t for v in vals]  # ring_test.go:602-608
    assert out[0].tolist()[:-1] == want[:-1]
    assert all(np.array_equal(out[0], out[j]) for j in range(nq))
    for j in list(range(0, N, 17)) + [N - 1]:  # literal Python restatement, a few columns
        assert py_scale(t, Q, pwi, pti, [int(x) for x in p[:, j]]) == int(out[0, j])


def test_simple_scaler_default_bfv_moduli():
    """the BFV default parameter shapes (bfv/params.go:47-88): 39-bit ... 59-bit moduli, t = 65537"""
    t, N = 65537, 64
    for logq in ([39, 39], [54, 54, 54], [56, 55, 55, 54, 54, 54], [59, 59, 59] + [58] * 9):
        Q, _, _ = orc.gen_moduli(6, logq, [])
        ctx = orc.Context(N, Q)
        sc = orc.Scaler(t, ctx)
        rng = random.Random(len(Q))
        Qp = prod(Q)
        vals = [rng.randrange(Qp) for _ in range(N)]
        out = sc.scale(crt_poly(vals, Q))
        assert out[0].tolist() == [div_round(v * t, Qp) % t for v in vals]


def _small_params(N=64):
    logn = N.bit_length() - 1
    return orc.gen_moduli(logn, [39, 39, 38], [40, 40], [60, 60, 60])


def test_bfv_index_matrix_is_a_permutation():
    for N in (8, 64, 4096):
        idx = orc.bfv_index_matrix(N)
        assert sorted(idx.tolist()) == list(range(N))


def test_bfv_encode_decode_roundtrip():
    N, t = 64, 65537
    Q, P, _ = _small_params(N)
    S = orc.BfvScheme(Q, P, N, t)
    rng = np.random.default_rng(5)
    m = rng.integers(0, t, size=N, dtype=np.uint64)
    pt = S.encode_uint(m)
    assert np.array_equal(S.decode_uint(pt), m)
    # short input: the remaining slots are zero (encoder.go:84-86)
    pt = S.encode_uint(m[:10])
    assert np.array_equal(S.decode_uint(pt), np.concatenate([m[:10], np.zeros(N - 10, np.uint64)]))
    mi = rng.integers(-(t // 2), t // 2 + 1, size=N)
    assert np.array_equal(S.decode_int(S.encode_int(mi)), mi)
    # plaintext = Delta * m in every limb
    Qp = prod(Q)
    coeffs = S.T.invntt(_slots(S, m))[0]
    for i, q in enumerate(Q):
        assert pt is not None and [int(x) for x in S.encode_uint(m)[i]] == [(Qp // t) * int(c) % q for c in coeffs]


def _slots(S, m):
    s = np.zeros((1, S.N), dtype=np.uint64)
    s[0, S.index_matrix.astype(np.int64)] = m
    return s


def test_bfv_scheme_pipeline_semantics():
    """keygen -> encode -> encrypt (pk, sk, sk fast) -> decrypt; Mul -> Relinearize -> RotateColumns -> decode"""
    N, t = 64, 65537
    Q, P, QMul = _small_params(N)
    S = orc.BfvScheme(Q, P, N, t)
    ev = orc.BfvEvaluator(S.Q, orc.Context(N, QMul), S.P, t)
    rng = np.random.default_rng(77)
    tern = lambda: rng.integers(-1, 2, size=N)
    gauss = lambda: np.rint(rng.normal(0, 3.2, size=N)).astype(np.int64)
    unif = lambda mods: np.ascontiguousarray(np.stack([rng.integers(0, q, size=N, dtype=np.uint64) for q in mods]))
    sk = S.gen_secret_key(tern())
    pk = S.gen_public_key(sk, gauss(), unif(Q + P))
    m0 = rng.integers(0, t, size=N, dtype=np.uint64)
    m1 = rng.integers(0, t, size=N, dtype=np.uint64)
    pt0, pt1 = S.encode_uint(m0), S.encode_uint(m1)
    ct_pk = S.encrypt_pk(pt0, pk, tern(), gauss(), gauss())
    ct_sk = S.encrypt_sk(pt1, sk, unif(Q + P), gauss())
    ct_skf = S.encrypt_sk(pt1, sk, unif(Q), gauss(), fast=True)
    assert np.array_equal(S.decode_uint(S.decrypt(ct_pk, sk)), m0)
    assert np.array_equal(S.decode_uint(S.decrypt(ct_sk, sk)), m1)
    assert np.array_equal(S.decode_uint(S.decrypt(ct_skf, sk)), m1)
    # the reference's pk fast path leaves (plaintext, 0) in a fresh ciphertext (encryptor.go:174-192, :221)
    ct_pkf = S.encrypt_pk(pt0, pk, tern(), gauss(), gauss(), fast=True)
    assert np.array_equal(ct_pkf[0], pt0) and not ct_pkf[1].any()

    rlk = S.gen_relin_key(sk, [gauss() for _ in range(S.beta)], [unif(Q + P) for _ in range(S.beta)])
    ct2 = ev.tensor_and_rescale(np.ascontiguousarray(ct_pk), np.ascontiguousarray(ct_sk))
    prod_slots = (m0.astype(object) * m1.astype(object)) % t
    assert [int(x) for x in S.decode_uint(S.decrypt(ct2, sk))] == list(prod_slots)  # degree 2 decrypts
    ct = ev.relinearize(np.ascontiguousarray(ct2), rlk)
    assert [int(x) for x in S.decode_uint(S.decrypt(ct, sk))] == list(prod_slots)
    # RotateColumns by k: Galois element 5^k (bfv/bfv.go:70), both rows rotate left by k
    k = 3
    gen = pow(5, k, 2 * N)
    rot = S.gen_rot_key(sk, gen, [gauss() for _ in range(S.beta)], [unif(Q + P) for _ in range(S.beta)])
    ctr = ev.permute(np.ascontiguousarray(ct), gen, rot)
    got = [int(x) for x in S.decode_uint(S.decrypt(ctr, sk))]
    row = N // 2
    want = [prod_slots[(i + k) % row] for i in range(row)] + [prod_slots[row + (i + k) % row] for i in range(row)]
    assert got == want
    # key switch to another secret
    sk2 = S.gen_secret_key(tern())
    swk = S.gen_switching_key(sk, sk2, [gauss() for _ in range(S.beta)], [unif(Q + P) for _ in range(S.beta)])
    assert [int(x) for x in S.decode_uint(S.decrypt(ev.switch_keys(np.ascontiguousarray(ct), swk), sk2))] == list(prod_slots)


def test_host_parameter_generators_match_oracle():
    """the product's host-side NewSimpleScaler / GenLiftParams (csrc/scaler.cu, no device needed) against the oracle"""
    import ctypes as C

    import lattigpu
    from lattigpu import ring

    for t in (65537, 0x3EE0001, 1 << 16):
        for logq in ([39, 39], [59, 59, 59] + [58] * 9, [60, 60, 60, 60]):
            Q, _, _ = orc.gen_moduli(10, logq, [])
            wi, ti = orc.Scaler(t, orc.Context(1024, Q)).params()
            gwi, gti, add_param, mul_param = ring.SimpleScalerParams(t, Q)
            assert np.array_equal(gwi, wi) and np.array_equal(gti, ti)
            if t & (t - 1):
                assert add_param == (1 << 128) // t >> 64 and (mul_param * t) & ((1 << 64) - 1) == 1
            else:
                assert add_param == mul_param == t - 1
            q = np.array(Q, dtype=np.uint64)
            d = np.zeros(len(Q), np.uint64)
            p64 = C.POINTER(C.c_uint64)
            assert lattigpu.lib().lg_bfv_lift_params_host(q.ctypes.data_as(p64), len(Q), t, d.ctypes.data_as(p64)) == 0
            assert [int(x) for x in d] == [((prod(Q) // t) % x << 64) % x for x in Q]
    assert np.array_equal(lattigpu.bfv_scheme.index_matrix(256), orc.bfv_index_matrix(256))


def test_general_degree_tensor_matches_degree1_restatement():
    """The general-degree restatement of tensorAndRescale (bfv/evaluator.go:374-417, composed from the oracle's ring ops)
    must return the degree-1 x degree-1 restatement's words on degree-1 operands: both branches compute the same
    canonical products, only the schedule of reductions differs."""
    logN = 12
    Q, P, M = orc.gen_moduli(logN, [39, 39], [30], [60, 60])
    N = 1 << logN
    ev = orc.BfvEvaluator(orc.Context(N, Q), orc.Context(N, M), orc.Context(N, P), 65537)
    rng = np.random.default_rng(5)
    a = np.ascontiguousarray(np.stack([rng.integers(0, q, size=(2, N), dtype=np.uint64) for q in Q], axis=1))
    b = np.ascontiguousarray(np.stack([rng.integers(0, q, size=(2, N), dtype=np.uint64) for q in Q], axis=1))
    assert np.array_equal(ev.tensor_and_rescale_general(a, b), ev.tensor_and_rescale(a, b))
    assert np.array_equal(ev.tensor_and_rescale_general(a, a, True), ev.tensor_and_rescale(a, a))
    # plaintext x ciphertext: linear in the ciphertext components
    pt = np.ascontiguousarray(a[:1])
    out = ev.tensor_and_rescale_general(b, pt)
    assert out.shape == (2, len(Q), N)
