"""GPU parity of the scheme-object wire formats (ckks/marshaler.go, bfv/marshaler.go): the bytes produced from
device-resident objects equal the oracle's, unmarshalling restores the same device words, and an unmarshalled
switching key drives the key switch to the same result."""
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


def test_ciphertexts_and_keys(lg):
    M, R = lg.marshaler, lg.ring
    rng = np.random.default_rng(3)
    N, nl = 4096, 4
    polys = [rng.integers(0, 1 << 64, size=(nl, N), dtype=np.uint64) for _ in range(3)]
    dev = [R.Poly.from_numpy(p) for p in polys]
    # CKKS: degree 1, level 2 (one limb dropped: metadata only on the device)
    ct = M.CkksCiphertext(dev[:2], scale=2.0**37 + 0.5, isNTT=True, level=2)
    data = ct.MarshalBinary()
    assert data == orc.ckks_ciphertext_marshal([p[:3] for p in polys[:2]], 2.0**37 + 0.5, True)
    assert ct.GetDataLen(True) == len(data) and ct.GetDataLen(False) == len(data) - 11 - 2 * 2
    back = M.CkksCiphertext().UnmarshalBinary(data)
    assert back.scale == 2.0**37 + 0.5 and back.isNTT and back.Degree() == 1 and back.Level() == 2
    assert all(np.array_equal(b.numpy(), p[:3]) for b, p in zip(back.value, polys))
    # BFV: degree 2
    bt = M.BfvCiphertext(dev, isNTT=False)
    data = bt.MarshalBinary()
    assert data == orc.bfv_ciphertext_marshal(polys, False) and bt.GetDataLen(True) == len(data)
    back = M.BfvCiphertext().UnmarshalBinary(data)
    assert not back.isNTT and all(np.array_equal(b.numpy(), p) for b, p in zip(back.value, polys))
    # keys
    sk = M.SecretKey(dev[0])
    assert sk.MarshalBinary() == orc.poly_marshal(polys[0]) and sk.GetDataLen(True) == 2 + 8 * N * nl
    assert np.array_equal(M.SecretKey().UnmarshalBinary(sk.MarshalBinary()).sk.numpy(), polys[0])
    pk = M.PublicKey(dev[:2])
    assert pk.MarshalBinary() == orc.public_key_marshal(polys[:2])
    b = M.PublicKey().UnmarshalBinary(pk.MarshalBinary())
    assert np.array_equal(b.pk[0].numpy(), polys[0]) and np.array_equal(b.pk[1].numpy(), polys[1])
    # truncated input
    with pytest.raises((ValueError, lg.LattigpuError)):
        M.CkksCiphertext().UnmarshalBinary(data[:100])


def test_switching_evaluation_rotation_keys(lg):
    M = lg.marshaler
    logN, lq, lp = 12, [45, 40, 40, 40], [50, 50]
    N = 1 << logN
    Q, P, _ = orc.gen_moduli(logN, lq, lp)
    nQ, nP = len(Q), len(P)
    beta = -(-nQ // nP)
    rng = np.random.default_rng(4)
    mk = lambda: np.ascontiguousarray(np.stack([rng.integers(0, q, size=(beta, 2, N), dtype=np.uint64) for q in Q + P], axis=2))
    evk, evk2, evk3 = mk(), mk(), mk()
    key = lg.ckks.SwitchingKey(evk)
    data = key.MarshalBinary()
    assert data == orc.swk_marshal(evk) and key.GetDataLen(True) == len(data)
    key_b = lg.ckks.SwitchingKey.UnmarshalBinary(data)
    assert (key_b.beta, key_b.nQP, key_b.N) == (beta, nQ + nP, N)
    assert key_b.MarshalBinary() == data
    # the decoded key is a working key: same key switch as the original, and as the oracle
    cQ, cP = lg.ring.NewContextWithParams(N, Q), lg.ring.NewContextWithParams(N, P)
    ev = lg.ckks.NewEvaluator(cQ, cP)
    cx = np.ascontiguousarray(np.stack([rng.integers(0, q, size=N, dtype=np.uint64) for q in Q]))
    outs = []
    for k in (key, key_b):
        p0, p1 = cQ.NewPoly(), cQ.NewPoly()
        ev.switchKeysInPlace(nQ - 1, lg.ring.Poly.from_numpy(cx), k, p0, p1)
        outs.append((p0.numpy(), p1.numpy()))
    assert np.array_equal(outs[0][0], outs[1][0]) and np.array_equal(outs[0][1], outs[1][1])
    oev = orc.CkksEvaluator(orc.Context(N, Q), orc.Context(N, P))
    w0, w1 = oev.switch_keys_in_place(nQ - 1, cx, evk)
    assert np.array_equal(outs[1][0], w0) and np.array_equal(outs[1][1], w1)
    # evaluation keys
    ek = M.CkksEvaluationKey(key)
    assert ek.MarshalBinary() == data and M.CkksEvaluationKey().UnmarshalBinary(data).evakey.MarshalBinary() == data
    key2 = lg.ckks.SwitchingKey(evk2)
    bek = M.BfvEvaluationKey([key, key2])
    bdata = bek.MarshalBinary()
    assert bdata == orc.bfv_evaluation_key_marshal([evk, evk2]) and bek.GetDataLen(True) == len(bdata)
    bb = M.BfvEvaluationKey().UnmarshalBinary(bdata)
    assert len(bb.evakey) == 2 and bb.evakey[1].MarshalBinary() == orc.swk_marshal(evk2)
    # rotation keys: left 3 and 70000, right 1, conjugate / row
    rk = M.RotationKeys()
    rk.evakeyRotColLeft[3] = key
    rk.evakeyRotColLeft[70000] = key2
    rk.evakeyRotColRight[1] = key2
    rk.evakeyConjugate = lg.ckks.SwitchingKey(evk3)
    rdata = rk.MarshalBinary()
    assert rdata == orc.rotation_keys_marshal({3: evk, 70000: evk2}, {1: evk2}, evk3)
    assert rk.GetDataLen(True) == len(rdata)
    rb = M.RotationKeys().UnmarshalBinary(rdata)
    assert sorted(rb.evakeyRotColLeft) == [3, 70000] and sorted(rb.evakeyRotColRight) == [1]
    assert rb.evakeyRotColLeft[70000].MarshalBinary() == orc.swk_marshal(evk2)
    assert rb.evakeyRotRow.MarshalBinary() == orc.swk_marshal(evk3)
