"""utils.PRNG / ring.CRPGenerator (utils/prng.go, ring/prng.go).

CPU part: the oracle against the RFC 7693 known answer and the reference's own PRNG test
(utils/prng_test.go:8-40: two generators with the same key / seed / clock agree), and the library's
host-only hash chain against the oracle (no kernel is launched: the chain is host code, csrc/crp.cu).
GPU part: CRPGenerator.Clock into a device polynomial, bit-exact against the oracle.
"""
import hashlib

import numpy as np
import pytest

from oracle import prng_oracle as po

KEY = bytes([0x49, 0x0a, 0x42, 0x3d, 0x97, 0x9d, 0xc1, 0x07, 0xa1, 0xd7, 0xe9, 0x7b, 0x3b, 0xce, 0xa1, 0xdb,
             0x42, 0xf3, 0xa6, 0xd5, 0x75, 0xd2, 0x0c, 0x92, 0xb7, 0x35, 0xce, 0x0c, 0xee, 0x09, 0x7c, 0x98])
SEED = bytes([0x48, 0xc3, 0x31, 0x12, 0x74, 0x98, 0xd3, 0xf2, 0x7b, 0x15, 0x15, 0x9b, 0x50, 0xc4, 0x9c, 0x00,
              0x7d, 0xa5, 0xea, 0x68, 0x1f, 0xed, 0x4f, 0x99, 0x54, 0xc0, 0x52, 0xc0, 0x75, 0xff, 0xf7, 0x5c])
RFC7693_ABC = ("ba80a53f981c4d0d6a2797b69f12f6e94c212f14685ac4b74b12bb6fdbffa2d1"
               "7d87c5392aab792dc252d5de4533cc9518d38aa8dbf1925ab92386edd4009923")


def test_oracle_hash_known_answer():
    assert hashlib.blake2b(b"abc", digest_size=64).hexdigest() == RFC7693_ABC


def test_oracle_prng_reference_test():
    """utils/prng_test.go:8-40"""
    Ha, Hb = po.PRNG(KEY), po.PRNG(KEY)
    Ha.Seed(SEED)
    Hb.Seed(SEED)
    Ha.SetClock(256)
    Hb.SetClock(256)
    assert Ha.Clock() == Hb.Clock() and Ha.clock == 257
    with pytest.raises(ValueError):
        Ha.SetClock(3)
    # digest k is the keyed hash of seed || d_1 || ... || d_(k-1)
    H = po.PRNG(KEY)
    H.Seed(SEED)
    d1 = H.Clock()
    d2 = H.Clock()
    assert d1 == hashlib.blake2b(SEED, key=KEY, digest_size=64).digest()
    assert d2 == hashlib.blake2b(SEED + d1, key=KEY, digest_size=64).digest()


def test_oracle_crp_is_uniform_below_q():
    g = po.CRPGenerator(None, 64, [576460752303439873, 1099511480321, 97])
    g.Seed(b"")
    a = g.Clock()
    b = g.Clock()
    assert a.shape == (3, 64) and (a != b).any()
    for j, q in enumerate(g.moduli):
        assert (a[j] < q).all()
    g2 = po.CRPGenerator(None, 64, g.moduli)
    g2.Seed(b"")
    g2.SetClock(0)
    assert (g2.Clock() == a).all()


@pytest.mark.parametrize("key,seed", [(None, None), (None, b""), (KEY, SEED), (b"k" * 64, b"s" * 300), (b"\x01", b"x" * 128)])
def test_library_prng_matches_oracle(key, seed):
    """host-only entry points of the C ABI (no device needed, no kernel launched)"""
    from lattigpu import ring as gring

    G, O = gring.NewPRNG(key), po.PRNG(key)
    if seed is not None:
        G.Seed(seed)
        O.Seed(seed)
    for _ in range(9):
        assert G.Clock() == O.Clock()
    G.SetClock(200)
    O.SetClock(200)
    assert G.GetClock() == 200 and G.Clock() == O.Clock()
    from lattigpu import LattigpuError

    with pytest.raises(LattigpuError, match="previous state"):
        G.SetClock(5)


def test_library_prng_rejects_long_key():
    from lattigpu import LattigpuError
    from lattigpu import ring as gring

    with pytest.raises(LattigpuError, match="invalid key size"):
        gring.NewPRNG(b"k" * 65)


@pytest.mark.gpu
@pytest.mark.parametrize("N,moduli,key", [
    (256, [576460752303439873, 576460752303702017], None),
    (1024, [1152921504606748673, 35184372121601, 8796093202433, 1099511480321], KEY),
])
def test_gpu_crp_generator(N, moduli, key):
    from lattigpu import ring as gring

    ctx = gring.NewContextWithParams(N, moduli)
    G = gring.NewCRPGenerator(key, ctx)
    O = po.CRPGenerator(key, N, moduli)
    G.Seed(SEED)
    O.Seed(SEED)
    for _ in range(2):
        crp = G.ClockNew()
        assert (crp.numpy() == O.Clock()).all()
        assert G.GetClock() == O.GetClock()
    # batched handle: entry 1 only
    p = ctx.NewPoly(batch=3)
    p.set(np.full((3, len(moduli), N), 7, dtype=np.uint64))
    before = p.numpy().copy()
    G.Clock(p, batch_index=1)
    got = p.numpy()
    assert (got[1] == O.Clock()).all() and (got[0] == before[0]).all() and (got[2] == before[2]).all()
    G.SetClock(G.GetClock() + 7)
    O.SetClock(O.GetClock() + 7)
    assert (G.ClockNew().numpy() == O.Clock()).all()
