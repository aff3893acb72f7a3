"""Wire formats of the scheme objects as the oracle restates them (ckks/marshaler.go, bfv/marshaler.go): header
bytes spelled out by hand for a tiny ring, and round trips."""
import struct

import numpy as np

from oracle import ring_oracle as orc


def _poly(rng, nl, N):
    return rng.integers(0, 1 << 64, size=(nl, N), dtype=np.uint64)


def test_ckks_ciphertext_bytes():
    rng = np.random.default_rng(1)
    N, nl = 4, 2
    c0, c1 = _poly(rng, nl, N), _poly(rng, nl, N)
    data = orc.ckks_ciphertext_marshal([c0, c1], 2.0**40, True)
    assert len(data) == 11 + 2 * (2 + 8 * N * nl)  # GetDataLen :10-20
    assert data[0] == 2 and data[1:9] == struct.pack("<d", 2.0**40) and data[9] == 0 and data[10] == 1
    assert data[11] == 2 and data[12] == nl  # log2(N), number of moduli (ring_object.go:169-170)
    assert data[13:21] == int(c0[0, 0]).to_bytes(8, "big")
    value, scale, ntt = orc.ckks_ciphertext_unmarshal(data)
    assert scale == 2.0**40 and ntt and np.array_equal(value[0], c0) and np.array_equal(value[1], c1)


def test_bfv_ciphertext_and_keys_bytes():
    rng = np.random.default_rng(2)
    N, nl = 8, 3
    polys = [_poly(rng, nl, N) for _ in range(3)]
    data = orc.bfv_ciphertext_marshal(polys, False)
    assert data[:2] == bytes([3, 0]) and len(data) == 2 + 3 * (2 + 8 * N * nl)
    value, ntt = orc.bfv_ciphertext_unmarshal(data)
    assert not ntt and all(np.array_equal(a, b) for a, b in zip(value, polys))
    pk = orc.public_key_marshal(polys[:2])
    assert pk == orc.poly_marshal(polys[0]) + orc.poly_marshal(polys[1])
    evk = rng.integers(0, 1 << 64, size=(2, 2, nl, N), dtype=np.uint64)
    sw = orc.swk_marshal(evk)
    assert sw[0] == 2 and len(sw) == 1 + 4 * (2 + 8 * N * nl)
    back, inc = orc.swk_unmarshal(sw)
    assert inc == len(sw) and np.array_equal(back, evk)
    ek = orc.bfv_evaluation_key_marshal([evk, evk])
    assert ek[0] == 2 and ek[1:] == sw + sw
    rk = orc.rotation_keys_marshal({5: evk}, {0x010203: evk}, evk)
    assert rk[:4] == bytes([2, 0, 0, 5])  # RotationLeft over the top byte of the big-endian amount
    off = 4 + len(sw)
    assert rk[off:off + 4] == bytes([1, 1, 2, 3])  # RotationRight, amount 0x010203
    off += 4 + len(sw)
    assert rk[off:off + 4] == bytes([3, 0, 0, 0]) and rk[off + 4:] == sw
