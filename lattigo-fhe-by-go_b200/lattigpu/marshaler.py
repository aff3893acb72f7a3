"""Wire formats of the scheme objects (ckks/marshaler.go, bfv/marshaler.go) over device-resident polynomials.

Every object is a short header around ring.Poly encodings (ring/ring_object.go:146-289: two header bytes, then
the coefficients limb-major as big-endian 64-bit words); the byte swap runs on the device (lg_poly_write_to /
lg_poly_decode), so ciphertexts, keys and shares stream between HBM and the wire without a host-side pass over
the coefficients.  Method names, header layouts and error behaviour follow the reference.
"""
import ctypes as C
import struct

from . import ring
from ._lib import check, lib, vp
from .ckks import SwitchingKey

# rotation types: ckks/keygen.go:45-49, bfv/keygen.go:41-45
RotationRight, RotationLeft, Conjugate = 1, 2, 3
RotationRow = 3


def _poly_len(data, pointer):
    """bytes taken by the ring.Poly encoded at data[pointer:] (DecodePolyNew, ring_object.go:277-289)"""
    if len(data) < pointer + 2:
        raise ValueError("error : invalid polynomial encoding")
    return 2 + (((1 << data[pointer]) * data[pointer + 1]) << 3)


def decode_poly_new(data, pointer=0, stream=None):
    """Poly.DecodePolyNew: a new device polynomial from data[pointer:], and the number of bytes read"""
    inc = _poly_len(data, pointer)
    if len(data) < pointer + inc:
        raise ValueError("error : invalid polynomial encoding")
    p = ring.Poly(1 << data[pointer], max(1, data[pointer + 1]), 1)
    p.UnmarshalBinary(bytes(data[pointer:pointer + inc]), stream=stream)
    return p, inc


class _PolyList:
    """shared by the ciphertexts and the public key: a sequence of ring.Poly encodings after `header`"""

    def _polys(self):
        raise NotImplementedError

    def _header(self):
        return b""

    def GetDataLen(self, WithMetaData=True):
        n = len(self._header()) if WithMetaData else 0
        return n + sum(p.GetDataLen(WithMetaData, nl) for p, nl in self._polys())

    def MarshalBinary(self, stream=None):
        return self._header() + b"".join(p.MarshalBinary(nl=nl, stream=stream) for p, nl in self._polys())


class CkksCiphertext(_PolyList):
    """ckks.Ciphertext (ckks/marshaler.go:9-91): data[0] = degree + 1, data[1:9] = scale (float64 bits, little
    endian), data[9] unused, data[10] = isNTT, then the polynomials at the ciphertext's level"""

    def __init__(self, value=None, scale=0.0, isNTT=True, level=None):
        self.value, self.scale, self.isNTT = list(value or []), float(scale), bool(isNTT)
        self.level = level

    def Degree(self):
        return len(self.value) - 1

    def Level(self):
        return self.value[0].nlimbs - 1 if self.level is None else self.level

    def _polys(self):
        return [(p, self.Level() + 1) for p in self.value]

    def _header(self):
        return bytes([self.Degree() + 1]) + struct.pack("<d", self.scale) + bytes([0, 1 if self.isNTT else 0])

    def UnmarshalBinary(self, data, stream=None):
        data = bytes(data)
        n = data[0]
        self.scale = struct.unpack("<d", data[1:9])[0]
        self.isNTT = data[10] == 1
        self.value, self.level = [], None
        pointer = 11
        for _ in range(n):
            p, inc = decode_poly_new(data, pointer, stream=stream)
            self.value.append(p)
            pointer += inc
        return self


class BfvCiphertext(_PolyList):
    """bfv.Ciphertext (bfv/marshaler.go:9-73): data[0] = number of polynomials, data[1] = isNTT"""

    def __init__(self, value=None, isNTT=False):
        self.value, self.isNTT = list(value or []), bool(isNTT)

    def _polys(self):
        return [(p, p.nlimbs) for p in self.value]

    def _header(self):
        return bytes([len(self.value), 1 if self.isNTT else 0])

    def UnmarshalBinary(self, data, stream=None):
        data = bytes(data)
        self.isNTT = data[1] == 1
        self.value = []
        pointer = 2
        for _ in range(data[0]):
            p, inc = decode_poly_new(data, pointer, stream=stream)
            self.value.append(p)
            pointer += inc
        return self


class SecretKey(_PolyList):
    """ckks/marshaler.go:93-120, bfv/marshaler.go:75-103: the polynomial alone"""

    def __init__(self, sk=None):
        self.sk = sk

    def _polys(self):
        return [(self.sk, self.sk.nlimbs)]

    def UnmarshalBinary(self, data, stream=None):
        self.sk, _ = decode_poly_new(bytes(data), 0, stream=stream)
        return self


class PublicKey(_PolyList):
    """ckks/marshaler.go:122-162, bfv/marshaler.go:105-150: pk[0] then pk[1]"""

    def __init__(self, pk=None):
        self.pk = list(pk) if pk is not None else [None, None]

    def _polys(self):
        return [(p, p.nlimbs) for p in self.pk]

    def UnmarshalBinary(self, data, stream=None):
        data = bytes(data)
        self.pk[0], inc = decode_poly_new(data, 0, stream=stream)
        self.pk[1], _ = decode_poly_new(data, inc, stream=stream)
        return self


def _swk_poly(key, digit, half):
    h = vp()
    check(lib().lg_swk_poly(key.h, digit, half, C.byref(h)))
    return ring.Poly(_handle=h, _keep=key)


def swk_get_data_len(key, WithMetaData=True):
    """SwitchingKey.GetDataLen (ckks/marshaler.go:193-205)"""
    per = (2 if WithMetaData else 0) + ((key.nQP * key.N) << 3)
    return (1 if WithMetaData else 0) + 2 * key.beta * per


def swk_encode(key, stream=None):
    """SwitchingKey.encode (:230-257): data[0] = number of digits, then evakey[j][0], evakey[j][1]"""
    out = [bytes([key.beta])]
    for j in range(key.beta):
        for h in (0, 1):
            out.append(_swk_poly(key, j, h).MarshalBinary(stream=stream))
    return b"".join(out)


def swk_decode(data, pointer=0, stream=None):
    """SwitchingKey.decode (:259-283): a new device key and the number of bytes read"""
    data = bytes(data)
    beta = data[pointer]
    start = pointer
    pointer += 1
    if beta == 0:
        raise ValueError("SwitchingKey: no digits encoded")
    N, nl = 1 << data[pointer], data[pointer + 1]
    h = vp()
    check(lib().lg_swk_alloc(N, beta, nl, C.byref(h)))
    key = SwitchingKey.__new__(SwitchingKey)
    key.h, key._keep, key.beta, key.nQP, key.N = h, None, beta, nl, N
    for j in range(beta):
        for hf in (0, 1):
            inc = _poly_len(data, pointer)
            if len(data) < pointer + inc:
                raise ValueError("error : invalid polynomial encoding")
            if (1 << data[pointer], data[pointer + 1]) != (N, nl):
                raise ValueError("SwitchingKey: polynomials of different shapes")
            _swk_poly(key, j, hf).UnmarshalBinary(data[pointer:pointer + inc], stream=stream)
            pointer += inc
    return key, pointer - start


SwitchingKey.GetDataLen = swk_get_data_len
SwitchingKey.MarshalBinary = swk_encode
SwitchingKey.UnmarshalBinary = staticmethod(lambda data, stream=None: swk_decode(data, 0, stream)[0])


class CkksEvaluationKey:
    """ckks.EvaluationKey (ckks/marshaler.go:164-191): the relinearisation key alone"""

    def __init__(self, evakey=None):
        self.evakey = evakey

    def GetDataLen(self, WithMetaData=True):
        return swk_get_data_len(self.evakey, WithMetaData)

    def MarshalBinary(self, stream=None):
        return swk_encode(self.evakey, stream)

    def UnmarshalBinary(self, data, stream=None):
        self.evakey, _ = swk_decode(data, 0, stream)
        return self


class BfvEvaluationKey:
    """bfv.EvaluationKey (bfv/marshaler.go:152-200): data[0] = number of keys (maxDegree), then each key"""

    def __init__(self, evakey=None):
        self.evakey = list(evakey or [])

    def GetDataLen(self, WithMetaData=True):
        return (1 if WithMetaData else 0) + sum(swk_get_data_len(k, WithMetaData) for k in self.evakey)

    def MarshalBinary(self, stream=None):
        return bytes([len(self.evakey)]) + b"".join(swk_encode(k, stream) for k in self.evakey)

    def UnmarshalBinary(self, data, stream=None):
        data = bytes(data)
        self.evakey = []
        pointer = 1
        for _ in range(data[0]):
            k, inc = swk_decode(data, pointer, stream)
            self.evakey.append(k)
            pointer += inc
        return self


class RotationKeys:
    """ckks.RotationKeys / bfv.RotationKeys (ckks/marshaler.go:285-438, bfv/marshaler.go:289-443): per key a 4-byte
    header -- the rotation type over the top byte of the big-endian rotation amount -- then the key.  Left keys
    first, then right keys, then the conjugate (CKKS) / row (BFV) key, both of type 3."""

    def __init__(self):
        self.evakeyRotColLeft, self.evakeyRotColRight, self.evakeyThird = {}, {}, None

    # the reference names the third key evakeyConjugate (ckks) / evakeyRotRow (bfv)
    evakeyConjugate = property(lambda s: s.evakeyThird, lambda s, v: setattr(s, "evakeyThird", v))
    evakeyRotRow = evakeyConjugate

    def GetDataLen(self, WithMetaData=True):
        keys = list(self.evakeyRotColLeft.values()) + list(self.evakeyRotColRight.values())
        keys += [self.evakeyThird] if self.evakeyThird is not None else []
        return sum((4 if WithMetaData else 0) + swk_get_data_len(k, WithMetaData) for k in keys)

    def MarshalBinary(self, stream=None):
        out = []
        for typ, keys in ((RotationLeft, self.evakeyRotColLeft), (RotationRight, self.evakeyRotColRight)):
            for i, k in keys.items():
                out.append(bytes([typ]) + struct.pack(">I", i & 0xFFFFFFFF)[1:] + swk_encode(k, stream))
        if self.evakeyThird is not None:
            out.append(bytes([Conjugate, 0, 0, 0]) + swk_encode(self.evakeyThird, stream))
        return b"".join(out)

    def UnmarshalBinary(self, data, stream=None):
        data = bytes(data)
        pointer = 0
        while pointer < len(data):
            typ = data[pointer]
            number = (data[pointer + 1] << 16) | (data[pointer + 2] << 8) | data[pointer + 3]
            pointer += 4
            if typ not in (RotationLeft, RotationRight, Conjugate):
                return self  # the reference stops here and returns its nil error (ckks/marshaler.go:427-430)
            k, inc = swk_decode(data, pointer, stream)
            if typ == RotationLeft:
                self.evakeyRotColLeft[number] = k
            elif typ == RotationRight:
                self.evakeyRotColRight[number] = k
            else:
                self.evakeyThird = k
            pointer += inc
        return self
