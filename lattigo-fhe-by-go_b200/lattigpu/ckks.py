"""Host-side mirror of the hot ops of the reference's ckks.evaluator
(ckks/evaluator.go:933-1591) over the C ABI: MulRelin, Relinearize, Rescale,
SwitchKeys, RotateColumns/Conjugate with a direct key, and the key-switch core.

A ciphertext is the pair (value[0], value[1]) of device Polys over Q (NTT
domain); scale / isNTT metadata stay with the caller as in ckks/operand.go.
"""
import ctypes as C

import numpy as np

from ._lib import check, lib, vp
from .ring import Poly, _arr, _ptr, _s


# ckks/params.go:36-87 DefaultParams: (LogN, LogQi, LogPi, Scale)
PN12QP109, PN13QP218, PN14QP438, PN15QP880, PN16QP1761 = range(5)
DefaultParams = [
    dict(LogN=12, LogQi=[37, 32], LogPi=[38], Scale=float(1 << 32)),
    dict(LogN=13, LogQi=[33, 30, 30, 30, 30, 30], LogPi=[35], Scale=float(1 << 30)),
    dict(LogN=14, LogQi=[45] + [34] * 9, LogPi=[43, 43], Scale=float(1 << 34)),
    dict(LogN=15, LogQi=[50] + [40] * 17, LogPi=[50, 50, 50], Scale=float(1 << 40)),
    dict(LogN=16, LogQi=[55] + [45] * 33, LogPi=[55, 55, 55, 55], Scale=float(1 << 45)),
]


def GenModuli(params):
    """ckks/utils.go:150-193: primes are generated per bit size (GenerateNTTPrimes) and
    dealt in order to Q, then P, so that all moduli are distinct.  Host only."""
    from .ring import GenerateNTTPrimes

    need = {}
    for b in list(params["LogQi"]) + list(params["LogPi"]):
        if b > 60:
            raise ValueError("cannot GenModuli: the provided LogQi/LogPi must be smaller than 61")
        need[b] = need.get(b, 0) + 1
    primes = {b: GenerateNTTPrimes(b, params["LogN"], n) for b, n in need.items()}
    Q, P = [], []
    for b in params["LogQi"]:
        Q.append(primes[b].pop(0))
    for b in params["LogPi"]:
        P.append(primes[b].pop(0))
    return Q, P


def scaleUpExact(value, n, q):
    """ckks/utils.go:22-49.  big.NewFloat(n*value) carries 53 bits, so adding 0.5 rounds like float64
    addition; Int() truncates.  A negative value whose magnitude is 0 mod q yields q (kept literal)."""
    x = -n * value if value < 0 else n * value
    res = int(x + 0.5) % q
    return q - res if value < 0 else res


def _mred(x, y, q, qinv):  # ring/modular_reduction.go:70-79
    M = (1 << 64) - 1
    t = x * y
    h = (((t & M) * qinv) & M) * q >> 64
    r = ((t >> 64) - h + q) & M
    return r - q if r >= q else r


def _mform(a, q, u0, u1):  # ring/modular_reduction.go:15-22
    M = (1 << 64) - 1
    r = (-((a * u0 + ((a * u1) >> 64)) & M) * q) & M
    return r - q if r >= q else r


class SwitchingKey:
    """ckks.SwitchingKey.evakey: [beta][2] polys over QP, NTT + Montgomery form
    (ckks/keygen.go:282-340), uploaded once and shared by every ciphertext of a batch."""

    def __init__(self, evakey=None, N=None, device_ptr=None, beta=None, nQP=None, keep=None):
        h = vp()
        if evakey is not None:
            a = _arr(evakey)
            assert a.ndim == 4 and a.shape[1] == 2, "evakey must be [beta][2][nQ+nP][N]"
            self.beta, self.nQP, self.N = a.shape[0], a.shape[2], a.shape[3]
            check(lib().lg_swk_create(self.N, self.beta, self.nQP, _ptr(a), C.byref(h)))
        else:
            self.beta, self.nQP, self.N = beta, nQP, N
            check(lib().lg_swk_wrap(vp(device_ptr), N, beta, nQP, C.byref(h)))
        self.h = h
        self._keep = keep

    def __del__(self):
        try:
            lib().lg_swk_destroy(self.h)
        except Exception:
            pass


class Evaluator:
    """ring part of ckks.NewEvaluator (ckks/evaluator.go:81-112): contexts Q and P,
    the FastBasisExtender and the Decomposer; scratch is stream-ordered."""

    def __init__(self, contextQ, contextP):
        self.contextQ, self.contextP = contextQ, contextP
        h = vp()
        check(lib().lg_ckks_eval_create(contextQ.h, contextP.h, C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            lib().lg_ckks_eval_destroy(self.h)
        except Exception:
            pass

    def switchKeysInPlace(self, level, cx, evakey, p0, p1, stream=None):
        check(lib().lg_ckks_switch_keys_in_place(self.h, level, cx.h, evakey.h, p0.h, p1.h, _s(stream)))

    def MulRelin(self, level, ct0, ct1, evakey, ctOut, stream=None):
        """MulRelin (:1016-1133).  An operand is its tuple of value polys: (v0, v1) a ciphertext, (v0,) a plaintext.
        ct0 is ct1 selects the squaring branch.  Ciphertext x ciphertext with a key is the fused device path; with
        evakey None the degree-2 result goes to the three polys of ctOut (:1059-1063, :1111-1117); plaintext x
        ciphertext (either order) is the :1121-1137 branch.  The last two are the reference's ring-op sequences."""
        K = self.contextQ
        if len(ct0) == 2 and len(ct1) == 2 and evakey is not None:
            check(lib().lg_ckks_mul_relin(self.h, level, ct0[0].h, ct0[1].h, ct1[0].h, ct1[1].h, evakey.h, ctOut[0].h,
                                          ctOut[1].h, _s(stream)))
            return
        batch = ct0[0].batch
        if len(ct0) == 2 and len(ct1) == 2:
            if len(ctOut) != 3:
                raise ValueError("cannot MulRelin: a degree 2 receiver is needed when no evaluation key is given")
            c00, c01 = K.NewPolyLvl(level, batch), K.NewPolyLvl(level, batch)  # ringpool[0], ringpool[1]
            alias = any(o is x for o in ctOut for x in tuple(ct0) + tuple(ct1))
            c0, c1, c2 = [K.NewPolyLvl(level, batch) for _ in range(3)] if alias else ctOut
            K.MFormLvl(level, ct0[0], c00, stream=stream)  # :1080-1081
            K.MFormLvl(level, ct0[1], c01, stream=stream)
            K.MulCoeffsMontgomeryLvl(level, c00, ct1[0], c0, stream=stream)
            K.MulCoeffsMontgomeryLvl(level, c00, ct1[1], c1, stream=stream)
            if ct0 is ct1:  # :1083-1088
                K.AddLvl(level, c1, c1, c1, stream=stream)
            else:  # :1092-1095
                K.MulCoeffsMontgomeryAndAddLvl(level, c01, ct1[0], c1, stream=stream)
            K.MulCoeffsMontgomeryLvl(level, c01, ct1[1], c2, stream=stream)
            if alias:  # :1111-1116
                for src, dst in zip((c0, c1, c2), ctOut):
                    K.CopyLvl(level, src, dst, stream=stream)
            return
        if len(ct0) + len(ct1) != 3:
            raise ValueError("cannot MulRelin: input elements must be of degree 0 or 1")
        tmp0, tmp1 = (ct1, ct0) if len(ct0) == 2 else (ct0, ct1)  # :1125-1129
        c00 = K.NewPolyLvl(level, batch)
        K.MFormLvl(level, tmp0[0], c00, stream=stream)  # :1134
        K.MulCoeffsMontgomeryLvl(level, c00, tmp1[0], ctOut[0], stream=stream)
        K.MulCoeffsMontgomeryLvl(level, c00, tmp1[1], ctOut[1], stream=stream)

    def Relinearize(self, level, ct0, evakey, ctOut, stream=None):
        check(lib().lg_ckks_relinearize(self.h, level, ct0[0].h, ct0[1].h, ct0[2].h, evakey.h, ctOut[0].h, ctOut[1].h,
                                        _s(stream)))

    def Rescale(self, nl, ct, nb=1, stream=None):
        """ckks/evaluator.go:955-960 applied nb times; the result lives in the first nl-nb limbs."""
        check(lib().lg_ckks_rescale(self.h, nl, ct[0].h, ct[1].h, nb, _s(stream)))

    def SwitchKeys(self, level, ct0, switchingKey, ctOut, stream=None):
        check(lib().lg_ckks_switch_keys(self.h, level, ct0[0].h, ct0[1].h, switchingKey.h, ctOut[0].h, ctOut[1].h, _s(stream)))

    def permuteNTT(self, level, ct0, index, evakey, ctOut, stream=None):
        """RotateColumns with a direct key (:1220) / Conjugate (:1449)"""
        check(lib().lg_ckks_permute_ntt(self.h, level, ct0[0].h, ct0[1].h, index.h, evakey.h, ctOut[0].h, ctOut[1].h,
                                        _s(stream)))


    # ---- constant ops (ckks/evaluator.go:373-833): the per-limb scalars are host arithmetic as in the
    # reference, the coefficient loops run on the device; scale / level bookkeeping stays with the caller
    def _tables(self):
        if not hasattr(self, "_tb"):
            t = self.contextQ.tables()
            self._tb = dict(q=[int(x) for x in self.contextQ.Modulus], qinv=[int(x) for x in t["mred"]],
                            bred=[(int(a), int(b)) for a, b in t["bred"]], psi2=[int(p[1]) for p in t["psi"]])
        return self._tb

    @staticmethod
    def _parts(constant):
        c = complex(constant)
        return float(c.real), float(c.imag)

    def _const_halves(self, level, cReal, cImag, scale, mont):
        """[a + b*psi^2] for the first N/2 NTT coefficients, [a - b*psi^2] for the rest (:413-444, :560-609, :684-727)"""
        tb = self._tables()
        lo, hi = [], []
        for i in range(level + 1):
            q, qinv, (u0, u1) = tb["q"][i], tb["qinv"][i], tb["bred"][i]
            re = im = sc = 0
            if cReal != 0:
                re = scaleUpExact(cReal, scale, q)
                sc = re
            if cImag != 0:
                im = _mred(scaleUpExact(cImag, scale, q), tb["psi2"][i], q, qinv)
                t = sc + im
                sc = t - q if t >= q else t
            first = _mform(sc, q, u0, u1) if mont else sc
            second = first
            if cImag != 0:
                t = re + (q - im)
                t = t - q if t >= q else t
                second = _mform(t, q, u0, u1) if mont else t
            lo.append(first)
            hi.append(second)
        return _arr(lo), _arr(hi)

    def AddConst(self, level, ct0, constant, scale, ctOut, stream=None):
        """:373-448 -- `scale` = ct0.Scale(); only value[0] is touched"""
        cReal, cImag = self._parts(constant)
        lo, hi = self._const_halves(level, cReal, cImag, scale, False)
        check(lib().lg_ring_add_scalar_halves(self.contextQ.h, level + 1, ct0[0].h, _ptr(lo), _ptr(hi), ctOut[0].h, _s(stream)))

    def const_scale(self, constant, default_scale):
        """the scaling MultByConst / MultByConstAndAdd apply to a constant with a fractional part (:631-668)"""
        cReal, cImag = self._parts(constant)
        scale = 1.0
        if isinstance(constant, (complex, float)):
            for v in (cReal, cImag):
                if v != 0 and v - float(int(v)) != 0:
                    scale = default_scale
        return scale

    def MultByConst(self, level, ct0, constant, scale, ctOut, stream=None):
        """:622-730 -- `scale` as returned by const_scale(); every value[u] is multiplied"""
        cReal, cImag = self._parts(constant)
        lo, hi = self._const_halves(level, cReal, cImag, scale, True)
        for a, c in zip(ct0, ctOut):
            check(lib().lg_ring_mul_scalar_montgomery_halves(self.contextQ.h, level + 1, a.h, _ptr(lo), _ptr(hi), c.h, _s(stream)))

    def MultByConstAndAdd(self, level, ct0, constant, scale, ctOut, stream=None):
        """:451-610 inner loops (:560-609) -- `scale` is the factor the reference derives from the two scales"""
        cReal, cImag = self._parts(constant)
        lo, hi = self._const_halves(level, cReal, cImag, scale, True)
        for a, c in zip(ct0, ctOut):
            check(lib().lg_ring_mul_scalar_montgomery_halves_and_add(self.contextQ.h, level + 1, a.h, _ptr(lo), _ptr(hi), c.h,
                                                                     _s(stream)))

    def _by_i(self, level, ct0, ctOut, div, stream):
        tb = self._tables()
        psi2 = [tb["psi2"][i] for i in range(level + 1)]
        neg = [tb["q"][i] - psi2[i] for i in range(level + 1)]
        lo, hi = (_arr(neg), _arr(psi2)) if div else (_arr(psi2), _arr(neg))
        for a, c in zip(ct0, ctOut):
            check(lib().lg_ring_mul_scalar_montgomery_halves(self.contextQ.h, level + 1, a.h, _ptr(lo), _ptr(hi), c.h, _s(stream)))

    def MultByi(self, level, ct0, ctOut, stream=None):
        """:746-784: multiplication by X^(N/2)"""
        self._by_i(level, ct0, ctOut, False, stream)

    def DivByi(self, level, ct0, ctOut, stream=None):
        """:795-833: multiplication by X^(3N/2)"""
        self._by_i(level, ct0, ctOut, True, stream)

    def RotateHoisted(self, level, ct0, rotations, ctOuts, stream=None):
        """RotateHoisted (:1252-1289): `rotations` is a list of (index, rotation key) pairs, i.e.
        (permuteNTTLeftIndex[k], evakeyRotColLeft[k]); ctOuts[i] receives ct0 rotated by rotations[i].
        The decomposition of ct0.value[1] is computed once and shared (switchKeyHoisted :1291-1392)."""
        h = vp()
        check(lib().lg_ckks_hoist(self.h, level, ct0[1].h, C.byref(h), _s(stream)))
        try:
            for (index, key), out in zip(rotations, ctOuts):
                check(lib().lg_ckks_switch_key_hoisted(self.h, h, ct0[0].h, index.h, key.h, out[0].h, out[1].h, _s(stream)))
        finally:
            lib().lg_hoisted_destroy(h)  # stream-ordered release: after the rotations issued on this stream

    # ---- drivers around permuteNTT / DivRoundByLastModulus*: host logic of the reference, device ops through the ABI
    def RotateColumns(self, level, ct0, k, evakey, ctOut, stream=None):
        """RotateColumns (:1201-1248): the key of rotation k when it exists, otherwise the power-of-two decomposition over
        the left or the right keys, whichever needs fewer rotations; `evakey` is a RotationKeys."""
        N = self.contextQ.N
        k &= (N >> 1) - 1  # :1207
        if k == 0:  # :1209-1211
            for a, c in zip(ct0, ctOut):
                self.contextQ.CopyLvl(level, a, c, stream=stream)
            return
        if evakey.evakeyRotColLeft.get(k) is not None:  # :1218-1220
            self.permuteNTT(level, ct0, evakey.permuteNTTLeftIndex[k], evakey.evakeyRotColLeft[k], ctOut, stream=stream)
            return
        i, has = 1, True  # :1225-1231
        while i < (N >> 1):
            if evakey.evakeyRotColLeft.get(i) is None or evakey.evakeyRotColRight.get(i) is None:
                has = False
                break
            i <<= 1
        if not has:
            raise ValueError("cannot RotateColumns: specific rotation and pow2 rotations have not been generated")  # :1245
        if bin(k).count("1") <= bin((N >> 1) - k).count("1"):  # :1236-1240
            self.rotateColumnsPow2(level, ct0, k, evakey.permuteNTTLeftIndex, evakey.evakeyRotColLeft, ctOut, stream=stream)
        else:
            self.rotateColumnsPow2(level, ct0, (N >> 1) - k, evakey.permuteNTTRightIndex, evakey.evakeyRotColRight, ctOut,
                                   stream=stream)

    def rotateColumnsPow2(self, level, ct0, k, permuteNTTIndex, evakeyRotCol, ctOut, stream=None):
        """:1402-1424: copy, then one in-place permuteNTT per set bit of k"""
        for a, c in zip(ct0, ctOut):
            self.contextQ.CopyLvl(level, a, c, stream=stream)
        evakeyIndex = 1
        while k > 0:
            if k & 1:
                self.permuteNTT(level, ctOut, permuteNTTIndex[evakeyIndex], evakeyRotCol[evakeyIndex], ctOut, stream=stream)
            evakeyIndex <<= 1
            k >>= 1

    def Conjugate(self, level, ct0, evakey, ctOut, stream=None):
        """:1437-1450"""
        if evakey.evakeyConjugate is None:
            raise ValueError("cannot Conjugate: rows rotation key not generated")
        self.permuteNTT(level, ct0, evakey.permuteNTTConjugateIndex, evakey.evakeyConjugate, ctOut, stream=stream)

    def RescaleMany(self, nl, ct, nbRescales, stream=None):
        """RescaleMany (:971-1000): DivRoundByLastModulusManyNTT on every value; the result lives in the first
        nl - nbRescales limbs.  Returns the factor the scale is divided by."""
        if nl - 1 < nbRescales:
            raise ValueError("cannot RescaleMany: input Ciphertext level too low")  # :974
        div = 1.0
        for i in range(nbRescales):  # :987-989
            div *= float(self.contextQ.Modulus[nl - 1 - i])
        for v in ct:  # :991-993
            self.contextQ.DivRoundByLastModulusManyNTT(v, nbRescales, nl=nl, stream=stream)
        return div

    def RescaleThreshold(self, nl, ct, scale, threshold, stream=None):
        """Rescale (:933-968): divide by the last modulus while scale >= threshold * q_level / 2; returns (scale, nl)"""
        if nl - 1 == 0:
            raise ValueError("cannot Rescale: input Ciphertext already at level 0")  # :938
        q = self.contextQ.Modulus
        while scale >= (threshold * float(q[nl - 1])) / 2 and nl - 1 != 0:  # :955
            scale /= float(q[nl - 1])
            for v in ct:
                self.contextQ.DivRoundByLastModulusNTT(v, nl=nl, stream=stream)
            nl -= 1
        return scale, nl


GaloisGen = 5  # ckks/ckks.go:13


class RotationKeys:
    """ckks.RotationKeys (ckks/keygen.go:24-33): switching keys of the column rotations and of the conjugation with the
    index tables permuteNTT gathers through.  SetRotKey mirrors keygen.go:418-478 (including its right-rotation index
    exponent 2N-1-k, where GenRot's power-of-two set uses 2N-n, :410)."""
    RotationLeft, RotationRight, Conjugate = 0, 1, 2

    def __init__(self, N):
        self.N = N
        self.evakeyRotColLeft, self.evakeyRotColRight = {}, {}
        self.permuteNTTLeftIndex, self.permuteNTTRightIndex = {}, {}
        self.evakeyConjugate, self.permuteNTTConjugateIndex = None, None

    def SetRotKey(self, key, rotType, k=0):
        from . import ring

        N = self.N
        if rotType == self.RotationLeft:
            if self.evakeyRotColLeft.get(k) is None and k != 0:
                self.permuteNTTLeftIndex[k] = ring.PermuteNTTIndex(GaloisGen, k, N)
                self.evakeyRotColLeft[k] = key
        elif rotType == self.RotationRight:
            if self.evakeyRotColRight.get(k) is None and k != 0:
                self.permuteNTTRightIndex[k] = ring.PermuteNTTIndex(GaloisGen, 2 * N - 1 - k, N)
                self.evakeyRotColRight[k] = key
        else:
            if self.evakeyConjugate is None:
                self.permuteNTTConjugateIndex = ring.PermuteNTTIndex(2 * N - 1, 1, N)
                self.evakeyConjugate = key

    def SetPow2(self, n, left_key, right_key):
        """one entry of GenRot's power-of-two set (keygen.go:405-413)"""
        from . import ring

        self.permuteNTTLeftIndex[n] = ring.PermuteNTTIndex(GaloisGen, n, self.N)
        self.permuteNTTRightIndex[n] = ring.PermuteNTTIndex(GaloisGen, 2 * self.N - n, self.N)
        self.evakeyRotColLeft[n], self.evakeyRotColRight[n] = left_key, right_key


def NewEvaluator(contextQ, contextP):
    return Evaluator(contextQ, contextP)
