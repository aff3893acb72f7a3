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
        """ct = (value0, value1).  ct0 is ct1 selects the squaring branch."""
        check(lib().lg_ckks_mul_relin(self.h, level, ct0[0].h, ct0[1].h, ct1[0].h, ct1[1].h, evakey.h, ctOut[0].h, ctOut[1].h,
                                      _s(stream)))

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
            lib().lg_stream_sync(_s(stream))  # the decomposition must outlive the kernels that read it
            lib().lg_hoisted_destroy(h)


def NewEvaluator(contextQ, contextP):
    return Evaluator(contextQ, contextP)
