"""Host-side mirror of the hot ops of the reference's bfv.evaluator (bfv/evaluator.go:278-813)
over the C ABI: Mul (tensorAndRescale), Relinearize, SwitchKeys, RotateColumns / RotateRows with a
direct key (permute) and the key-switch core.  Ciphertexts are tuples of device Polys over Q in the
COEFFICIENT domain, as in the reference.
"""
import ctypes as C

from ._lib import check, lib, vp
from .ckks import SwitchingKey  # same layout: [beta][2][#Q+#P][N], NTT + Montgomery (bfv/keygen.go)
from .ring import _s

# bfv/params.go:47-88 DefaultParams
PN12QP109, PN13QP218, PN14QP438, PN15QP880 = range(4)
DefaultParams = [
    dict(LogN=12, T=65537, LogQi=[39, 39], LogPi=[30], LogQiMul=[60, 60]),
    dict(LogN=13, T=65537, LogQi=[54, 54, 54], LogPi=[55], LogQiMul=[60, 60, 60]),
    dict(LogN=14, T=65537, LogQi=[56, 55, 55, 54, 54, 54], LogPi=[55, 55], LogQiMul=[60] * 6),
    dict(LogN=15, T=65537, LogQi=[59, 59, 59] + [58] * 9, LogPi=[60, 60, 60], LogQiMul=[60] * 12),
]
GaloisGen = 5  # bfv/bfv.go


def GenModuli(params):
    """bfv/utils.go:26-85: primes per bit size, dealt to Q, then P, then QMul.  Host only."""
    from .ring import GenerateNTTPrimes

    need = {}
    for b in list(params["LogQi"]) + list(params["LogPi"]) + list(params["LogQiMul"]):
        if b > 60:
            raise ValueError("cannot GenModuli: the provided moduli sizes must be smaller than 61")
        need[b] = need.get(b, 0) + 1
    primes = {b: GenerateNTTPrimes(b, params["LogN"], n) for b, n in need.items()}
    out = []
    for key in ("LogQi", "LogPi", "LogQiMul"):
        out.append([primes[b].pop(0) for b in params[key]])
    return out  # Q, P, QMul


class Evaluator:
    """ring part of bfv.NewEvaluator (bfv/evaluator.go:62-104)"""

    def __init__(self, contextQ, contextQMul, contextP, t):
        self.contextQ, self.contextQMul, self.contextP, self.t = contextQ, contextQMul, contextP, t
        h = vp()
        check(lib().lg_bfv_eval_create(contextQ.h, contextQMul.h, contextP.h, t, C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            lib().lg_bfv_eval_destroy(self.h)
        except Exception:
            pass

    def Mul(self, ct0, ct1, ctOut, stream=None):
        """degree 1 x degree 1 -> degree 2 (ctOut = three Polys); ct0 is ct1 selects the squaring branch"""
        check(lib().lg_bfv_mul(self.h, ct0[0].h, ct0[1].h, ct1[0].h, ct1[1].h, ctOut[0].h, ctOut[1].h, ctOut[2].h, _s(stream)))

    def switchKeys(self, cx, evakey, p0, p1, stream=None):
        check(lib().lg_bfv_switch_keys_core(self.h, cx.h, evakey.h, p0.h, p1.h, _s(stream)))

    def Relinearize(self, ct0, evakey, ctOut, stream=None):
        check(lib().lg_bfv_relinearize(self.h, ct0[0].h, ct0[1].h, ct0[2].h, evakey.h, ctOut[0].h, ctOut[1].h, _s(stream)))

    def SwitchKeys(self, ct0, switchKey, ctOut, stream=None):
        check(lib().lg_bfv_switch_keys(self.h, ct0[0].h, ct0[1].h, switchKey.h, ctOut[0].h, ctOut[1].h, _s(stream)))

    def permute(self, ct0, generator, switchKey, ctOut, stream=None):
        check(lib().lg_bfv_permute(self.h, ct0[0].h, ct0[1].h, generator, switchKey.h, ctOut[0].h, ctOut[1].h, _s(stream)))


def NewEvaluator(contextQ, contextQMul, contextP, t):
    return Evaluator(contextQ, contextQMul, contextP, t)
