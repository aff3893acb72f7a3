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
    """ring part of bfv.NewEvaluator (bfv/evaluator.go:62-104).  Degree 1 x degree 1 products and every key switch are
    fused device paths (lg_bfv_*); the general-degree tensor (:278-464), relinearisation of higher degrees (:480-507)
    and the rotation drivers (:578-690) are the reference's ring-op sequences, each op on the device through the ABI."""

    def __init__(self, contextQ, contextQMul, contextP, t):
        self.contextQ, self.contextQMul, self.contextP, self.t = contextQ, contextQMul, contextP, t
        h = vp()
        check(lib().lg_bfv_eval_create(contextQ.h, contextQMul.h, contextP.h, t, C.byref(h)))
        self.h = h
        self._q1q2 = None
        self.pHalf = 1  # :98 pHalf = QMul.ModulusBigint >> 1
        for q in contextQMul.Modulus:
            self.pHalf *= int(q)
        self.pHalf >>= 1

    def __del__(self):
        try:
            lib().lg_bfv_eval_destroy(self.h)
        except Exception:
            pass

    def Mul(self, ct0, ct1, ctOut, stream=None):
        """Mul = tensorAndRescale (:278-464, :467-470).  An operand is its tuple of value polys in the coefficient domain
        (a plaintext is a 1-tuple); ctOut has len(ct0) + len(ct1) - 1 polys.  ct0 is ct1 selects the squaring branches."""
        if len(ct0) == 2 and len(ct1) == 2:
            check(lib().lg_bfv_mul(self.h, ct0[0].h, ct0[1].h, ct1[0].h, ct1[1].h, ctOut[0].h, ctOut[1].h, ctOut[2].h,
                                   _s(stream)))
            return
        self._tensorAndRescaleGeneral(ct0, ct1, ctOut, stream)

    def _tensorAndRescaleGeneral(self, ct0, ct1, ctOut, stream):
        from .ring import NewFastBasisExtender

        Q, M = self.contextQ, self.contextQMul
        if self._q1q2 is None:
            self._q1q2 = NewFastBasisExtender(Q, M)  # baseconverterQ1Q2 (:95)
        bc = self._q1q2
        levelQ, levelQMul = Q.nl - 1, M.nl - 1
        nout = len(ct0) + len(ct1) - 1
        if len(ctOut) != nout:
            raise ValueError("cannot Mul: receiver must be of degree %d" % (nout - 1))
        batch = ct0[0].batch
        sq = ct0 is ct1

        def extend(ct):  # :299-312
            q1, q2 = [], []
            for v in ct:
                a, b = Q.NewPoly(batch), M.NewPoly(batch)
                bc.ModUpSplitQP(levelQ, v, b, stream=stream)
                Q.NTT(v, a, stream=stream)
                M.NTT(b, b, stream=stream)
                q1.append(a)
                q2.append(b)
            return q1, q2

        c0Q1, c0Q2 = extend(ct0)
        c1Q1, c1Q2 = (c0Q1, c0Q2) if sq else extend(ct1)
        c2Q1 = [Q.NewPoly(batch) for _ in range(nout)]  # NewPoly zeroes: :376-379
        c2Q2 = [M.NewPoly(batch) for _ in range(nout)]
        if sq:  # :382-404
            c00Q1, c00Q2 = [Q.NewPoly(batch) for _ in ct0], [M.NewPoly(batch) for _ in ct0]
            for i in range(len(ct0)):
                Q.MForm(c0Q1[i], c00Q1[i], stream=stream)
                M.MForm(c0Q2[i], c00Q2[i], stream=stream)
            for i in range(len(ct0)):
                for j in range(i + 1, len(ct0)):
                    Q.MulCoeffsMontgomery(c00Q1[i], c0Q1[j], c2Q1[i + j], stream=stream)
                    M.MulCoeffsMontgomery(c00Q2[i], c0Q2[j], c2Q2[i + j], stream=stream)
                    Q.Add(c2Q1[i + j], c2Q1[i + j], c2Q1[i + j], stream=stream)
                    M.Add(c2Q2[i + j], c2Q2[i + j], c2Q2[i + j], stream=stream)
            for i in range(len(ct0)):
                Q.MulCoeffsMontgomeryAndAdd(c00Q1[i], c0Q1[i], c2Q1[i << 1], stream=stream)
                M.MulCoeffsMontgomeryAndAdd(c00Q2[i], c0Q2[i], c2Q2[i << 1], stream=stream)
        else:  # :407-416
            for i in range(len(ct0)):
                Q.MForm(c0Q1[i], c0Q1[i], stream=stream)
                M.MForm(c0Q2[i], c0Q2[i], stream=stream)
                for j in range(len(ct1)):
                    Q.MulCoeffsMontgomeryAndAdd(c0Q1[i], c1Q1[j], c2Q1[i + j], stream=stream)
                    M.MulCoeffsMontgomeryAndAdd(c0Q2[i], c1Q2[j], c2Q2[i + j], stream=stream)
        t = self.t
        for i in range(nout):  # :424-463
            Q.InvNTT(c2Q1[i], c2Q1[i], stream=stream)
            M.InvNTT(c2Q2[i], c2Q2[i], stream=stream)
            bc.ModDownSplitedQP(levelQ, levelQMul, c2Q1[i], c2Q2[i], c2Q2[i], stream=stream)
            M.AddScalarBigint(c2Q2[i], self.pHalf, c2Q2[i], stream=stream)
            bc.ModUpSplitPQ(levelQMul, c2Q2[i], ctOut[i], stream=stream)
            Q.SubScalarBigint(ctOut[i], self.pHalf, ctOut[i], stream=stream)
            Q.MulScalar(ctOut[i], t, ctOut[i], stream=stream)

    def switchKeys(self, cx, evakey, p0, p1, stream=None):
        check(lib().lg_bfv_switch_keys_core(self.h, cx.h, evakey.h, p0.h, p1.h, _s(stream)))

    def Relinearize(self, ct0, evakey, ctOut, stream=None):
        """Relinearize (:480-530).  evakey: the SwitchingKey of degree 2, or the list evakey[deg-2] of an EvaluationKey
        for ciphertexts of higher degree; ctOut = two polys."""
        keys = list(evakey) if isinstance(evakey, (list, tuple)) else [evakey]
        deg = len(ct0) - 1
        if deg - 1 > len(keys):
            raise ValueError("cannot Relinearize: input ciphertext degree too large to allow relinearization")  # :518
        Q = self.contextQ
        if deg < 2:  # :521-524
            for a, c in zip(ct0, ctOut):
                if a is not c:
                    Q.Copy(a, c, stream=stream)
            return
        if deg == 2:
            check(lib().lg_bfv_relinearize(self.h, ct0[0].h, ct0[1].h, ct0[2].h, keys[0].h, ctOut[0].h, ctOut[1].h, _s(stream)))
            return
        if ctOut[0] is not ct0[0]:  # :484-487
            Q.Copy(ct0[0], ctOut[0], stream=stream)
            Q.Copy(ct0[1], ctOut[1], stream=stream)
        batch = ct0[0].batch
        p0, p1 = Q.NewPoly(batch), Q.NewPoly(batch)
        for d in range(deg, 1, -1):  # :492-496
            self.switchKeys(ct0[d], keys[d - 2], p0, p1, stream=stream)
            Q.Add(ctOut[0], p0, ctOut[0], stream=stream)
            Q.Add(ctOut[1], p1, ctOut[1], stream=stream)

    def SwitchKeys(self, ct0, switchKey, ctOut, stream=None):
        check(lib().lg_bfv_switch_keys(self.h, ct0[0].h, ct0[1].h, switchKey.h, ctOut[0].h, ctOut[1].h, _s(stream)))

    def permute(self, ct0, generator, switchKey, ctOut, stream=None):
        check(lib().lg_bfv_permute(self.h, ct0[0].h, ct0[1].h, generator, switchKey.h, ctOut[0].h, ctOut[1].h, _s(stream)))

    # ---- rotation drivers (:578-690); `evakey` is a RotationKeys --------------------------------------------------
    def RotateColumns(self, ct0, k, evakey, ctOut, stream=None):
        """:578-625"""
        N = self.contextQ.N
        k &= (N >> 1) - 1
        if k == 0:
            for a, c in zip(ct0, ctOut):
                self.contextQ.Copy(a, c, stream=stream)
            return
        if evakey.evakeyRotColLeft.get(k) is not None:
            self.permute(ct0, pow(GaloisGen, k, 2 * N), evakey.evakeyRotColLeft[k], ctOut, stream=stream)  # galElRotColLeft[k]
            return
        i, has = 1, True
        while i < (N >> 1):
            if evakey.evakeyRotColLeft.get(i) is None or evakey.evakeyRotColRight.get(i) is None:
                has = False
                break
            i <<= 1
        if not has:
            raise ValueError("cannot RotateColumns: specific rotation and pow2 rotations have not been generated")
        if bin(k).count("1") <= bin((N >> 1) - k).count("1"):
            self.rotateColumnsPow2(ct0, GaloisGen, k, evakey.evakeyRotColLeft, ctOut, stream=stream)  # :628-630
        else:
            genInv = pow(GaloisGen, 2 * N - 1, 2 * N)  # :634
            self.rotateColumnsPow2(ct0, genInv, (N >> 1) - k, evakey.evakeyRotColRight, ctOut, stream=stream)

    def rotateColumnsPow2(self, ct0, generator, k, evakeyRotCol, ctOut, stream=None):
        """:637-666"""
        N = self.contextQ.N
        mask = (N << 1) - 1
        if ct0[0] is not ctOut[0]:
            self.contextQ.Copy(ct0[0], ctOut[0], stream=stream)
            self.contextQ.Copy(ct0[1], ctOut[1], stream=stream)
        evakeyIndex = 1
        while k > 0:
            if k & 1:
                self.permute(ctOut, generator, evakeyRotCol[evakeyIndex], ctOut, stream=stream)
            generator = (generator * generator) & mask
            evakeyIndex <<= 1
            k >>= 1

    def RotateRows(self, ct0, evakey, ctOut, stream=None):
        """:669-680: galElRotRow = 2N - 1"""
        if evakey.evakeyRotRow is None:
            raise ValueError("cannot RotateRows: rotation key not generated")
        self.permute(ct0, 2 * self.contextQ.N - 1, evakey.evakeyRotRow, ctOut, stream=stream)


class RotationKeys:
    """bfv.RotationKeys (bfv/keygen.go): switching keys of the column rotations (left / right) and of the row rotation"""

    def __init__(self):
        self.evakeyRotColLeft, self.evakeyRotColRight = {}, {}
        self.evakeyRotRow = None


def NewEvaluator(contextQ, contextQMul, contextP, t):
    return Evaluator(contextQ, contextQMul, contextP, t)
