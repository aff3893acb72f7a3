"""Host-side mirror of the dbfv protocols on the ring hot path: CKG (dbfv/publickey_gen.go:44-67),
PCKS (dbfv/public_keyswitching.go:98-165), CKS (dbfv/keyswitching.go:66-122), RTG (dbfv/rotkey_gen.go:126-215)
RKG (dbfv/relinkey_gen.go:190-355) and Refresh (dbfv/public_refresh.go:77-214).  Same conventions as lattigpu.dckks; BFV ciphertexts are in the
coefficient domain, so PCKS / CKS add the noise after InvNTT and use the coefficient-domain ModDowns."""
import ctypes as C

import numpy as np

from . import dckks, ring
from ._lib import check, lib, vp
from .ring import _s

# The ring sequences of the BFV rotation-key and relinearisation-key protocols are the CKKS ones line for line
# (dbfv/rotkey_gen.go:150-196 vs dckks/rotkey_gen.go:95-141; dbfv/relinkey_gen.go:214-355 vs
# dckks/relinkey_gen.go:65-223): both run over contextQP with the same digit structure.
RTGProtocol = dckks.RTGProtocol
RKGProtocol = dckks.RKGProtocol


def RKGProtocolNaive(contextQ, contextP, contextQP):
    """dbfv/relinkey_gen_naive.go: the dckks sequence, with the second round-one sample going to shareOut[i][1] (:76)"""
    return dckks.RKGProtocolNaive(contextQ, contextP, contextQP, second_error_into=1)


class CKGProtocol:
    def __init__(self, contextQP):
        self.context = contextQP

    def AllocateShares(self, batch=1):
        return self.context.NewPoly(batch)

    def GenShare(self, sk, crs, shareOut, e, stream=None):
        self.context.NTT(e, shareOut, stream=stream)  # :54-57
        self.context.MulCoeffsMontgomeryAndSub(sk, crs, shareOut, stream=stream)

    def AggregateShares(self, share1, share2, shareOut, stream=None):
        self.context.Add(share1, share2, shareOut, stream=stream)  # :60-62


class PCKSProtocol:
    def __init__(self, contextQ, contextP, contextQP):
        self.contextQ, self.contextP, self.contextQP = contextQ, contextP, contextQP
        self.baseconverter = ring.NewFastBasisExtender(contextQ, contextP)
        self._pool = {}  # tmp / share0tmp / share1tmp of the reference (public_keyswitching.go), per batch size

    def AllocateShares(self, batch=1):
        return (self.contextQ.NewPoly(batch), self.contextQ.NewPoly(batch))

    def GenShare(self, sk, pk, ct1, shareOut, u, e0, e1, stream=None):
        """:111-146; u ternary (Montgomery, coefficient domain, QP); e0/e1 noise over QP (coefficient domain)"""
        Q, K = self.contextQ, self.contextQP
        batch = u.batch
        level = Q.nl - 1
        if batch not in self._pool:
            self._pool[batch] = (K.NewPoly(batch), K.NewPoly(batch), K.NewPoly(batch), self.contextQ.NewPoly(batch))
        tmp, s0, s1, tq = self._pool[batch]
        K.NTT(u, tmp, stream=stream)  # :118
        K.MulCoeffsMontgomery(tmp, pk[0], s0, stream=stream)  # :121-122
        K.MulCoeffsMontgomery(tmp, pk[1], s1, stream=stream)
        K.InvNTT(s0, s0, stream=stream)  # :124-125
        K.InvNTT(s1, s1, stream=stream)
        K.Add(s0, e0, s0, stream=stream)  # SampleAndAdd :128-129
        K.Add(s1, e1, s1, stream=stream)
        self.baseconverter.ModDownPQ(level, s0, shareOut[0], stream=stream)  # :132-135
        self.baseconverter.ModDownPQ(level, s1, shareOut[1], stream=stream)
        t = tq
        Q.NTT(ct1, t, stream=stream)  # :138-140
        Q.MulCoeffsMontgomery(t, sk, t, stream=stream)
        Q.InvNTT(t, t, stream=stream)
        Q.Add(shareOut[0], t, shareOut[0], stream=stream)  # :143

    def AggregateShares(self, share1, share2, shareOut, stream=None):
        self.contextQ.Add(share1[0], share2[0], shareOut[0], stream=stream)  # :149-153
        self.contextQ.Add(share1[1], share2[1], shareOut[1], stream=stream)

    def KeySwitch(self, combined, ct, ctOut, stream=None):
        self.contextQ.Add(ct[0], combined[0], ctOut[0], stream=stream)  # :156-160
        self.contextQ.Copy(combined[1], ctOut[1], stream=stream)


class CKSProtocol:
    """dbfv/keyswitching.go:9-122"""

    def __init__(self, contextQ, contextP, contextQP):
        self.contextQ, self.contextP, self.contextQP = contextQ, contextP, contextQP
        self.baseconverter = ring.NewFastBasisExtender(contextQ, contextP)
        self.Pbig = 1
        for p in contextP.Modulus:
            self.Pbig *= int(p)

    def AllocateShare(self, batch=1):
        return self.contextQ.NewPoly(batch)

    def GenShare(self, skInput, skOutput, ct1, shareOut, e, stream=None):
        """:73-106.  ct1 = ct.Value()[1] (coefficient domain); e = smudging sample over QP (coefficient domain).
        shareOut = (InvNTT((skIn - skOut) * NTT(ct1) * P) + e) / P"""
        Q = self.contextQ
        nQ, level = Q.nl, Q.nl - 1
        delta, t = Q.NewPoly(skInput.batch), Q.NewPoly(ct1.batch)
        Q.Sub(skInput, skOutput, delta, stream=stream)  # :75
        Q.NTT(ct1, t, stream=stream)  # :87
        Q.MulCoeffsMontgomery(t, delta, shareOut, stream=stream)  # :88
        Q.MulScalarBigint(shareOut, self.Pbig, shareOut, stream=stream)  # :89
        Q.InvNTT(shareOut, shareOut, stream=stream)  # :91
        Q.Add(shareOut, e.view(0, nQ), shareOut, stream=stream)  # :93-94 (the sample is not transformed)
        hP = e.view(nQ, self.contextP.nl).CopyNew(stream=stream)  # :96-102 (hP starts at zero)
        self.baseconverter.ModDownSplitedPQ(level, shareOut, hP, shareOut, stream=stream)  # :104

    def AggregateShares(self, share1, share2, shareOut, stream=None):
        self.contextQ.Add(share1, share2, shareOut, stream=stream)  # :112-114

    def KeySwitch(self, combined, ct, ctOut, stream=None):
        self.contextQ.Add(ct[0], combined, ctOut[0], stream=stream)  # :117-121
        self.contextQ.Copy(ct[1], ctOut[1], stream=stream)


class RefreshProtocol:
    """dbfv/public_refresh.go:77-214.  hP is protocol state as in the reference: GenShares adds the P limbs of the
    error sample into it with plain uint64 additions and never clears it (:128-134)."""

    def __init__(self, contextQ, contextP, contextQP, t):
        self.contextQ, self.contextP, self.contextQP, self.t = contextQ, contextP, contextQP, int(t)
        self.tmp1, self.tmp2 = contextQP.NewPoly(), contextQP.NewPoly()
        self.hP = contextP.NewPoly()
        self.hP.Zero()
        self.baseconverter = ring.NewFastBasisExtender(contextQ, contextP)
        self.scaler = ring.NewSimpleScaler(self.t, contextQ)  # Recode builds one per call (:184)
        self.Pbig = 1
        for p in contextP.Modulus:
            self.Pbig *= int(p)
        h = vp()
        check(lib().lg_bfv_lift_create(contextQ.h, self.t, C.byref(h)))  # deltaMont, dbfv.go / bfv/utils.go:9-23
        self._lift = h

    def __del__(self):
        try:
            lib().lg_bfv_lift_destroy(self._lift)
        except Exception:
            pass

    def AllocateShares(self):
        return (self.contextQ.NewPoly(), self.contextQ.NewPoly())  # :99-102

    def _lift_into(self, m, out, stream):
        """lift (:207-214): out.Coeffs[i] = MRed(m.Coeffs[0], deltaMont[i]) for every limb of Q"""
        check(lib().lg_bfv_lift_apply(self._lift, m.h, out.h, _s(stream)))

    def GenShares(self, sk, ct1, crs, share, e, ePrime, mask, stream=None):
        """:105-169.  sk over QP (NTT + Montgomery); ct1 = ciphertext.Value()[1] (coefficient domain, over Q);
        crs over QP (coefficient domain); e / ePrime = the two gaussian samples over QP (coefficient domain);
        mask = the uniform plaintext coefficients in [0, t) (host array [N] or device poly [1][1][N])."""
        Q, K, P = self.contextQ, self.contextQP, self.contextP
        nQ, level = Q.nl, Q.nl - 1
        h0, h1 = share
        tq = self.tmp1.view(0, nQ)
        Q.NTT(ct1, tq, stream=stream)  # :116
        Q.MulCoeffsMontgomery(sk, tq, h0, stream=stream)  # :117
        Q.InvNTT(h0, h0, stream=stream)  # :119
        Q.MulScalarBigint(h0, self.Pbig, h0, stream=stream)  # :122
        Q.Add(h0, e.view(0, nQ), h0, stream=stream)  # :125-126
        P.AddNoMod(self.hP, e.view(nQ, P.nl), self.hP, stream=stream)  # :128-134
        self.baseconverter.ModDownSplitedPQ(level, h0, self.hP, h0, stream=stream)  # :137
        K.NTT(crs, self.tmp1, stream=stream)  # :140
        K.MulCoeffsMontgomery(sk, self.tmp1, self.tmp2, stream=stream)  # :141
        K.Neg(self.tmp2, self.tmp2, stream=stream)  # :142
        K.InvNTT(self.tmp2, self.tmp2, stream=stream)  # :143
        K.Add(self.tmp2, ePrime, self.tmp2, stream=stream)  # SampleAndAdd :146
        self.baseconverter.ModDownPQ(level, self.tmp2, h1, stream=stream)  # :149
        if not isinstance(mask, ring.Poly):
            mask = ring.Poly.from_numpy(np.asarray(mask, dtype=np.uint64).reshape(1, 1, -1))
        self._lift_into(mask, tq, stream)  # :152-153
        Q.Add(h0, tq, h0, stream=stream)  # :156
        Q.Sub(h1, tq, h1, stream=stream)  # :159

    def Aggregate(self, share1, share2, shareOut, stream=None):
        self.contextQ.Add(share1[0], share2[0], shareOut[0], stream=stream)  # :172-175
        self.contextQ.Add(share1[1], share2[1], shareOut[1], stream=stream)

    def Decrypt(self, ct0, shareDecrypt, sharePlaintext, stream=None):
        self.contextQ.Add(ct0, shareDecrypt, sharePlaintext, stream=stream)  # :178-180

    def Recode(self, sharePlaintext, sharePlaintextOut, stream=None):
        """:183-188: t/Q scaling, then the lift back to Q"""
        m = ring.Poly(self.contextQ.N, 1, sharePlaintext.batch)
        self.scaler.Scale(sharePlaintext, m, stream=stream)
        self._lift_into(m, sharePlaintextOut, stream)

    def Recrypt(self, sharePlaintext, crs, shareRecrypt, ctOut, stream=None):
        self.contextQ.Add(sharePlaintext, shareRecrypt, ctOut[0], stream=stream)  # :194
        self.baseconverter.ModDownPQ(self.contextQ.nl - 1, crs, ctOut[1], stream=stream)  # :197

    def Finalize(self, ct, crs, share, ctOut, stream=None):
        """:202-206"""
        tq = self.tmp1.view(0, self.contextQ.nl)
        self.Decrypt(ct[0], share[0], tq, stream=stream)
        self.Recode(tq, tq, stream=stream)
        self.Recrypt(tq, crs, share[1], ctOut, stream=stream)
