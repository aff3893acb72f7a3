"""Host-side mirror of the dbfv protocols on the ring hot path: CKG (dbfv/publickey_gen.go:44-67),
PCKS (dbfv/public_keyswitching.go:98-165), CKS (dbfv/keyswitching.go:66-122), RTG (dbfv/rotkey_gen.go:126-215)
and RKG (dbfv/relinkey_gen.go:190-355).  Same conventions as lattigpu.dckks; BFV ciphertexts are in the
coefficient domain, so PCKS / CKS add the noise after InvNTT and use the coefficient-domain ModDowns."""
from . import dckks, ring

# The ring sequences of the BFV rotation-key and relinearisation-key protocols are the CKKS ones line for line
# (dbfv/rotkey_gen.go:150-196 vs dckks/rotkey_gen.go:95-141; dbfv/relinkey_gen.go:214-355 vs
# dckks/relinkey_gen.go:65-223): both run over contextQP with the same digit structure.
RTGProtocol = dckks.RTGProtocol
RKGProtocol = dckks.RKGProtocol


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
