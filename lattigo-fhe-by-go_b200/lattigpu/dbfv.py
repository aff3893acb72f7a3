"""Host-side mirror of the dbfv protocols on the ring hot path: CKG (dbfv/publickey_gen.go:44-67) and
PCKS (dbfv/public_keyswitching.go:98-165).  Same conventions as lattigpu.dckks; BFV ciphertexts
are in the coefficient domain, so PCKS adds the noise after InvNTT and uses ModDownPQ."""
from . import ring


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
