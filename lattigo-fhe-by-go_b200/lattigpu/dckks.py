"""Host-side mirror of the dckks protocols that sit on the ring hot path: collective public-key
generation (CKG, dckks/publickey_gen.go:18-52) and public collective key switching (PCKS,
dckks/public_keyswitching.go:9-113).  As in the reference the protocol logic is host code that
calls ring ops; here every ring op runs on the GPU through the C ABI.

Sampling stays with the caller (the reference draws from crypto/rand on the host): GenShare takes
the sampled polynomials in the coefficient domain and only the trailing NTT / arithmetic is done
here (what `SampleNTT` does after sampling, ring/gaussianSampler.go:289-292).
Shares of different parties are combined either pairwise (AggregateShares, as the reference) or,
with one party per GPU, by lattigpu.dist.Comm.AggregateShares (all-reduce + Reduce).
"""
from . import ring


class CKGProtocol:
    """dckks/publickey_gen.go:9-52 over contextQP"""

    def __init__(self, contextQP):
        self.contextQP = contextQP

    def AllocateShares(self, batch=1):
        return self.contextQP.NewPoly(batch)

    def GenShare(self, sk, crs, shareOut, e, stream=None):
        """shareOut = NTT(e) - sk*crs   (:39-42; e = gaussian sample, coefficient domain)"""
        self.contextQP.NTT(e, shareOut, stream=stream)
        self.contextQP.MulCoeffsMontgomeryAndSub(sk, crs, shareOut, stream=stream)

    def AggregateShares(self, share1, share2, shareOut, stream=None):
        self.contextQP.Add(share1, share2, shareOut, stream=stream)  # :45-47


class PCKSProtocol:
    """dckks/public_keyswitching.go:9-113"""

    def __init__(self, contextQ, contextP, contextQP):
        self.contextQ, self.contextP, self.contextQP = contextQ, contextP, contextQP
        self.baseconverter = ring.NewFastBasisExtender(contextQ, contextP)
        self._pool = {}  # tmp / share0tmp / share1tmp of the reference (public_keyswitching.go), per batch size
        self.nQP = contextQP.nl

    def AllocateShares(self, level, batch=1):
        return (self.contextQ.NewPolyLvl(level, batch), self.contextQ.NewPolyLvl(level, batch))

    def GenShare(self, level, sk, pk, ct1, shareOut, u, e0, e1, stream=None):
        """:63-96.  u = ternary sample already in Montgomery form (coefficient domain, over QP),
        e0 / e1 = smudging / gaussian samples (coefficient domain, over QP); pk = (pk0, pk1) over QP;
        ct1 = ct.Value()[1]; sk over Q (NTT + Montgomery)."""
        K = self.contextQP
        batch = u.batch
        if batch not in self._pool:
            self._pool[batch] = (K.NewPoly(batch), K.NewPoly(batch), K.NewPoly(batch), self.contextQ.NewPoly(batch))
        tmp, s0, s1, tq = self._pool[batch]
        K.NTT(u, tmp, stream=stream)  # SampleTernaryMontgomeryNTT :68
        K.MulCoeffsMontgomery(tmp, pk[0], s0, stream=stream)  # :71-72
        K.MulCoeffsMontgomery(tmp, pk[1], s1, stream=stream)
        K.NTT(e0, tmp, stream=stream)  # :75-76
        K.Add(s0, tmp, s0, stream=stream)
        K.NTT(e1, tmp, stream=stream)  # :77-78
        K.Add(s1, tmp, s1, stream=stream)
        self.baseconverter.ModDownNTTPQ(level, s0, shareOut[0], stream=stream)  # :81
        self.baseconverter.ModDownNTTPQ(level, s1, shareOut[1], stream=stream)  # :84
        self.contextQ.MulCoeffsMontgomeryAndAddLvl(level, ct1, sk, shareOut[0], stream=stream)  # :87

    def AggregateShares(self, share1, share2, shareOut, level, stream=None):
        self.contextQ.AddLvl(level, share1[0], share2[0], shareOut[0], stream=stream)  # :99-103
        self.contextQ.AddLvl(level, share1[1], share2[1], shareOut[1], stream=stream)

    def KeySwitch(self, combined, ct, ctOut, level, stream=None):
        self.contextQ.AddLvl(level, ct[0], combined[0], ctOut[0], stream=stream)  # :107-111
        self.contextQ.CopyLvl(level, combined[1], ctOut[1], stream=stream)
