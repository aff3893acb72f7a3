"""Host-side mirror of the dckks protocols that sit on the ring hot path: collective public-key
generation (CKG, dckks/publickey_gen.go:18-52), public collective key switching (PCKS,
dckks/public_keyswitching.go:9-113), collective key switching (CKS, dckks/keyswitching.go:55-108),
rotation-key generation (RTG, dckks/rotkey_gen.go:75-174) and the three-round relinearisation-key
generation (RKG, dckks/relinkey_gen.go:60-223), its two-round "naive" variant (relinkey_gen_naive.go) and the collective refresh (dckks/public_refresh.go:9-147).  As in the reference the protocol logic is host code that
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


def _add_digit_limbs(K, src, dst, i, alpha, levels, tmp, stream):
    """dst[index] = CRed(dst[index] + src[index]) for the limbs index = i*alpha + j of digit i that lie
    below `levels` (the j-loops of relinkey_gen.go:88-105 / rotkey_gen.go:116-132); dst is canonical, so adding a
    poly that is zero elsewhere leaves the other limbs unchanged."""
    lo = i * alpha
    hi = min(lo + alpha, levels)
    tmp.Zero(stream=stream)
    K.CopyLvl(hi - lo - 1, src.view(lo, hi - lo), tmp.view(lo, hi - lo), stream=stream)
    K.Add(dst, tmp, dst, stream=stream)


class CKSProtocol:
    """dckks/keyswitching.go:9-108"""

    def __init__(self, contextQ, contextP, contextQP):
        self.contextQ, self.contextP, self.contextQP = contextQ, contextP, contextQP
        self.baseconverter = ring.NewFastBasisExtender(contextQ, contextP)
        self.Pbig = 1
        for p in contextP.Modulus:
            self.Pbig *= int(p)

    def AllocateShare(self, level, batch=1):
        return self.contextQ.NewPolyLvl(level, batch)

    def GenShare(self, level, skInput, skOutput, ct1, shareOut, e, stream=None):
        """:62-96.  skInput / skOutput over Q (NTT + Montgomery), ct1 = ct.Value()[1], e = smudging sample over
        QP in the coefficient domain.  shareOut = ((skIn - skOut) * ct1 * P + NTT(e)) / P"""
        Q, K = self.contextQ, self.contextQP
        nQ = Q.nl
        delta = Q.NewPoly(skInput.batch)
        Q.Sub(skInput, skOutput, delta, stream=stream)  # :64
        Q.MulCoeffsMontgomeryLvl(level, ct1, delta, shareOut, stream=stream)  # :74
        Q.MulScalarBigintLvl(level, shareOut, self.Pbig, shareOut, stream=stream)  # :76
        tmp = K.NewPoly(e.batch)
        K.NTT(e, tmp, stream=stream)  # SampleNTT :79
        Q.AddLvl(level, shareOut, tmp.view(0, nQ), shareOut, stream=stream)  # :80
        hP = tmp.view(nQ, self.contextP.nl)  # :82-88 (hP starts at zero)
        self.baseconverter.ModDownSplitedNTTPQ(level, shareOut, hP, shareOut, stream=stream)  # :90

    def AggregateShares(self, level, share1, share2, shareOut, stream=None):
        self.contextQ.AddLvl(level, share1, share2, shareOut, stream=stream)  # :100-102

    def KeySwitch(self, level, combined, ct, ctOut, stream=None):
        self.contextQ.AddLvl(level, ct[0], combined, ctOut[0], stream=stream)  # :105-108
        self.contextQ.CopyLvl(level, ct[1], ctOut[1], stream=stream)


class RTGProtocol:
    """dckks/rotkey_gen.go:9-174 over contextQP"""

    def __init__(self, contextQ, contextP, contextQP):
        self.contextQP = contextQP
        self.levels, self.alpha = contextQ.nl, contextP.nl
        self.beta = -(-self.levels // self.alpha)
        self.Pbig = 1
        for p in contextP.Modulus:
            self.Pbig *= int(p)

    def genShare(self, sk, galEl, crp, errors, stream=None):
        """:95-141.  sk over QP (NTT + Montgomery); crp[i] uniform over QP; errors[i] gaussian coefficients
        (device polys, coefficient domain).  Returns the beta share polys."""
        K = self.contextQP
        tmpPoly, tmp = K.NewPoly(), K.NewPoly()
        ring.PermuteNTT(sk, galEl, tmpPoly, stream=stream)  # :99
        K.MulScalarBigint(tmpPoly, self.Pbig, tmpPoly, stream=stream)  # :101
        K.InvMForm(tmpPoly, tmpPoly, stream=stream)  # :103
        out = []
        for i in range(self.beta):
            ek = K.NewPoly()
            K.NTT(errors[i], ek, stream=stream)  # SampleNTTNew :110
            _add_digit_limbs(K, tmpPoly, ek, i, self.alpha, self.levels, tmp, stream)  # :116-132
            K.MulCoeffsMontgomeryAndSub(crp[i], sk, ek, stream=stream)  # :135
            K.MForm(ek, ek, stream=stream)  # :136
            out.append(ek)
        return out

    def Aggregate(self, share1, share2, shareOut, stream=None):
        for a, b, c in zip(share1, share2, shareOut):
            self.contextQP.Add(a, b, c, stream=stream)  # :158-160

    def Finalize(self, share, crp, stream=None):
        """:164-174: evakey[i] = (share[i], MForm(crp[i])) -> [beta][2] polys over QP"""
        K = self.contextQP
        key = []
        for i in range(self.beta):
            k1 = K.NewPoly()
            K.MForm(crp[i], k1, stream=stream)
            key.append((share[i].CopyNew(stream=stream), k1))
        return key


class RKGProtocol:
    """dckks/relinkey_gen.go:9-223 over contextQP; u = ephemeral key, sk = secret share (NTT + Montgomery)"""

    def __init__(self, contextQ, contextP, contextQP):
        self.contextQP = contextQP
        self.levels, self.alpha = contextQ.nl, contextP.nl
        self.beta = -(-self.levels // self.alpha)
        self.Pbig = 1
        for p in contextP.Modulus:
            self.Pbig *= int(p)

    def GenShareRoundOne(self, u, sk, crp, errors, stream=None):
        """:65-112: share[i] = -u*crp[i] + P*s*w_i + NTT(e_i)"""
        K = self.contextQP
        pool, tmp = sk.CopyNew(stream=stream), K.NewPoly()
        K.MulScalarBigint(pool, self.Pbig, pool, stream=stream)  # :77
        K.InvMForm(pool, pool, stream=stream)  # :79
        out = []
        for i in range(self.beta):
            h = K.NewPoly()
            K.NTT(errors[i], h, stream=stream)  # :84
            _add_digit_limbs(K, pool, h, i, self.alpha, self.levels, tmp, stream)  # :87-105
            K.MulCoeffsMontgomeryAndSub(u, crp[i], h, stream=stream)  # :108
            out.append(h)
        return out

    def AggregateShareRoundOne(self, share1, share2, shareOut, stream=None):
        for a, b, c in zip(share1, share2, shareOut):
            self.contextQP.Add(a, b, c, stream=stream)

    def GenShareRoundTwo(self, round1, sk, crp, errors1, errors2, stream=None):
        """:135-163: (round1[i]*sk + NTT(e1_i), sk*crp[i] + NTT(e2_i))"""
        K = self.contextQP
        out = []
        for i in range(self.beta):
            s0, s1, pool = K.NewPoly(), K.NewPoly(), K.NewPoly()
            K.MulCoeffsMontgomery(round1[i], sk, s0, stream=stream)  # :146
            K.NTT(errors1[i], pool, stream=stream)  # :149
            K.Add(s0, pool, s0, stream=stream)
            K.NTT(errors2[i], s1, stream=stream)  # :154
            K.MulCoeffsMontgomeryAndAdd(sk, crp[i], s1, stream=stream)  # :156
            out.append((s0, s1))
        return out

    def AggregateShareRoundTwo(self, share1, share2, shareOut, stream=None):
        for a, b, c in zip(share1, share2, shareOut):
            self.contextQP.Add(a[0], b[0], c[0], stream=stream)
            self.contextQP.Add(a[1], b[1], c[1], stream=stream)

    def GenShareRoundThree(self, round2, u, sk, errors, stream=None):
        """:186-199: (u - sk) * round2[i][1] + NTT(e3_i)"""
        K = self.contextQP
        pool = K.NewPoly()
        K.Sub(u, sk, pool, stream=stream)  # :191
        out = []
        for i in range(self.beta):
            h = K.NewPoly()
            K.NTT(errors[i], h, stream=stream)  # :196
            K.MulCoeffsMontgomeryAndAdd(pool, round2[i][1], h, stream=stream)  # :197
            out.append(h)
        return out

    def AggregateShareRoundThree(self, share1, share2, shareOut, stream=None):
        for a, b, c in zip(share1, share2, shareOut):
            self.contextQP.Add(a, b, c, stream=stream)

    def GenRelinearizationKey(self, round2, round3, stream=None):
        """:210-223: key[i] = (MForm(round2[i][0] + round3[i]), MForm(round2[i][1]))"""
        K = self.contextQP
        key = []
        for i in range(self.beta):
            k0, k1 = K.NewPoly(), K.NewPoly()
            K.Add(round2[i][0], round3[i], k0, stream=stream)
            K.MForm(k0, k0, stream=stream)
            K.MForm(round2[i][1], k1, stream=stream)
            key.append((k0, k1))
        return key


class RKGProtocolNaive:
    """dckks/relinkey_gen_naive.go:11-200 (and dbfv/relinkey_gen_naive.go with `second_error_into=1`) over contextQP.
    pk = (pk0, pk1) collective public key; sk = secret share (NTT + Montgomery).  Round one draws two gaussian samples
    per digit: the dbfv file transforms them into shareOut[i][0] and shareOut[i][1] (:74-76); the dckks file writes
    BOTH into shareOut[i][0] (:73,:75), so the first is overwritten and shareOut[i][1] keeps its content (zero for
    freshly allocated shares).  `second_error_into` selects which (0 = dckks, 1 = dbfv)."""

    def __init__(self, contextQ, contextP, contextQP, second_error_into=0):
        self.contextQP = contextQP
        self.levels, self.alpha = contextQ.nl, contextP.nl
        self.beta = -(-self.levels // self.alpha)
        self.second = int(second_error_into)
        self.Pbig = 1
        for p in contextP.Modulus:
            self.Pbig *= int(p)

    def AllocateShares(self):
        K = self.contextQP
        mk = lambda: [(self._zero(K), self._zero(K)) for _ in range(self.beta)]
        return mk(), mk()

    @staticmethod
    def _zero(K):
        p = K.NewPoly()
        p.Zero()
        return p

    def GenShareRoundOne(self, sk, pk, shareOut, errors, us, stream=None):
        """:53-108.  errors[i] = (first, second) gaussian samples of digit i (coefficient domain, over QP);
        us[i] = ternary sample of digit i (Montgomery form, coefficient domain)."""
        K = self.contextQP
        pool, tmp = sk.CopyNew(stream=stream), K.NewPoly()
        K.MulScalarBigint(pool, self.Pbig, pool, stream=stream)  # :59
        K.InvMForm(pool, pool, stream=stream)  # :61
        for i in range(self.beta):
            K.NTT(errors[i][0], shareOut[i][0], stream=stream)  # :73
            K.NTT(errors[i][1], shareOut[i][self.second], stream=stream)  # :75 (dbfv :76)
            _add_digit_limbs(K, pool, shareOut[i][0], i, self.alpha, self.levels, tmp, stream)  # :78-95
        for i in range(self.beta):
            K.NTT(us[i], pool, stream=stream)  # SampleTernaryMontgomeryNTT :100
            K.MulCoeffsMontgomeryAndAdd(pk[0], pool, shareOut[i][0], stream=stream)  # :101
            K.MulCoeffsMontgomeryAndAdd(pk[1], pool, shareOut[i][1], stream=stream)  # :102

    def AggregateShareRoundOne(self, share1, share2, shareOut, stream=None):
        for a, b, c in zip(share1, share2, shareOut):  # :111-120
            self.contextQP.Add(a[0], b[0], c[0], stream=stream)
            self.contextQP.Add(a[1], b[1], c[1], stream=stream)

    def GenShareRoundTwo(self, round1, sk, pk, shareOut, us, errors, stream=None):
        """:129-166.  us[i] = ternary sample; errors[i] = (e0, e1) gaussian samples of digit i"""
        K = self.contextQP
        pool = K.NewPoly()
        for i in range(self.beta):
            K.MulCoeffsMontgomery(round1[i][0], sk, shareOut[i][0], stream=stream)  # :141
            K.MulCoeffsMontgomery(round1[i][1], sk, shareOut[i][1], stream=stream)  # :142
            K.NTT(us[i], pool, stream=stream)  # :145
            K.MulCoeffsMontgomeryAndAdd(pk[0], pool, shareOut[i][0], stream=stream)  # :148
            K.MulCoeffsMontgomeryAndAdd(pk[1], pool, shareOut[i][1], stream=stream)  # :151
            K.NTT(errors[i][0], pool, stream=stream)  # :154
            K.Add(shareOut[i][0], pool, shareOut[i][0], stream=stream)
            K.NTT(errors[i][1], pool, stream=stream)  # :158
            K.Add(shareOut[i][1], pool, shareOut[i][1], stream=stream)

    AggregateShareRoundTwo = AggregateShareRoundOne  # :169-177

    def GenRelinearizationKey(self, round2, stream=None):
        """:180-200: key[i] = (MForm(round2[i][0]), MForm(round2[i][1]))"""
        K = self.contextQP
        key = []
        for i in range(self.beta):
            k0, k1 = K.NewPoly(), K.NewPoly()
            K.MForm(round2[i][0], k0, stream=stream)
            K.MForm(round2[i][1], k1, stream=stream)
            key.append((k0, k1))
        return key


class RefreshProtocol:
    """dckks/public_refresh.go:9-147 over contextQ.  The ring ops run on the device; the big-integer steps of the
    reference (ring.RandInt mask, SetCoefficientsBigint, PolyToBigint: math/big host code, ring_context.go:343-421)
    stay host code here as well (Python integers).  One ciphertext per call (batch 1), like the reference."""

    def __init__(self, contextQ):
        self.contextQ = contextQ
        self.tmp = contextQ.NewPoly()

    def AllocateShares(self, levelStart):
        return self.contextQ.NewPolyLvl(levelStart), self.contextQ.NewPoly()  # :38-40

    def _modulus(self, nl):
        q = 1
        for m in self.contextQ.Modulus[:nl]:
            q *= int(m)
        return q

    def _set_bigint(self, nl, coeffs, p):
        """SetCoefficientsBigintLvl (ring_context.go:356-367): Euclidean residues, like big.Int.Mod"""
        import numpy as np
        a = np.array([[int(c) % int(q) for c in coeffs] for q in self.contextQ.Modulus[:nl]], dtype=np.uint64)
        p.set(a)

    def GenShares(self, sk, levelStart, nParties, ct1, crs, shareDecrypt, shareRecrypt, mask, e0, e1, stream=None):
        """:43-98.  mask = the N values ring.RandInt(bound) returned (bound = Q_levelStart / (2 nParties), :48-54),
        centred here as :57-63 do; e0 / e1 = the two gaussian samples (coefficient domain, over contextQ);
        sk over Q (NTT + Montgomery), ct1 = ciphertext.Value()[1], crs in the NTT domain."""
        K = self.contextQ
        bound = self._modulus(levelStart + 1) // (2 * nParties)
        half = bound >> 1
        mask = [int(m) - bound if int(m) >= half else int(m) for m in mask]
        self._set_bigint(levelStart + 1, mask, shareDecrypt)  # :66
        self._set_bigint(K.nl, mask, shareRecrypt)  # :68
        K.NTTLvl(levelStart, shareDecrypt, shareDecrypt, stream=stream)  # :75
        K.NTT(shareRecrypt, shareRecrypt, stream=stream)  # :76
        K.MulCoeffsMontgomeryAndAddLvl(levelStart, sk, ct1, shareDecrypt, stream=stream)  # :79
        K.MulCoeffsMontgomeryAndAdd(sk, crs, shareRecrypt, stream=stream)  # :82
        K.NTT(e0, self.tmp, stream=stream)  # SampleNTT :85
        K.AddLvl(levelStart, shareDecrypt, self.tmp, shareDecrypt, stream=stream)
        K.NTT(e1, self.tmp, stream=stream)  # :89
        K.Add(shareRecrypt, self.tmp, shareRecrypt, stream=stream)
        K.Neg(shareRecrypt, shareRecrypt, stream=stream)  # :93
        self.tmp.Zero(stream=stream)  # :95

    def Aggregate(self, share1, share2, shareOut, stream=None):
        self.contextQ.AddLvl(share1.nlimbs - 1, share1, share2, shareOut, stream=stream)  # :101-103

    def Decrypt(self, level, ct0, shareDecrypt, stream=None):
        """:106-108 on ciphertext.Value()[0] at `level`"""
        self.contextQ.AddLvl(level, ct0, shareDecrypt, ct0, stream=stream)

    def Recode(self, level, ct0, stream=None):
        """:111-139.  Returns value[0] grown to every limb of Q (the reference appends the missing limbs to Coeffs)."""
        K = self.contextQ
        nl = level + 1
        K.InvNTTLvl(level, ct0, ct0, stream=stream)
        a = ct0.numpy(nl=nl, stream=stream, squeeze=False)[0]
        mods = [int(q) for q in K.Modulus[:nl]]
        big = self._modulus(nl)
        rec = [(big // q) * pow(big // q, -1, q) for q in mods]  # PolyToBigint, ring_context.go:384-421
        cols = [a[i].tolist() for i in range(nl)]
        half = big >> 1
        vals = []
        for x in range(K.N):
            v = sum(cols[i][x] * rec[i] for i in range(nl)) % big
            vals.append(v - big if v >= half else v)  # :131-136
        out = K.NewPoly()
        self._set_bigint(K.nl, vals, out)  # :138
        K.NTT(out, out, stream=stream)  # :140
        return out

    def Recrypt(self, ct0, crs, shareRecrypt, stream=None):
        """:142-147: returns (value[0] + shareRecrypt, copy of crs)"""
        self.contextQ.Add(ct0, shareRecrypt, ct0, stream=stream)
        return ct0, crs.CopyNew(stream=stream)


def evakey_to_numpy(key):
    """[beta] x (poly, poly) -> [beta][2][nQP][N] uint64, the layout of ckks.SwitchingKey"""
    import numpy as np

    return np.ascontiguousarray(np.stack([np.stack([k0.numpy(), k1.numpy()]) for k0, k1 in key]))
