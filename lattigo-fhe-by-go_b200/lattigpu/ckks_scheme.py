"""Host-side mirror of the ring-op sequences of the CKKS key generator, encryptor and decryptor
(ckks/keygen.go:138-494, ckks/encryptor.go:179-362, ckks/decryptor.go:53-78) over the C ABI, so that
"encrypt -> MulRelin -> Rescale -> decrypt" stays on the device end to end.

As in lattigpu.dckks, sampling stays with the caller: the reference draws from crypto/rand on the host
(ring/ternarySampler.go, ring/gaussianSampler.go, ring.NewUniformPoly), so every entry point takes the
sampled values -- ternary / gaussian coefficients as small signed integers, uniform polynomials as
residues -- and runs what follows the sampling (MForm, NTT, the multiply-adds, ModDown) on the GPU.
Secret and public keys live over QP in NTT + Montgomery form, exactly like `SecretKey.sk` /
`PublicKey.pk` of the reference.
"""
import numpy as np

from . import ring
from .ckks import SwitchingKey


def signed_to_poly(context, coeffs, batch=1):
    """Residues of small signed coefficients over every modulus of `context`: x >= 0 -> x, x < 0 -> q + x
    (what the samplers write, e.g. ring/gaussianSampler.go:271).  coeffs: int array [N] or [batch][N], or a
    device Poly that already holds the residues (returned as is)."""
    if isinstance(coeffs, ring.Poly):
        return coeffs
    c = np.asarray(coeffs, dtype=np.int64)
    if c.ndim == 1:
        c = c[None, :]
    out = np.empty((c.shape[0], context.nl, context.N), dtype=np.uint64)
    for i, q in enumerate(context.Modulus):
        out[:, i, :] = np.where(c < 0, np.int64(q) + c, c).astype(np.uint64)
    assert out.shape[0] == batch
    return ring.Poly.from_numpy(out)


class KeyGenerator:
    """ckks.keyGenerator over contextQP (ckks/keygen.go:60-112)"""

    def __init__(self, contextQ, contextP):
        self.contextQ, self.contextP = contextQ, contextP
        self.contextQP = ring.NewContextWithParams(contextQ.N, list(contextQ.Modulus) + list(contextP.Modulus))
        self.levels, self.alpha = contextQ.nl, contextP.nl
        self.beta = -(-self.levels // self.alpha)
        self.Pbig = 1
        for p in contextP.Modulus:
            self.Pbig *= int(p)

    def GenSecretKey(self, ternary, stream=None):
        """:96-112: SampleTernaryMontgomeryNTT -- the ternary coefficients in Montgomery form, then NTT"""
        K = self.contextQP
        sk = signed_to_poly(K, ternary)
        K.MForm(sk, sk, stream=stream)
        K.NTT(sk, sk, stream=stream)
        return sk

    def GenPublicKey(self, sk, e, a, stream=None):
        """:138-151: pk[0] = -(sk*a + NTT(e)), pk[1] = a (uniform, taken as NTT + Montgomery)"""
        K = self.contextQP
        pk0 = signed_to_poly(K, e)
        K.NTT(pk0, pk0, stream=stream)  # SampleNTTNew
        pk1 = ring.Poly.from_numpy(np.ascontiguousarray(a)[None])
        K.MulCoeffsMontgomeryAndAdd(sk, pk1, pk0, stream=stream)
        K.Neg(pk0, pk0, stream=stream)
        return pk0, pk1

    def newSwitchingKey(self, skIn, skOut, errors, uniforms, stream=None):
        """:282-340.  errors[i]: gaussian coefficients of digit i, uniforms[i]: uniform poly over QP.
        skIn is consumed like the reference's polypool (multiplied by P in place)."""
        K = self.contextQP
        N, nQP = K.N, K.nl
        K.MulScalarBigint(skIn, self.Pbig, skIn, stream=stream)  # :290
        tmp = K.NewPoly()
        evk = np.zeros((self.beta, 2, nQP, N), dtype=np.uint64)
        for i in range(self.beta):
            k0 = signed_to_poly(K, errors[i])
            K.NTT(k0, k0, stream=stream)  # SampleNTTNew :303
            K.MForm(k0, k0, stream=stream)  # :304
            k1 = ring.Poly.from_numpy(np.ascontiguousarray(uniforms[i])[None])  # :307
            # :316-331  k0[index] = CRed(k0[index] + skIn[index]) on the digit's limbs below `levels`
            lo = i * self.alpha
            hi = min(lo + self.alpha, self.levels)
            tmp.Zero(stream=stream)
            K.CopyLvl(hi - lo - 1, skIn.view(lo, hi - lo), tmp.view(lo, hi - lo), stream=stream)
            K.Add(k0, tmp, k0, stream=stream)
            K.MulCoeffsMontgomeryAndSub(k1, skOut, k0, stream=stream)  # :334
            evk[i, 0] = k0.numpy(stream=stream)
            evk[i, 1] = k1.numpy(stream=stream)
        return SwitchingKey(evk), evk

    def GenRelinKey(self, sk, errors, uniforms, stream=None):
        """:190-203: switching key from sk^2 to sk"""
        K = self.contextQP
        pool = sk.CopyNew(stream=stream)
        K.MulCoeffsMontgomery(pool, sk, pool, stream=stream)
        return self.newSwitchingKey(pool, sk, errors, uniforms, stream=stream)

    def GenSwitchingKey(self, skInput, skOutput, errors, uniforms, stream=None):
        """:239-251"""
        return self.newSwitchingKey(skInput.CopyNew(stream=stream), skOutput, errors, uniforms, stream=stream)

    def genrotKey(self, skOutput, gen, errors, uniforms, stream=None):
        """:487-494: switching key from the Galois image of sk back to sk"""
        pool = self.contextQP.NewPoly()
        ring.PermuteNTT(skOutput, gen, pool, stream=stream)
        return self.newSwitchingKey(pool, skOutput, errors, uniforms, stream=stream)


class Encryptor:
    """pkEncryptor / skEncryptor (ckks/encryptor.go:13-362).  Plaintexts and ciphertexts are device polys over
    Q in the NTT domain (batch of independent messages); the QP pools are kept per batch size like the reference's
    encryptor.polypool (:16, :117-119) -- a synchronous cudaMalloc / cudaFree pair per call costs more than the ring ops.

    Like the reference, the non-fast paths hand the full QP pool to ModDownPQ(level, ...), which reads its
    "P part" at Coeffs[level+1 : level+1+#P] (ring_basis_extension.go:254): at the top level those are the
    special primes, below it they are the next Q limbs.  The call is kept literal."""

    def __init__(self, contextQ, contextP, contextQP, pk=None, sk=None):
        self.contextQ, self.contextP, self.contextQP = contextQ, contextP, contextQP
        self.pk, self.sk = pk, sk
        self.baseconverter = ring.NewFastBasisExtender(contextQ, contextP)
        self.nQ = contextQ.nl
        self._pools = {}

    def _pool(self, batch):
        if batch not in self._pools:
            self._pools[batch] = (self.contextQP.NewPoly(batch), self.contextQP.NewPoly(batch))
        return self._pools[batch]

    def EncryptPk(self, level, plaintext, ctOut, u, e0, e1, fast=False, stream=None):
        """pkEncryptor.encrypt :179-237.  u: ternary coefficients [batch][N]; e0, e1: gaussian coefficients.
        ct = [pk0*u + e0 (+m), pk1*u + e1], divided by P unless `fast`."""
        Q, K = self.contextQ, self.contextQP
        batch = plaintext.batch
        if fast:
            up = signed_to_poly(Q, u, batch)
            Q.MForm(up, up, stream=stream)
            Q.NTT(up, up, stream=stream)  # SampleTernaryMontgomeryNTT :187
            Q.MulCoeffsMontgomery(up, self.pk[0].view(0, self.nQ), ctOut[0], stream=stream)
            Q.MulCoeffsMontgomery(up, self.pk[1].view(0, self.nQ), ctOut[1], stream=stream)
            for e, c in ((e0, ctOut[0]), (e1, ctOut[1])):
                ep = signed_to_poly(Q, e, batch)
                Q.NTT(ep, ep, stream=stream)  # SampleNTT :195,:199
                Q.Add(c, ep, c, stream=stream)
        else:
            up = signed_to_poly(K, u, batch)
            K.MForm(up, up, stream=stream)
            K.NTT(up, up, stream=stream)  # :206
            p0, p1 = self._pool(batch)
            K.MulCoeffsMontgomery(up, self.pk[0], p0, stream=stream)  # :209
            K.MulCoeffsMontgomery(up, self.pk[1], p1, stream=stream)  # :211
            K.InvNTT(p0, p0, stream=stream)  # :214-215
            K.InvNTT(p1, p1, stream=stream)
            K.Add(p0, signed_to_poly(K, e0, batch), p0, stream=stream)  # SampleAndAdd :218,:220
            K.Add(p1, signed_to_poly(K, e1, batch), p1, stream=stream)
            self.baseconverter.ModDownPQ(level, p0, ctOut[0], stream=stream)  # :223
            self.baseconverter.ModDownPQ(level, p1, ctOut[1], stream=stream)  # :226
            Q.NTTLvl(level, ctOut[0], ctOut[0], stream=stream)  # :229-230
            Q.NTTLvl(level, ctOut[1], ctOut[1], stream=stream)
        Q.AddLvl(level, ctOut[0], plaintext, ctOut[0], stream=stream)  # :234

    def EncryptSk(self, level, plaintext, ctOut, crp, e, fast=False, stream=None):
        """skEncryptor.encrypt :318-362.  crp: uniform poly over QP (over Q when fast), NTT domain,
        [batch][limbs][N] residues; e: gaussian coefficients.  ct = [-a*s + e (+m), a]."""
        Q, K = self.contextQ, self.contextQP
        batch = plaintext.batch
        a = ring.Poly.from_numpy(np.ascontiguousarray(crp))
        if fast:
            Q.MulCoeffsMontgomery(a, self.sk.view(0, self.nQ), ctOut[0], stream=stream)  # :324
            Q.Neg(ctOut[0], ctOut[0], stream=stream)
            ep = signed_to_poly(Q, e, batch)
            Q.NTT(ep, ep, stream=stream)  # :327
            Q.Add(ctOut[0], ep, ctOut[0], stream=stream)
            Q.Copy(a, ctOut[1], stream=stream)  # :330
        else:
            p0, _ = self._pool(batch)
            K.MulCoeffsMontgomery(a, self.sk, p0, stream=stream)  # :337
            K.Neg(p0, p0, stream=stream)
            K.InvNTT(p0, p0, stream=stream)  # :341
            K.Add(p0, signed_to_poly(K, e, batch), p0, stream=stream)  # :344
            self.baseconverter.ModDownPQ(level, p0, ctOut[0], stream=stream)  # :348
            self.baseconverter.ModDownNTTPQ(level, a, ctOut[1], stream=stream)  # :352 (consumes the P part of a)
            Q.NTTLvl(level, ctOut[0], ctOut[0], stream=stream)  # :355
        Q.AddLvl(level, ctOut[0], plaintext, ctOut[0], stream=stream)  # :359


class Decryptor:
    """ckks/decryptor.go:53-78: Horner evaluation of the ciphertext at sk, ReduceLvl on the reference's cadence"""

    def __init__(self, contextQ, sk):
        self.contextQ, self.sk = contextQ, sk

    def Decrypt(self, level, ct, ptOut, stream=None):
        Q = self.contextQ
        sk = self.sk.view(0, Q.nl)
        degree = len(ct) - 1
        Q.CopyLvl(level, ct[degree], ptOut, stream=stream)
        for i in range(degree, 0, -1):
            Q.MulCoeffsMontgomeryLvl(level, ptOut, sk, ptOut, stream=stream)
            Q.AddLvl(level, ptOut, ct[i - 1], ptOut, stream=stream)
            if i & 7 == 7:
                Q.ReduceLvl(level, ptOut, ptOut, stream=stream)
        if degree & 7 != 7:
            Q.ReduceLvl(level, ptOut, ptOut, stream=stream)
