"""Host-side mirror of the ring-op sequences of the BFV key generator, encryptor, decryptor and batch encoder
(bfv/keygen.go:82-441, bfv/encryptor.go:168-345, bfv/decryptor.go:55-75, bfv/encoder.go:28-182) over the C ABI,
so that BASELINE config 3 runs "encode -> encrypt -> Mul -> Relinearize -> RotateColumns -> decrypt -> decode"
device-resident.

As in lattigpu.ckks_scheme, sampling stays with the caller (the reference draws from crypto/rand on the host): every
entry point takes the sampled values and runs what follows the sampling on the GPU.  BFV ciphertexts and plaintexts
live over Q in the COEFFICIENT domain; keys live over QP in NTT + Montgomery form.
"""
import ctypes as C

import numpy as np

from . import ring
from ._lib import check, lib, vp
from .bfv import GaloisGen
from .ckks import SwitchingKey
from .ckks_scheme import signed_to_poly
from .ring import _arr, _ptr, _s


class KeyGenerator:
    """bfv.keyGenerator over contextQP (bfv/keygen.go:60-80)"""

    def __init__(self, contextQ, contextP):
        self.contextQ, self.contextP = contextQ, contextP
        self.contextQP = ring.NewContextWithParams(contextQ.N, list(contextQ.Modulus) + list(contextP.Modulus))
        self.nQ, self.alpha = contextQ.nl, contextP.nl
        self.beta = -(-self.nQ // self.alpha)
        self.Pbig = 1
        for p in contextP.Modulus:
            self.Pbig *= int(p)

    def GenSecretKey(self, ternary, stream=None):
        """:82-96: SampleTernaryMontgomeryNTTNew"""
        K = self.contextQP
        sk = signed_to_poly(K, ternary)
        K.MForm(sk, sk, stream=stream)
        K.NTT(sk, sk, stream=stream)
        return sk

    def GenPublicKey(self, sk, e, a, stream=None):
        """:120-135: pk[0] = -(sk*a + NTT(e)), pk[1] = a"""
        K = self.contextQP
        pk0 = signed_to_poly(K, e)
        K.NTT(pk0, pk0, stream=stream)
        pk1 = ring.Poly.from_numpy(np.ascontiguousarray(a)[None])
        K.MulCoeffsMontgomeryAndAdd(sk, pk1, pk0, stream=stream)
        K.Neg(pk0, pk0, stream=stream)
        return pk0, pk1

    def newswitchingkey(self, skIn, skOut, errors, uniforms, stream=None):
        """:285-334.  skIn: already multiplied by P (the callers do it).  The in-digit loop stops at the last limb of
        the QP context (:323), not of Q -- kept literal."""
        K = self.contextQP
        N, nQP = K.N, K.nl
        tmp = K.NewPoly()
        evk = np.zeros((self.beta, 2, nQP, N), dtype=np.uint64)
        for i in range(self.beta):
            k0 = signed_to_poly(K, errors[i])
            K.NTT(k0, k0, stream=stream)  # SampleNTTNew :301
            K.MForm(k0, k0, stream=stream)  # :302
            k1 = ring.Poly.from_numpy(np.ascontiguousarray(uniforms[i])[None])  # :304
            lo = i * self.alpha
            hi = min(lo + self.alpha, nQP)  # :309-326
            tmp.Zero(stream=stream)
            K.CopyLvl(hi - lo - 1, skIn.view(lo, hi - lo), tmp.view(lo, hi - lo), stream=stream)
            K.Add(k0, tmp, k0, stream=stream)
            K.MulCoeffsMontgomeryAndSub(k1, skOut, k0, stream=stream)  # :331
            evk[i, 0] = k0.numpy(stream=stream)
            evk[i, 1] = k1.numpy(stream=stream)
        return SwitchingKey(evk), evk

    def GenRelinKey(self, sk, errors, uniforms, stream=None):
        """:171-195 with maxDegree = 1: key from P*sk^2 to sk"""
        K = self.contextQP
        pool = sk.CopyNew(stream=stream)
        K.MulScalarBigint(pool, self.Pbig, pool, stream=stream)
        K.MulCoeffsMontgomery(pool, sk, pool, stream=stream)
        return self.newswitchingkey(pool, sk, errors, uniforms, stream=stream)

    def GenSwitchingKey(self, skIn, skOut, errors, uniforms, stream=None):
        """:248-262"""
        K = self.contextQP
        pool = K.NewPoly()
        K.MulScalarBigint(skIn, self.Pbig, pool, stream=stream)
        return self.newswitchingkey(pool, skOut, errors, uniforms, stream=stream)

    def genrotkey(self, sk, gen, errors, uniforms, stream=None):
        """:429-441"""
        K = self.contextQP
        pool = K.NewPoly()
        ring.PermuteNTT(sk, gen, pool, stream=stream)
        K.MulScalarBigint(pool, self.Pbig, pool, stream=stream)
        return self.newswitchingkey(pool, sk, errors, uniforms, stream=stream)


class Encryptor:
    """pkEncryptor / skEncryptor (bfv/encryptor.go:57-345); batches of independent plaintexts"""

    def __init__(self, contextQ, contextP, contextQP, pk=None, sk=None):
        self.contextQ, self.contextP, self.contextQP = contextQ, contextP, contextQP
        self.pk, self.sk = pk, sk
        self.baseconverter = ring.NewFastBasisExtender(contextQ, contextP)
        self.nQ = contextQ.nl

    def EncryptPk(self, plaintext, ctOut, u, e0, e1, fast=False, stream=None):
        """pkEncryptor.encrypt :168-222.  The fast branch (:174-192) leaves its result in the encryptor's pools and
        never copies it to the ciphertext, so the receiver only has the plaintext added (:221); mirrored as is."""
        Q, K = self.contextQ, self.contextQP
        batch = plaintext.batch
        level = self.nQ - 1
        if not fast:
            up = signed_to_poly(K, u, batch)
            K.MForm(up, up, stream=stream)
            K.NTT(up, up, stream=stream)  # :196
            p0, p1 = K.NewPoly(batch), K.NewPoly(batch)
            K.MulCoeffsMontgomery(up, self.pk[0], p0, stream=stream)  # :200-201
            K.MulCoeffsMontgomery(up, self.pk[1], p1, stream=stream)
            K.InvNTT(p0, p0, stream=stream)  # :203-204
            K.InvNTT(p1, p1, stream=stream)
            K.Add(p0, signed_to_poly(K, e0, batch), p0, stream=stream)  # :207-212
            K.Add(p1, signed_to_poly(K, e1, batch), p1, stream=stream)
            self.baseconverter.ModDownPQ(level, p0, ctOut[0], stream=stream)  # :215-216
            self.baseconverter.ModDownPQ(level, p1, ctOut[1], stream=stream)
        Q.Add(ctOut[0], plaintext, ctOut[0], stream=stream)  # :221

    def EncryptSk(self, plaintext, ctOut, crp, e, fast=False, stream=None):
        """skEncryptor.encrypt :296-345.  crp: uniform poly over QP (over Q when fast), NTT domain residues
        [batch][limbs][N]; e: gaussian coefficients.  ct = [-a*s + e (+m), a] in the coefficient domain."""
        Q, K = self.contextQ, self.contextQP
        batch = plaintext.batch
        level = self.nQ - 1
        a = ring.Poly.from_numpy(np.ascontiguousarray(crp))
        if fast:
            Q.MulCoeffsMontgomery(a, self.sk.view(0, self.nQ), ctOut[0], stream=stream)  # :304
            Q.Neg(ctOut[0], ctOut[0], stream=stream)
            Q.InvNTT(ctOut[0], ctOut[0], stream=stream)  # :307-308
            Q.InvNTT(a, ctOut[1], stream=stream)
            Q.Add(ctOut[0], signed_to_poly(Q, e, batch), ctOut[0], stream=stream)  # SampleGaussianAndAdd :310
        else:
            p0 = K.NewPoly(batch)
            K.MulCoeffsMontgomery(a, self.sk, p0, stream=stream)  # :316
            K.Neg(p0, p0, stream=stream)
            K.InvNTT(p0, p0, stream=stream)  # :320-321
            K.InvNTT(a, a, stream=stream)
            K.Add(p0, signed_to_poly(K, e, batch), p0, stream=stream)  # :323
            self.baseconverter.ModDownPQ(level, p0, ctOut[0], stream=stream)  # :325-326
            self.baseconverter.ModDownPQ(level, a, ctOut[1], stream=stream)
        Q.Add(ctOut[0], plaintext, ctOut[0], stream=stream)  # :344


class Decryptor:
    """bfv/decryptor.go:55-75: Horner evaluation at sk in the NTT domain, back to coefficients"""

    def __init__(self, contextQ, sk):
        self.contextQ, self.sk = contextQ, sk

    def Decrypt(self, ct, ptOut, stream=None):
        Q = self.contextQ
        sk = self.sk.view(0, Q.nl)
        degree = len(ct) - 1
        pool = Q.NewPoly(ptOut.batch)
        Q.NTT(ct[degree], ptOut, stream=stream)
        for i in range(degree, 0, -1):
            Q.MulCoeffsMontgomery(ptOut, sk, ptOut, stream=stream)
            Q.NTT(ct[i - 1], pool, stream=stream)
            Q.Add(ptOut, pool, ptOut, stream=stream)
            if i & 7 == 7:
                Q.Reduce(ptOut, ptOut, stream=stream)
        if degree & 7 != 7:
            Q.Reduce(ptOut, ptOut, stream=stream)
        Q.InvNTT(ptOut, ptOut, stream=stream)


def index_matrix(N):
    """bfv/encoder.go:36-58: slot i of row 0 / row 1 sits at bit-reversed position of (5^i - 1)/2 / (2N - 5^i - 1)/2"""
    logN = N.bit_length() - 1
    rev = lambda x: int(format(x, "0%db" % logN)[::-1], 2) if logN else 0
    row, m, pos = N >> 1, N << 1, 1
    idx = np.zeros(N, dtype=np.uint64)
    for i in range(row):
        idx[i] = rev((pos - 1) >> 1)
        idx[i | row] = rev((m - pos - 1) >> 1)
        pos = (pos * GaloisGen) & (m - 1)
    return idx


class Encoder:
    """bfv.encoder (bfv/encoder.go:18-182) on the device: slots are [batch][1][N] polys of values below t.
    encode = scatter through indexMatrix, InvNTT over contextT, lift by Delta = floor(Q/t) into every limb of Q;
    decode = SimpleScaler.Scale, NTT over contextT, gather through indexMatrix."""

    def __init__(self, contextQ, t):
        self.contextQ, self.t = contextQ, int(t)
        self.contextT = ring.NewContextWithParams(contextQ.N, [self.t])  # bfv/bfv.go:47
        self.indexMatrix = index_matrix(contextQ.N)
        inv = np.zeros(contextQ.N, dtype=np.uint64)
        inv[self.indexMatrix.astype(np.int64)] = np.arange(contextQ.N, dtype=np.uint64)
        self._gather = ring.GaloisIndex(index=self.indexMatrix)  # coeffs[i] = pool[indexMatrix[i]]
        self._scatter = ring.GaloisIndex(index=inv)  # pt[indexMatrix[i]] = coeffs[i]
        self.simplescaler = ring.NewSimpleScaler(self.t, contextQ)
        h = vp()
        check(lib().lg_bfv_lift_create(contextQ.h, self.t, C.byref(h)))
        self._lift = h

    def __del__(self):
        try:
            lib().lg_bfv_lift_destroy(self._lift)
        except Exception:
            pass

    def deltaMont(self):
        out = np.zeros(self.contextQ.nl, np.uint64)
        check(lib().lg_bfv_lift_get_params(self._lift, _ptr(out)))
        return out

    def slots_from_host(self, coeffs):
        """[batch][n <= N] unsigned (or signed: EncodeInt :94-119 maps x < 0 to t + x) values -> device slots,
        zero-padded (:84-86)"""
        c = np.asarray(coeffs)
        if c.ndim == 1:
            c = c[None]
        if c.shape[1] > self.contextQ.N:
            raise ValueError("cannot EncodeUint: invalid input to encode (number of coefficients must be smaller or equal to the context)")
        if np.issubdtype(c.dtype, np.signedinteger):
            c = np.where(c < 0, c + self.t, c)
        full = np.zeros((c.shape[0], 1, self.contextQ.N), dtype=np.uint64)
        full[:, 0, : c.shape[1]] = c.astype(np.uint64)
        return ring.Poly.from_numpy(full)

    def EncodeUint(self, slots, plaintext, stream=None):
        """:69-90 + encodePlaintext :121-136.  slots: device poly [batch][1][N] (see slots_from_host)"""
        if not isinstance(slots, ring.Poly):
            slots = self.slots_from_host(slots)
        m = ring.Poly(self.contextQ.N, 1, slots.batch)
        ring.PermuteNTTWithIndex(slots, self._scatter, m, stream=stream)
        self.contextT.InvNTT(m, m, stream=stream)
        check(lib().lg_bfv_lift_apply(self._lift, m.h, plaintext.h, _s(stream)))

    EncodeInt = EncodeUint

    def DecodeUintDevice(self, plaintext, stream=None):
        """:139-153 up to the host copy: returns the device poly [batch][1][N] of slot values"""
        pool = ring.Poly(self.contextQ.N, 1, plaintext.batch)
        self.simplescaler.Scale(plaintext, pool, stream=stream)
        self.contextT.NTT(pool, pool, stream=stream)
        out = ring.Poly(self.contextQ.N, 1, plaintext.batch)
        ring.PermuteNTTWithIndex(pool, self._gather, out, stream=stream)
        return out

    def DecodeUint(self, plaintext, stream=None):
        return self.DecodeUintDevice(plaintext, stream=stream).numpy(stream=stream, squeeze=False)[:, 0, :]

    def DecodeInt(self, plaintext, stream=None):
        """:157-182: centred around t"""
        v = self.DecodeUint(plaintext, stream=stream).astype(np.int64)
        return np.where(v > (self.t >> 1), v - self.t, v)
