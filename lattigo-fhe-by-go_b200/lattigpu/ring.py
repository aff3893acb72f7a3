"""Host-side mirror of the reference's `ring` package interface over the C ABI.

Method names, argument order ("output last") and level semantics follow
ring/ring_context.go, ring/ring.go, ring/ntt.go, ring/ring_galois.go,
ring/ring_scaling.go and ring/ring_basis_extension.go of Lattigo v1.3.1, so that
tests read like the reference's.  Everything executes on the GPU through
liblattigpu.so; this module holds no arithmetic.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, lib, p64, u64, vp


def _arr(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64))


def _ptr(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(p64)


def device_count():
    n = C.c_int(0)
    lib().lg_device_count(C.byref(n))
    return n.value


def set_device(i):
    check(lib().lg_set_device(i))


def launch_count():
    return int(lib().lg_launch_count())


def debug_set_switch(name, value):
    """lg_debug_set_switch: diagnostic A/B switches between kernel variants that return identical words"""
    check(lib().lg_debug_set_switch(name.encode(), int(value)))


def GenerateNTTPrimes(logQ, logN, levels):
    """ring/utils.go:133-175 (host only)"""
    out = np.zeros(levels, dtype=np.uint64)
    check(lib().lg_generate_ntt_primes(logQ, logN, levels, _ptr(out)))
    return [int(x) for x in out]


def IsPrime(num):
    return bool(lib().lg_is_prime(num))


class Stream:
    def __init__(self, handle=None):
        if handle is None:
            h = vp()
            check(lib().lg_stream_create(C.byref(h)))
            self.h, self.owns = h, True
        else:
            self.h, self.owns = vp(handle), False

    def sync(self):
        check(lib().lg_stream_sync(self.h))

    def __del__(self):
        try:
            if self.owns and self.h:
                lib().lg_stream_destroy(self.h)
        except Exception:
            pass


_default_stream = vp(None)


def _s(stream):
    if stream is None:
        return _default_stream
    return stream.h if isinstance(stream, Stream) else vp(stream)


class Poly:
    """ring.Poly resident on the device, laid out [batch][nlimbs][N] (ring/ring_object.go:11-13)."""

    def __init__(self, N=None, nlimbs=None, batch=1, _handle=None, _keep=None):
        if _handle is None:
            h = vp()
            check(lib().lg_poly_create(N, nlimbs, batch, C.byref(h)))
            _handle = h
        self.h = _handle
        self._keep = _keep
        L = lib()
        self.N = int(L.lg_poly_n(self.h))
        self.nlimbs = int(L.lg_poly_nlimbs(self.h))
        self.batch = int(L.lg_poly_batch(self.h))

    def __del__(self):
        try:
            if self.h:
                lib().lg_poly_destroy(self.h)
                self.h = None
        except Exception:
            pass

    @classmethod
    def from_numpy(cls, a, stream=None):
        """a: [nlimbs, N] or [batch, nlimbs, N] uint64"""
        a = _arr(a)
        if a.ndim == 2:
            a = a[None]
        p = cls(a.shape[2], a.shape[1], a.shape[0])
        p.set(a, stream=stream)
        return p

    @classmethod
    def wrap(cls, device_ptr, N, nlimbs, batch, keep=None):
        h = vp()
        check(lib().lg_poly_wrap(vp(device_ptr), N, nlimbs, batch, C.byref(h)))
        return cls(_handle=h, _keep=keep)

    def view(self, limb0, nlimbs):
        """p.Coeffs[limb0:limb0+nlimbs] as a non-owning handle"""
        h = vp()
        check(lib().lg_poly_view(self.h, limb0, nlimbs, C.byref(h)))
        return Poly(_handle=h, _keep=self)

    def set(self, a, limb0=0, batch0=0, stream=None):
        a = _arr(a)
        if a.ndim == 2:
            a = a[None]
        check(lib().lg_poly_upload(self.h, batch0, a.shape[0], limb0, a.shape[1], _ptr(a), _s(stream)))

    def numpy(self, nl=None, limb0=0, stream=None, squeeze=True):
        nl = self.nlimbs - limb0 if nl is None else nl
        out = np.empty((self.batch, nl, self.N), dtype=np.uint64)
        check(lib().lg_poly_download(self.h, 0, self.batch, limb0, nl, _ptr(out), _s(stream)))
        return out[0] if (squeeze and self.batch == 1) else out

    # ring/ring_object.go:146-289 -- wire format (2-byte header, big-endian words, limb-major)
    def GetDataLen(self, WithMetadata=True, nl=None):
        return int(lib().lg_poly_get_data_len(self.h, self.nlimbs if nl is None else nl, 1 if WithMetadata else 0))

    def MarshalBinary(self, batch_index=0, nl=None, WithMetadata=True, stream=None):
        """WriteTo / MarshalBinary (:161-175, :224-231); WithMetadata=False is WriteCoeffs (:177-184)"""
        nl = self.nlimbs if nl is None else nl
        buf = C.create_string_buffer(self.GetDataLen(WithMetadata, nl))
        check(lib().lg_poly_write_to(self.h, batch_index, nl, buf, len(buf), 1 if WithMetadata else 0, _s(stream)))
        return buf.raw

    def UnmarshalBinary(self, data, batch_index=0, WithMetadata=True, nl=None, stream=None):
        """UnmarshalBinary / DecodePolyNew (:257-289); WithMetadata=False is DecodeCoeffs over `nl` moduli"""
        data = bytes(data)
        check(lib().lg_poly_decode(self.h, batch_index, data, len(data), 1 if WithMetadata else 0,
                                   self.nlimbs if nl is None else nl, _s(stream)))

    def device_ptr(self):
        return lib().lg_poly_device_ptr(self.h)

    def Zero(self, stream=None):
        check(lib().lg_poly_zero(self.h, _s(stream)))

    def CopyNew(self, stream=None):
        q = Poly(self.N, self.nlimbs, self.batch)
        check(lib().lg_poly_copy(self.h, self.nlimbs, q.h, _s(stream)))
        return q


def _op3(name):
    def full(self, p1, p2, p3, stream=None):
        check(getattr(lib(), name)(self.h, self.nl, p1.h, p2.h, p3.h, _s(stream)))

    def lvl(self, level, p1, p2, p3, stream=None):
        check(getattr(lib(), name)(self.h, level + 1, p1.h, p2.h, p3.h, _s(stream)))

    return full, lvl


def _op2(name):
    def full(self, p1, p2, stream=None):
        check(getattr(lib(), name)(self.h, self.nl, p1.h, p2.h, _s(stream)))

    def lvl(self, level, p1, p2, stream=None):
        check(getattr(lib(), name)(self.h, level + 1, p1.h, p2.h, _s(stream)))

    return full, lvl


class Context:
    """ring.Context (ring/ring_context.go:18-51).  `NewContextWithParams(N, Moduli)`
    generates the NTT tables natively; `from_tables` takes the ones Go computed."""

    def __init__(self, N, Modulus, _handle=None):
        self.N = int(N)
        self.Modulus = [int(q) for q in Modulus]
        self.nl = len(self.Modulus)
        if _handle is None:
            h = vp()
            m = _arr(self.Modulus)
            check(lib().lg_ring_create(self.N, self.nl, _ptr(m), C.byref(h)))
            _handle = h
        self.h = _handle

    @classmethod
    def from_tables(cls, N, Modulus, bred, mred, psi, psi_inv, ninv, rescale=None):
        h = vp()
        m = _arr(Modulus)
        arrs = [_arr(x) for x in (bred, mred, psi, psi_inv, ninv)]
        res = _arr(rescale) if rescale is not None and len(rescale) else None
        check(lib().lg_ring_create_from_tables(int(N), len(m), _ptr(m), *[_ptr(a) for a in arrs],
                                               _ptr(res) if res is not None else None, C.byref(h)))
        return cls(N, Modulus, _handle=h)

    def __del__(self):
        try:
            if self.h:
                lib().lg_ring_destroy(self.h)
                self.h = None
        except Exception:
            pass

    def tables(self):
        nl, N = self.nl, self.N
        out = dict(moduli=np.zeros(nl, np.uint64), bred=np.zeros((nl, 2), np.uint64), mred=np.zeros(nl, np.uint64),
                   psi=np.zeros((nl, N), np.uint64), psi_inv=np.zeros((nl, N), np.uint64), ninv=np.zeros(nl, np.uint64),
                   rescale=np.zeros(max(nl * (nl - 1) // 2, 1), np.uint64))
        check(lib().lg_ring_get_tables(self.h, *[_ptr(out[k]) for k in
                                                 ("moduli", "bred", "mred", "psi", "psi_inv", "ninv", "rescale")]))
        return out

    def NewPoly(self, batch=1):
        return Poly(self.N, self.nl, batch)

    def NewPolyLvl(self, level, batch=1):
        return Poly(self.N, level + 1, batch)

    # ring/ntt.go:4-29
    NTT, NTTLvl = _op2("lg_ring_ntt")
    InvNTT, InvNTTLvl = _op2("lg_ring_invntt")
    # ring/ring.go
    Add, AddLvl = _op3("lg_ring_add")
    AddNoMod, AddNoModLvl = _op3("lg_ring_add_nomod")
    Sub, SubLvl = _op3("lg_ring_sub")
    SubNoMod, SubNoModLvl = _op3("lg_ring_sub_nomod")
    Neg, NegLvl = _op2("lg_ring_neg")
    Reduce, ReduceLvl = _op2("lg_ring_reduce")
    MulCoeffs, MulCoeffsLvl = _op3("lg_ring_mul_coeffs")
    MulCoeffsAndAdd, _ = _op3("lg_ring_mul_coeffs_and_add")
    MulCoeffsAndAddNoMod, _ = _op3("lg_ring_mul_coeffs_and_add_nomod")
    MulCoeffsConstant, _ = _op3("lg_ring_mul_coeffs_constant")
    MulCoeffsMontgomery, MulCoeffsMontgomeryLvl = _op3("lg_ring_mul_coeffs_montgomery")
    MulCoeffsMontgomeryAndAdd, MulCoeffsMontgomeryAndAddLvl = _op3("lg_ring_mul_coeffs_montgomery_and_add")
    MulCoeffsMontgomeryAndAddNoMod, MulCoeffsMontgomeryAndAddNoModLvl = _op3("lg_ring_mul_coeffs_montgomery_and_add_nomod")
    _, MulCoeffsMontgomeryConstantAndAddNoModLvl = _op3("lg_ring_mul_coeffs_montgomery_constant_and_add_nomod")
    MulCoeffsMontgomeryAndSub, _ = _op3("lg_ring_mul_coeffs_montgomery_and_sub")
    MulCoeffsMontgomeryAndSubNoMod, _ = _op3("lg_ring_mul_coeffs_montgomery_and_sub_nomod")
    MulCoeffsMontgomeryConstant, _ = _op3("lg_ring_mul_coeffs_montgomery_constant")
    MForm, MFormLvl = _op2("lg_ring_mform")
    InvMForm, _ = _op2("lg_ring_invmform")
    BitReverse, _ = _op2("lg_ring_bitreverse")

    # ring/ring.go:357-437, :439-464, :574-580, :772-800; ring/ring_context.go:423-467 (csrc/ringext.cu)
    def MulPoly(self, p1, p2, p3, stream=None):
        check(lib().lg_ring_mul_poly(self.h, p1.h, p2.h, p3.h, 0, _s(stream)))

    def MulPolyMontgomery(self, p1, p2, p3, stream=None):
        check(lib().lg_ring_mul_poly(self.h, p1.h, p2.h, p3.h, 1, _s(stream)))

    def MulPolyNaive(self, p1, p2, p3, stream=None):
        check(lib().lg_ring_mul_poly_naive(self.h, p1.h, p2.h, p3.h, 0, _s(stream)))

    def MulPolyNaiveMontgomery(self, p1, p2, p3, stream=None):
        check(lib().lg_ring_mul_poly_naive(self.h, p1.h, p2.h, p3.h, 1, _s(stream)))

    def Exp(self, p1, e, p2, stream=None):
        check(lib().lg_ring_exp(self.h, p1.h, u64(e), p2.h, _s(stream)))

    def Shift(self, p1, n, p2, stream=None):
        check(lib().lg_ring_shift(self.h, p1.h, u64(n), p2.h, _s(stream)))

    def Rotate(self, p1, n, p2=None, stream=None):
        """p2 is accepted and ignored, as the reference never writes it (ring.go:791)"""
        check(lib().lg_ring_rotate(self.h, p1.h, u64(n), _s(stream)))

    def Equal(self, p1, p2, stream=None):
        return self.EqualLvl(self.nl - 1, p1, p2, stream)

    def EqualLvl(self, level, p1, p2, stream=None):
        import ctypes

        eq = ctypes.c_int(0)
        check(lib().lg_ring_equal(self.h, level + 1, p1.h, p2.h, ctypes.byref(eq), _s(stream)))
        return bool(eq.value)

    def _word(self, name, p1, m, p2, stream=None):
        check(getattr(lib(), name)(self.h, self.nl, p1.h, u64(m), p2.h, _s(stream)))

    def Mod(self, p1, m, p2, stream=None):
        self._word("lg_ring_mod", p1, m, p2, stream)

    def AND(self, p1, m, p2, stream=None):
        self._word("lg_ring_and", p1, m, p2, stream)

    def OR(self, p1, m, p2, stream=None):
        self._word("lg_ring_or", p1, m, p2, stream)

    def XOR(self, p1, m, p2, stream=None):
        self._word("lg_ring_xor", p1, m, p2, stream)

    # scalar ops: the big.Int reduction mod q_i stays on the host, as in ring.go:477-572
    def _scalars(self, scalar, nl):
        return _arr([int(scalar) % (1 << 64)] * nl)

    def _bigint(self, scalar, nl):
        return _arr([int(scalar) % q for q in self.Modulus[:nl]])

    def AddScalar(self, p1, scalar, p2=None, stream=None):
        s = self._scalars(scalar, self.nl)
        check(lib().lg_ring_add_scalar(self.h, self.nl, p1.h, _ptr(s), _s(stream)))

    def AddScalarBigint(self, p1, scalar, p2=None, stream=None):
        s = self._bigint(scalar, self.nl)
        check(lib().lg_ring_add_scalar(self.h, self.nl, p1.h, _ptr(s), _s(stream)))

    def SubScalar(self, p1, scalar, p2=None, stream=None):
        s = self._scalars(scalar, self.nl)
        check(lib().lg_ring_sub_scalar(self.h, self.nl, p1.h, _ptr(s), _s(stream)))

    def SubScalarBigint(self, p1, scalar, p2=None, stream=None):
        s = self._bigint(scalar, self.nl)
        check(lib().lg_ring_sub_scalar(self.h, self.nl, p1.h, _ptr(s), _s(stream)))

    def MulScalar(self, p1, scalar, p2, stream=None):
        self.MulScalarLvl(self.nl - 1, p1, scalar, p2, stream)

    def MulScalarLvl(self, level, p1, scalar, p2, stream=None):
        s = self._scalars(scalar, level + 1)
        check(lib().lg_ring_mul_scalar(self.h, level + 1, p1.h, _ptr(s), p2.h, _s(stream)))

    def MulScalarBigint(self, p1, scalar, p2, stream=None):
        self.MulScalarBigintLvl(self.nl - 1, p1, scalar, p2, stream)

    def MulScalarBigintLvl(self, level, p1, scalar, p2, stream=None):
        s = self._bigint(scalar, level + 1)
        check(lib().lg_ring_mul_scalar(self.h, level + 1, p1.h, _ptr(s), p2.h, _s(stream)))

    def MulByPow2(self, p1, pow2, p2, stream=None):
        check(lib().lg_ring_mul_by_pow2(self.h, self.nl, p1.h, pow2, p2.h, _s(stream)))

    def MulByPow2Lvl(self, level, p1, pow2, p2, stream=None):
        check(lib().lg_ring_mul_by_pow2(self.h, level + 1, p1.h, pow2, p2.h, _s(stream)))

    def MultByMonomial(self, p1, monomialDeg, p2, stream=None):
        check(lib().lg_ring_mult_by_monomial(self.h, self.nl, p1.h, monomialDeg, p2.h, _s(stream)))

    def MulByVectorMontgomery(self, p1, vector, p2, stream=None):
        check(lib().lg_ring_mul_by_vector_montgomery(self.h, self.nl, p1.h, vector.h, p2.h, _s(stream)))

    def MulByVectorMontgomeryAndAddNoMod(self, p1, vector, p2, stream=None):
        check(lib().lg_ring_mul_by_vector_montgomery_and_add_nomod(self.h, self.nl, p1.h, vector.h, p2.h, _s(stream)))

    def Copy(self, p0, p1, stream=None):
        check(lib().lg_poly_copy(p0.h, self.nl, p1.h, _s(stream)))

    def CopyLvl(self, level, p0, p1, stream=None):
        check(lib().lg_poly_copy(p0.h, level + 1, p1.h, _s(stream)))

    # ring/ring_galois.go:106-127
    def Permute(self, polIn, gen, polOut, stream=None):
        check(lib().lg_ring_permute(self.h, self.nl, polIn.h, gen, polOut.h, _s(stream)))

    # ring/ring_scaling.go -- `nl` = len(p0.Coeffs); the caller drops the last limb(s)
    def _div(self, name, p0, nl, stream, nb=None):
        nl = p0.nlimbs if nl is None else nl
        if nb is None:
            check(getattr(lib(), name)(self.h, nl, p0.h, _s(stream)))
        else:
            check(getattr(lib(), name)(self.h, nl, p0.h, nb, _s(stream)))

    def DivFloorByLastModulusNTT(self, p0, nl=None, stream=None):
        self._div("lg_ring_div_floor_by_last_modulus_ntt", p0, nl, stream)

    def DivFloorByLastModulus(self, p0, nl=None, stream=None):
        self._div("lg_ring_div_floor_by_last_modulus", p0, nl, stream)

    def DivFloorByLastModulusManyNTT(self, p0, nbRescales, nl=None, stream=None):
        self._div("lg_ring_div_floor_by_last_modulus_many_ntt", p0, nl, stream, nbRescales)

    def DivFloorByLastModulusMany(self, p0, nbRescales, nl=None, stream=None):
        self._div("lg_ring_div_floor_by_last_modulus_many", p0, nl, stream, nbRescales)

    def DivRoundByLastModulusNTT(self, p0, nl=None, stream=None):
        self._div("lg_ring_div_round_by_last_modulus_ntt", p0, nl, stream)

    def DivRoundByLastModulus(self, p0, nl=None, stream=None):
        self._div("lg_ring_div_round_by_last_modulus", p0, nl, stream)

    def DivRoundByLastModulusManyNTT(self, p0, nbRescales, nl=None, stream=None):
        self._div("lg_ring_div_round_by_last_modulus_many_ntt", p0, nl, stream, nbRescales)

    def DivRoundByLastModulusMany(self, p0, nbRescales, nl=None, stream=None):
        self._div("lg_ring_div_round_by_last_modulus_many", p0, nl, stream, nbRescales)


def NewContextWithParams(N, Moduli):
    """ring_context.go:60-64: raises LattigpuError when the moduli do not allow the NTT."""
    return Context(N, Moduli)


def NTT(context, table_limb, coeffsIn, limbIn, coeffsOut, limbOut, stream=None):
    """free function ring.NTT (ntt.go:53) on one limb of a Poly with the tables of `table_limb`"""
    check(lib().lg_ring_ntt_limb(context.h, table_limb, coeffsIn.h, limbIn, coeffsOut.h, limbOut, _s(stream)))


def InvNTT(context, table_limb, coeffsIn, limbIn, coeffsOut, limbOut, stream=None):
    check(lib().lg_ring_invntt_limb(context.h, table_limb, coeffsIn.h, limbIn, coeffsOut.h, limbOut, _s(stream)))


class GaloisIndex:
    """the []uint64 returned by ring.PermuteNTTIndex (ring_galois.go:29-50), device resident"""

    def __init__(self, gen=None, power=None, N=None, index=None):
        h = vp()
        if index is not None:
            idx = _arr(index)
            check(lib().lg_galois_create_from_index(_ptr(idx), len(idx), C.byref(h)))
            self.N = len(idx)
        else:
            check(lib().lg_galois_create(gen, power, N, C.byref(h)))
            self.N = N
        self.h = h

    def numpy(self):
        out = np.zeros(self.N, np.uint64)
        check(lib().lg_galois_get_index(self.h, _ptr(out)))
        return out

    def __del__(self):
        try:
            lib().lg_galois_destroy(self.h)
        except Exception:
            pass


def PermuteNTTIndex(gen, power, N):
    return GaloisIndex(gen, power, N)


def PermuteNTTWithIndex(polIn, index, polOut, stream=None):
    check(lib().lg_ring_permute_ntt_with_index(min(polIn.nlimbs, polOut.nlimbs), polIn.h, index.h, polOut.h, _s(stream)))


def PermuteNTT(polIn, gen, polOut, stream=None):
    check(lib().lg_ring_permute_ntt(min(polIn.nlimbs, polOut.nlimbs), polIn.h, gen, polOut.h, _s(stream)))


class FastBasisExtender:
    """ring.FastBasisExtender (ring_basis_extension.go:9-350)"""

    def __init__(self, contextQ, contextP):
        self.contextQ, self.contextP = contextQ, contextP
        h = vp()
        check(lib().lg_extender_create(contextQ.h, contextP.h, C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            lib().lg_extender_destroy(self.h)
        except Exception:
            pass

    def ModUpSplitQP(self, level, p1, p2, stream=None):
        check(lib().lg_extender_modup_split_qp(self.h, level, p1.h, p2.h, _s(stream)))

    def ModUpSplitPQ(self, level, p1, p2, stream=None):
        check(lib().lg_extender_modup_split_pq(self.h, level, p1.h, p2.h, _s(stream)))

    def ModDownNTTPQ(self, level, p1, p2, stream=None):
        check(lib().lg_extender_moddown_ntt_pq(self.h, level, p1.h, p2.h, _s(stream)))

    def ModDownSplitedNTTPQ(self, level, p1Q, p1P, p2, stream=None):
        check(lib().lg_extender_moddown_splited_ntt_pq(self.h, level, p1Q.h, p1P.h, p2.h, _s(stream)))

    def ModDownPQ(self, level, p1, p2, stream=None):
        check(lib().lg_extender_moddown_pq(self.h, level, p1.h, p2.h, _s(stream)))

    def ModDownSplitedPQ(self, level, p1Q, p1P, p2, stream=None):
        check(lib().lg_extender_moddown_splited_pq(self.h, level, p1Q.h, p1P.h, p2.h, _s(stream)))

    def ModDownSplitedQP(self, levelQ, levelP, p1Q, p1P, p2, stream=None):
        check(lib().lg_extender_moddown_splited_qp(self.h, levelQ, levelP, p1Q.h, p1P.h, p2.h, _s(stream)))


def NewFastBasisExtender(contextQ, contextP):
    return FastBasisExtender(contextQ, contextP)


class Decomposer:
    """ring.Decomposer (ring_basis_extension.go:398-713)"""

    def __init__(self, N, Q, P):
        q, p = _arr(Q), _arr(P)
        h = vp()
        check(lib().lg_decomposer_create(N, _ptr(q), len(q), _ptr(p), len(p), C.byref(h)))
        self.h = h
        self.beta = lib().lg_decomposer_beta(h)

    def __del__(self):
        try:
            lib().lg_decomposer_destroy(self.h)
        except Exception:
            pass

    def Xalpha(self):
        return [lib().lg_decomposer_xalpha(self.h, i) for i in range(self.beta)]

    def Decompose(self, level, crtDecompLevel, p0, p1, stream=None):
        check(lib().lg_decomposer_decompose(self.h, level, crtDecompLevel, p0.h, p1.h, _s(stream)))

    def DecomposeAndSplit(self, level, crtDecompLevel, p0, p1Q, p1P, stream=None):
        check(lib().lg_decomposer_decompose_and_split(self.h, level, crtDecompLevel, p0.h, p1Q.h, p1P.h, _s(stream)))


def NewDecomposer(N, Q, P):
    return Decomposer(N, Q, P)


class PRNG:
    """utils.PRNG (utils/prng.go:11-72): keyed BLAKE2b-512 hash chain; host code in the library"""

    def __init__(self, key=None):
        key = bytes(key or b"")
        h = vp()
        check(lib().lg_prng_create(key, len(key), C.byref(h)))
        self.h = h
        self._seed = b""

    def __del__(self):
        try:
            lib().lg_prng_destroy(self.h)
        except Exception:
            pass

    def GetClock(self):
        return int(lib().lg_prng_get_clock(self.h))

    def Seed(self, seed):
        self._seed = bytes(seed)
        check(lib().lg_prng_seed(self.h, self._seed, len(self._seed)))

    def GetSeed(self):
        return self._seed

    def Clock(self):
        out = C.create_string_buffer(64)
        check(lib().lg_prng_clock(self.h, out))
        return out.raw

    def SetClock(self, n):
        check(lib().lg_prng_set_clock(self.h, n))


def NewPRNG(key=None):
    return PRNG(key)


class CRPGenerator:
    """ring.CRPGenerator (ring/prng.go:11-103): deterministic uniform polynomials of `context` from the PRNG"""

    def __init__(self, key, context):
        key = bytes(key or b"")
        h = vp()
        check(lib().lg_crp_create(key, len(key), context.h, C.byref(h)))
        self.h, self.context = h, context
        self._seed = b""

    def __del__(self):
        try:
            lib().lg_crp_destroy(self.h)
        except Exception:
            pass

    def GetClock(self):
        return int(lib().lg_crp_get_clock(self.h))

    def Seed(self, seed):
        self._seed = bytes(seed)
        check(lib().lg_crp_seed(self.h, self._seed, len(self._seed)))

    def GetSeed(self):
        return self._seed

    def SetClock(self, n):
        check(lib().lg_crp_set_clock(self.h, n))

    def Clock(self, crp, batch_index=0, stream=None):
        check(lib().lg_crp_clock(self.h, crp.h, batch_index, _s(stream)))

    def ClockNew(self):
        crp = self.context.NewPoly()
        self.Clock(crp)
        return crp


def NewCRPGenerator(key, context):
    return CRPGenerator(key, context)


class SimpleScaler:
    """ring.SimpleScaler (ring/ring_scaling.go:166-300): Scale(p1, p2) writes round(t/Q * p1) mod t to every limb of
    p2; the Float128 accumulation of ring/float128.go runs on the device, one coefficient per thread."""

    def __init__(self, t, context):
        h = vp()
        check(lib().lg_scaler_create(t, context.h, C.byref(h)))
        self.h, self.context, self.t = h, context, int(t)

    def __del__(self):
        try:
            lib().lg_scaler_destroy(self.h)
        except Exception:
            pass

    def params(self):
        wi = np.zeros(self.context.nl, np.uint64)
        ti = np.zeros((self.context.nl, 2), np.float64)
        check(lib().lg_scaler_get_params(self.h, _ptr(wi), ti.ctypes.data_as(C.POINTER(C.c_double))))
        return wi, ti

    def Scale(self, p1, p2, stream=None):
        check(lib().lg_scaler_scale(self.h, p1.h, p2.h, _s(stream)))


def NewSimpleScaler(t, context):
    return SimpleScaler(t, context)


def SimpleScalerParams(t, moduli):
    """NewSimpleScaler's tables for a bare modulus list, computed on the host (no device needed)"""
    q = _arr(list(moduli))
    wi = np.zeros(len(q), np.uint64)
    ti = np.zeros((len(q), 2), np.float64)
    a, m = C.c_uint64(), C.c_uint64()
    check(lib().lg_scaler_params_host(t, _ptr(q), len(q), _ptr(wi), ti.ctypes.data_as(C.POINTER(C.c_double)), C.byref(a), C.byref(m)))
    return wi, ti, int(a.value), int(m.value)
