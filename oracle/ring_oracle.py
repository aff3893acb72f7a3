"""ctypes loader for the CPU oracle (oracle/ring_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never by the product package.
Polynomials are numpy uint64 arrays of shape [nlimbs, N] (C-contiguous).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libring_oracle.so")

u64 = C.c_uint64
p64 = C.POINTER(C.c_uint64)
vp = C.c_void_p


def build(force=False):
    src = os.path.join(_HERE, "ring_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _declare(_lib)
    return _lib


def _declare(L):
    def f(name, res, *args):
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = list(args)

    f("orc_bred_params", None, u64, p64)
    f("orc_mred_params", u64, u64)
    f("orc_mform", u64, u64, u64, p64)
    f("orc_mform_constant", u64, u64, u64, p64)
    f("orc_invmform", u64, u64, u64, u64)
    f("orc_invmform_constant", u64, u64, u64, u64)
    f("orc_mred", u64, u64, u64, u64, u64)
    f("orc_mred_constant", u64, u64, u64, u64, u64)
    f("orc_bred_add", u64, u64, u64, p64)
    f("orc_bred_add_constant", u64, u64, u64, p64)
    f("orc_bred", u64, u64, u64, u64, p64)
    f("orc_bred_constant", u64, u64, u64, u64, p64)
    f("orc_cred", u64, u64, u64)
    f("orc_power_of_2", u64, u64, u64, u64, u64)
    f("orc_modexp", u64, u64, u64, u64)
    f("orc_bitreverse64", u64, u64, u64)
    f("orc_small_prime", u64, C.c_int)
    f("orc_is_prime", C.c_int, u64)
    f("orc_generate_ntt_primes", C.c_int, u64, u64, u64, p64)
    f("orc_primitive_root", u64, u64)
    f("orc_ctx_new", vp, u64, C.c_int, p64)
    f("orc_ctx_free", None, vp)
    f("orc_ctx_scalars", None, vp, p64, p64, p64, p64, p64, p64)
    f("orc_ctx_tables", None, vp, C.c_int, p64, p64)
    f("orc_ctx_rescale_param", u64, vp, C.c_int, C.c_int)
    f("orc_ntt_limb", None, p64, p64, u64, p64, u64, u64, p64)
    f("orc_invntt_limb", None, p64, p64, u64, p64, u64, u64, u64)
    f("orc_ntt", None, vp, C.c_int, p64, p64)
    f("orc_invntt", None, vp, C.c_int, p64, p64)
    f("orc_ntt_one", None, vp, C.c_int, p64, p64)
    f("orc_invntt_one", None, vp, C.c_int, p64, p64)
    for name in ("add", "add_nomod", "sub", "sub_nomod", "mulcoeffs", "mulcoeffs_and_add", "mulcoeffs_and_add_nomod",
                 "mulcoeffs_constant", "mulcoeffs_montgomery", "mulcoeffs_montgomery_and_add",
                 "mulcoeffs_montgomery_and_add_nomod", "mulcoeffs_montgomery_constant_and_add_nomod",
                 "mulcoeffs_montgomery_and_sub", "mulcoeffs_montgomery_and_sub_nomod",
                 "mulcoeffs_montgomery_constant"):
        f("orc_" + name, None, vp, C.c_int, p64, p64, p64)
    for name in ("neg", "reduce", "mform_poly", "invmform_poly", "bitreverse_poly"):
        f("orc_" + name, None, vp, C.c_int, p64, p64)
    f("orc_add_scalar", None, vp, C.c_int, p64, p64)
    f("orc_sub_scalar", None, vp, C.c_int, p64, p64)
    f("orc_mul_scalar", None, vp, C.c_int, p64, p64, p64)
    f("orc_mul_by_pow2", None, vp, C.c_int, p64, u64, p64)
    f("orc_mult_by_monomial", None, vp, C.c_int, p64, u64, p64)
    f("orc_mul_by_vector_montgomery", None, vp, C.c_int, p64, p64, p64)
    f("orc_mul_by_vector_montgomery_and_add_nomod", None, vp, C.c_int, p64, p64, p64)
    f("orc_gen_galois_params", None, u64, u64, p64)
    f("orc_permute_ntt_index", None, u64, u64, u64, p64)
    f("orc_permute_ntt_with_index", None, u64, C.c_int, p64, p64, p64)
    f("orc_permute_ntt", None, u64, C.c_int, p64, u64, p64)
    f("orc_permute", None, vp, C.c_int, p64, u64, p64)
    f("orc_extender_new", vp, vp, vp)
    f("orc_extender_free", None, vp)
    f("orc_extender_params", None, vp, p64, p64)
    f("orc_modup_split_qp", None, vp, C.c_int, p64, p64)
    f("orc_modup_split_pq", None, vp, C.c_int, p64, p64)
    f("orc_moddown_ntt_pq", None, vp, C.c_int, p64, p64)
    f("orc_moddown_splited_ntt_pq", None, vp, C.c_int, p64, p64, p64)
    f("orc_moddown_pq", None, vp, C.c_int, p64, p64)
    f("orc_moddown_splited_pq", None, vp, C.c_int, p64, p64, p64)
    f("orc_moddown_splited_qp", None, vp, C.c_int, C.c_int, p64, p64, p64)
    f("orc_decomposer_new", vp, p64, C.c_int, p64, C.c_int)
    f("orc_decomposer_free", None, vp)
    f("orc_decomposer_beta", C.c_int, vp)
    f("orc_decomposer_xalpha", C.c_int, vp, C.c_int)
    f("orc_decompose_and_split", None, vp, u64, C.c_int, C.c_int, p64, p64, p64)
    f("orc_decompose", None, vp, u64, C.c_int, C.c_int, p64, p64)
    for name in ("div_floor_by_last_modulus_ntt", "div_floor_by_last_modulus", "div_round_by_last_modulus_ntt",
                 "div_round_by_last_modulus"):
        f("orc_" + name, None, vp, C.c_int, p64)
    for name in ("div_floor_by_last_modulus_many", "div_floor_by_last_modulus_many_ntt",
                 "div_round_by_last_modulus_many", "div_round_by_last_modulus_many_ntt"):
        f("orc_" + name, None, vp, C.c_int, p64, C.c_int)
    f("orc_ckks_eval_new", vp, vp, vp)
    f("orc_ckks_eval_free", None, vp)
    f("orc_ckks_switch_keys_in_place", None, vp, C.c_int, p64, p64, p64, p64)
    f("orc_ckks_mul_relin", None, vp, C.c_int, p64, p64, p64, p64)
    f("orc_ckks_rescale", None, vp, C.c_int, p64, C.c_int)
    f("orc_ckks_permute_ntt", None, vp, C.c_int, p64, p64, p64, p64)
    f("orc_ckks_switch_keys", None, vp, C.c_int, p64, p64, p64)
    f("orc_ckks_hoist", None, vp, C.c_int, p64, p64, p64)
    f("orc_ckks_switch_key_hoisted", None, vp, C.c_int, p64, p64, p64, p64, p64, p64)


def ptr(a):
    assert a.dtype == np.uint64 and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(p64)


def arr(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.uint64))


def bred_params(q):
    u = (u64 * 2)()
    lib().orc_bred_params(q, u)
    return [int(u[0]), int(u[1])]


def generate_ntt_primes(logq, logn, levels):
    out = np.zeros(levels, dtype=np.uint64)
    n = lib().orc_generate_ntt_primes(logq, logn, levels, ptr(out))
    assert n == levels
    return [int(x) for x in out]


def gen_moduli(logn, log_qi, log_pi, log_extra=()):
    """ckks/utils.go:150-193 GenModuli (also bfv/utils.go:26-85 with a third
    QiMul list): primes are generated per bit size, then dealt in order to Q,
    then P (then QMul)."""
    need = {}
    for b in list(log_qi) + list(log_pi) + list(log_extra):
        assert b <= 60
        need[b] = need.get(b, 0) + 1
    primes = {b: generate_ntt_primes(b, logn, n) for b, n in need.items()}
    out = []
    for group in (log_qi, log_pi, log_extra):
        lst = []
        for b in group:
            lst.append(primes[b][0])
            primes[b] = primes[b][1:]
        out.append(lst)
    return out


class Context:
    """ring.Context restated (ring/ring_context.go:18-209)."""

    def __init__(self, N, moduli):
        self.N = int(N)
        self.moduli = [int(q) for q in moduli]
        self.nl = len(self.moduli)
        m = arr(self.moduli)
        self.h = lib().orc_ctx_new(self.N, self.nl, ptr(m))
        if not self.h:
            raise ValueError("moduli do not allow NTT / invalid N")
        self.bred = np.zeros((self.nl, 2), dtype=np.uint64)
        self.mred = np.zeros(self.nl, dtype=np.uint64)
        self.ninv = np.zeros(self.nl, dtype=np.uint64)
        self.psi_mont = np.zeros(self.nl, dtype=np.uint64)
        self.psi_inv_mont = np.zeros(self.nl, dtype=np.uint64)
        lib().orc_ctx_scalars(self.h, None, ptr(self.bred), ptr(self.mred), ptr(self.ninv), ptr(self.psi_mont),
                              ptr(self.psi_inv_mont))

    def __del__(self):
        try:
            if self.h:
                lib().orc_ctx_free(self.h)
                self.h = None
        except Exception:
            pass

    def tables(self, limb):
        psi = np.zeros(self.N, dtype=np.uint64)
        psi_inv = np.zeros(self.N, dtype=np.uint64)
        lib().orc_ctx_tables(self.h, limb, ptr(psi), ptr(psi_inv))
        return psi, psi_inv

    def all_tables(self):
        psi = np.zeros((self.nl, self.N), dtype=np.uint64)
        psi_inv = np.zeros((self.nl, self.N), dtype=np.uint64)
        for i in range(self.nl):
            lib().orc_ctx_tables(self.h, i, ptr(psi[i]), ptr(psi_inv[i]))
        return psi, psi_inv

    def rescale_params(self):
        """flat triangular list rescaleParams[j-1][i], i<j (ring_context.go:148-158)"""
        return [[int(lib().orc_ctx_rescale_param(self.h, j, i)) for i in range(j)] for j in range(1, self.nl)]

    def new_poly(self, nl=None):
        return np.zeros((self.nl if nl is None else nl, self.N), dtype=np.uint64)

    # --- generic dispatch helpers -------------------------------------------------
    def _nl(self, a, nl):
        return a.shape[0] if nl is None else nl

    def op3(self, name, p1, p2, p3=None, nl=None):
        nl = self._nl(p1, nl)
        if p3 is None:
            p3 = np.zeros((nl, self.N), dtype=np.uint64)
        getattr(lib(), "orc_" + name)(self.h, nl, ptr(p1), ptr(p2), ptr(p3))
        return p3

    def op2(self, name, p1, p2=None, nl=None):
        nl = self._nl(p1, nl)
        if p2 is None:
            p2 = np.zeros((nl, self.N), dtype=np.uint64)
        getattr(lib(), "orc_" + name)(self.h, nl, ptr(p1), ptr(p2))
        return p2

    def ntt(self, p, out=None, nl=None):
        nl = self._nl(p, nl)
        if out is None:
            out = np.zeros((nl, self.N), dtype=np.uint64)
        lib().orc_ntt(self.h, nl, ptr(p), ptr(out))
        return out

    def invntt(self, p, out=None, nl=None):
        nl = self._nl(p, nl)
        if out is None:
            out = np.zeros((nl, self.N), dtype=np.uint64)
        lib().orc_invntt(self.h, nl, ptr(p), ptr(out))
        return out

    def permute(self, p, gen, nl=None):
        nl = self._nl(p, nl)
        out = np.zeros((nl, self.N), dtype=np.uint64)
        lib().orc_permute(self.h, nl, ptr(p), gen, ptr(out))
        return out

    def mul_scalar(self, p, scalars, nl=None):
        nl = self._nl(p, nl)
        out = np.zeros((nl, self.N), dtype=np.uint64)
        s = arr(scalars)
        lib().orc_mul_scalar(self.h, nl, ptr(p), ptr(s), ptr(out))
        return out

    def div_round_ntt(self, p):
        """DivRoundByLastModulusNTT; returns the nl-1 limb result (input not modified)."""
        q = p.copy()
        lib().orc_div_round_by_last_modulus_ntt(self.h, q.shape[0], ptr(q))
        return q[:-1].copy()


def _mred1(x, y, q, qinv):
    """MRed, modular_reduction.go:70-79, on Python integers"""
    M = (1 << 64) - 1
    t = x * y
    h = ((((t & M) * qinv) & M) * q) >> 64
    r = ((t >> 64) - h + q) & M
    return r - q if r >= q else r


def _go_mask(N):
    """(1 << N) - 1 on uint64 as Go evaluates it: a shift by 64 or more gives 0"""
    return (1 << 64) - 1 if N >= 64 else (1 << N) - 1


def mul_poly(ctx, p1, p2, montgomery=False):
    """MulPoly ring/ring.go:358-367, MulPolyMontgomery :371-380"""
    a, b = ctx.ntt(p1), ctx.ntt(p2)
    p3 = ctx.op3("mulcoeffs_montgomery" if montgomery else "mulcoeffs", a, b)
    return ctx.invntt(p3)


def mul_poly_naive(ctx, p1, p2, montgomery=False):
    """MulPolyNaive ring/ring.go:383-410 (p1 to Montgomery form first), MulPolyNaiveMontgomery :413-437; small N only"""
    N = ctx.N
    c1 = p1.copy() if montgomery else ctx.op2("mform_poly", p1)
    out = np.zeros_like(p1)
    for x, q in enumerate(ctx.moduli):
        qinv = int(ctx.mred[x])
        a, b = c1[x].tolist(), p2[x].tolist()
        r = [0] * N
        for i in range(N):
            for j in range(i):
                v = r[j] + (q - _mred1(a[i], b[N - i + j], q, qinv))
                r[j] = v - q if v >= q else v
            for j in range(i, N):
                v = r[j] + _mred1(a[i], b[j - i], q, qinv)
                r[j] = v - q if v >= q else v
        out[x] = np.array(r, dtype=np.uint64)
    return out


def ring_exp(ctx, p1, e):
    """Exp ring/ring.go:441-464; returns (p1 after the call, p2)"""
    p1 = ctx.ntt(p1)
    tmp = ctx.op3("add", ctx.new_poly(), p1)
    p2 = np.ones_like(p1)
    i = int(e)
    while i > 0:
        if i & 1:
            p2 = ctx.op3("mulcoeffs", p2, tmp)
        tmp = ctx.op3("mulcoeffs", tmp, p1)
        i >>= 1
    p2 = ctx.invntt(p2)
    p2 = ctx.invntt(p1)  # :463 overwrites the power
    return p1, p2


def ring_shift(ctx, p1, n):
    """Shift ring/ring.go:575-580: append(p1[n:], p1[:n]...)"""
    k = int(n) & _go_mask(ctx.N)
    if k > ctx.N:
        raise IndexError("slice bounds out of range")
    return np.concatenate([p1[:, k:], p1[:, :k]], axis=1)


def ring_rotate(ctx, p1, n):
    """Rotate ring/ring.go:775-800: written into p1 (p2 is never touched); returns p1 after the call; small N only"""
    n = int(n) & _go_mask(ctx.N)
    out = p1.copy()
    for i, q in enumerate(ctx.moduli):
        qinv = int(ctx.mred[i])
        psi = int(ctx.psi_mont[i])
        root = _mred1(psi, psi, q, qinv)
        res, x, e = (1 << 64) % q, root, n  # modexpMontgomery, ring/utils.go:39-50
        while e > 0:
            if e & 1:
                res = _mred1(res, x, q, qinv)
            x = _mred1(x, x, q, qinv)
            e >>= 1
        root = res
        gal = (1 << 64) % q
        row = out[i].tolist()
        for j in range(1, ctx.N):
            gal = _mred1(gal, root, q, qinv)
            row[j] = _mred1(row[j], gal, q, qinv)
        out[i] = np.array(row, dtype=np.uint64)
    return out


def ring_equal(ctx, p1, p2, level=None):
    """Equal / EqualLvl ring/ring_context.go:424-467: returns (equal, p1 reduced, p2 reduced)"""
    nl = ctx.nl if level is None else level + 1
    a, b = p1.copy(), p2.copy()
    a[:nl] = ctx.op2("reduce", p1[:nl].copy(), nl=nl)
    b[:nl] = ctx.op2("reduce", p2[:nl].copy(), nl=nl)
    return bool(np.array_equal(a[:nl], b[:nl])), a, b


def permute_ntt_index(gen, power, N):
    idx = np.zeros(N, dtype=np.uint64)
    lib().orc_permute_ntt_index(gen, power, N, ptr(idx))
    return idx


def permute_ntt_with_index(p, index):
    out = np.zeros_like(p)
    lib().orc_permute_ntt_with_index(p.shape[1], p.shape[0], ptr(p), ptr(index), ptr(out))
    return out


class Extender:
    """ring.FastBasisExtender restated (ring/ring_basis_extension.go:9-350)."""

    def __init__(self, ctxQ, ctxP):
        self.Q, self.P = ctxQ, ctxP
        self.h = lib().orc_extender_new(ctxQ.h, ctxP.h)

    def __del__(self):
        try:
            lib().orc_extender_free(self.h)
        except Exception:
            pass

    def modup_split_qp(self, level, p1):
        out = self.P.new_poly()
        lib().orc_modup_split_qp(self.h, level, ptr(p1), ptr(out))
        return out

    def modup_split_pq(self, level, p1):
        out = self.Q.new_poly()
        lib().orc_modup_split_pq(self.h, level, ptr(p1), ptr(out))
        return out

    def moddown_ntt_pq(self, level, p1):
        p1 = p1.copy()
        out = self.Q.new_poly(level + 1)
        lib().orc_moddown_ntt_pq(self.h, level, ptr(p1), ptr(out))
        return out

    def moddown_splited_ntt_pq(self, level, p1Q, p1P):
        p1P = p1P.copy()
        out = self.Q.new_poly(level + 1)
        lib().orc_moddown_splited_ntt_pq(self.h, level, ptr(p1Q), ptr(p1P), ptr(out))
        return out

    def moddown_pq(self, level, p1):
        out = self.Q.new_poly(level + 1)
        lib().orc_moddown_pq(self.h, level, ptr(p1), ptr(out))
        return out

    def moddown_splited_pq(self, level, p1Q, p1P):
        out = self.Q.new_poly(level + 1)
        lib().orc_moddown_splited_pq(self.h, level, ptr(p1Q), ptr(p1P), ptr(out))
        return out

    def moddown_splited_qp(self, levelQ, levelP, p1Q, p1P):
        out = self.P.new_poly(levelP + 1)
        lib().orc_moddown_splited_qp(self.h, levelQ, levelP, ptr(p1Q), ptr(p1P), ptr(out))
        return out


class Decomposer:
    """ring.Decomposer restated (ring/ring_basis_extension.go:398-713)."""

    def __init__(self, Q, P, N):
        self.Qm, self.Pm, self.N = list(Q), list(P), N
        q, p = arr(Q), arr(P)
        self.h = lib().orc_decomposer_new(ptr(q), len(Q), ptr(p), len(P))
        self.beta = lib().orc_decomposer_beta(self.h)

    def __del__(self):
        try:
            lib().orc_decomposer_free(self.h)
        except Exception:
            pass

    def decompose_and_split(self, level, crt, p0):
        outQ = np.zeros((level + 1, self.N), dtype=np.uint64)
        outP = np.zeros((len(self.Pm), self.N), dtype=np.uint64)
        lib().orc_decompose_and_split(self.h, self.N, level, crt, ptr(p0), ptr(outQ), ptr(outP))
        return outQ, outP

    def decompose(self, level, crt, p0):
        out = np.zeros((level + 1 + len(self.Pm), self.N), dtype=np.uint64)
        lib().orc_decompose(self.h, self.N, level, crt, ptr(p0), ptr(out))
        return out


class CkksEvaluator:
    """Hot ops of ckks.evaluator restated (ckks/evaluator.go:933-1591)."""

    def __init__(self, ctxQ, ctxP):
        self.Q, self.P = ctxQ, ctxP
        self.h = lib().orc_ckks_eval_new(ctxQ.h, ctxP.h)

    def __del__(self):
        try:
            lib().orc_ckks_eval_free(self.h)
        except Exception:
            pass

    def switch_keys_in_place(self, level, cx, evk):
        p0 = np.zeros((level + 1, self.Q.N), dtype=np.uint64)
        p1 = np.zeros((level + 1, self.Q.N), dtype=np.uint64)
        lib().orc_ckks_switch_keys_in_place(self.h, level, ptr(cx), ptr(evk), ptr(p0), ptr(p1))
        return p0, p1

    def mul_relin(self, level, ct0, ct1, evk):
        out = np.zeros((2, level + 1, self.Q.N), dtype=np.uint64)
        lib().orc_ckks_mul_relin(self.h, level, ptr(ct0), ptr(ct1), ptr(evk), ptr(out))
        return out

    def rescale(self, ct, nb=1):
        nl = ct.shape[1]
        buf = ct.copy()
        lib().orc_ckks_rescale(self.h, nl, ptr(buf), nb)
        return buf.reshape(-1)[: 2 * (nl - nb) * self.Q.N].reshape(2, nl - nb, self.Q.N).copy()

    def permute_ntt(self, level, ct, index, evk):
        out = np.zeros((2, level + 1, self.Q.N), dtype=np.uint64)
        lib().orc_ckks_permute_ntt(self.h, level, ptr(ct), ptr(index), ptr(evk), ptr(out))
        return out

    def switch_keys(self, level, ct, evk):
        out = np.zeros((2, level + 1, self.Q.N), dtype=np.uint64)
        lib().orc_ckks_switch_keys(self.h, level, ptr(ct), ptr(evk), ptr(out))
        return out

    def rotate_columns_pow2(self, level, ct, k, indexes, evks):
        """rotateColumnsPow2 (ckks/evaluator.go:1402-1424): indexes / evks = maps of the power-of-two rotations"""
        out = np.ascontiguousarray(ct[:, : level + 1]).copy()
        i = 1
        while k > 0:
            if k & 1:
                out = self.permute_ntt(level, out, indexes[i], evks[i])
            i <<= 1
            k >>= 1
        return out

    def rotate_columns(self, level, ct, k, left, right):
        """RotateColumns (ckks/evaluator.go:1201-1248); left / right = {k: (index, evk)} as RotationKeys holds them"""
        N = self.Q.N
        k &= (N >> 1) - 1
        if k == 0:
            return np.ascontiguousarray(ct[:, : level + 1]).copy()
        if k in left:
            return self.permute_ntt(level, ct, left[k][0], left[k][1])
        i = 1
        while i < (N >> 1):
            if i not in left or i not in right:
                raise ValueError("cannot RotateColumns: specific rotation and pow2 rotations have not been generated")
            i <<= 1
        if bin(k).count("1") <= bin((N >> 1) - k).count("1"):
            return self.rotate_columns_pow2(level, ct, k, {j: v[0] for j, v in left.items()}, {j: v[1] for j, v in left.items()})
        return self.rotate_columns_pow2(level, ct, (N >> 1) - k, {j: v[0] for j, v in right.items()},
                                        {j: v[1] for j, v in right.items()})

    def rescale_many(self, ct, nb):
        """RescaleMany (ckks/evaluator.go:971-1000): DivRoundByLastModulusManyNTT (ring_scaling.go:152-156) per value"""
        nl = ct.shape[1]
        outs = []
        for v in ct:
            buf = np.ascontiguousarray(v).copy()
            lib().orc_div_round_by_last_modulus_many_ntt(self.Q.h, nl, ptr(buf), nb)
            outs.append(buf[: nl - nb].copy())
        return np.stack(outs)

    def rotate_hoisted(self, level, ct, indexes, evks):
        """RotateHoisted (ckks/evaluator.go:1252-1289): one decomposition of ct.value[1] shared by every
        rotation; indexes[k] = permuteNTTLeftIndex, evks[k] = evakeyRotColLeft of rotation k."""
        nl = level + 1
        beta = -(-nl // self.P.nl)
        qdec = np.zeros((beta, self.Q.nl, self.Q.N), dtype=np.uint64)
        pdec = np.zeros((beta, self.P.nl, self.Q.N), dtype=np.uint64)
        lib().orc_ckks_hoist(self.h, level, ptr(ct), ptr(qdec), ptr(pdec))
        outs = []
        for index, evk in zip(indexes, evks):
            out = np.zeros((2, nl, self.Q.N), dtype=np.uint64)
            lib().orc_ckks_switch_key_hoisted(self.h, level, ptr(ct), ptr(qdec), ptr(pdec), ptr(index), ptr(evk), ptr(out))
            outs.append(out)
        return outs


# ---------------------------------------------------------------------------
# CKKS key generator / encryptor / decryptor ring sequences (ckks/keygen.go, ckks/encryptor.go,
# ckks/decryptor.go), restated as compositions of the oracle's ring ops.  The sampled values are inputs
# (the reference draws them from crypto/rand): ternary / gaussian coefficients as signed ints.
# ---------------------------------------------------------------------------
def signed_residues(moduli, coeffs):
    c = np.asarray(coeffs, dtype=np.int64)
    return np.ascontiguousarray(np.stack([np.where(c < 0, np.int64(q) + c, c).astype(np.uint64) for q in moduli]))


class CkksScheme:
    """The non-fast encrypt paths call ModDownPQ(level, pool) with the full QP pool, literally as
    encryptor.go:223,:348 do (below the top level its "P part" Coeffs[level+1:...] are Q limbs)."""

    def __init__(self, Q, P, N):
        self.Qm, self.Pm, self.N = list(Q), list(P), N
        self.Q, self.P, self.QP = Context(N, self.Qm), Context(N, self.Pm), Context(N, self.Qm + self.Pm)
        self.ext = Extender(self.Q, self.P)
        self.levels, self.alpha = len(self.Qm), len(self.Pm)
        self.beta = -(-self.levels // self.alpha)
        self.Pbig = 1
        for p in self.Pm:
            self.Pbig *= int(p)

    def gen_secret_key(self, ternary):  # keygen.go:96-112
        return self.QP.ntt(self.QP.op2("mform_poly", signed_residues(self.Qm + self.Pm, ternary)))

    def gen_public_key(self, sk, e, a):  # keygen.go:138-151
        pk0 = self.QP.ntt(signed_residues(self.Qm + self.Pm, e))
        self.QP.op3("mulcoeffs_montgomery_and_add", sk, np.ascontiguousarray(a), pk0)
        return self.QP.op2("neg", pk0), np.ascontiguousarray(a).copy()

    def new_switching_key(self, sk_in, sk_out, errors, uniforms):  # keygen.go:282-340
        QP = self.Qm + self.Pm
        sk_in = self.QP.mul_scalar(sk_in, [self.Pbig % q for q in QP])
        evk = np.zeros((self.beta, 2, len(QP), self.N), dtype=np.uint64)
        for i in range(self.beta):
            k0 = self.QP.op2("mform_poly", self.QP.ntt(signed_residues(QP, errors[i])))
            k1 = np.ascontiguousarray(uniforms[i]).copy()
            for j in range(self.alpha):
                index = i * self.alpha + j
                qi = np.uint64(QP[index])
                t = k0[index] + sk_in[index]
                k0[index] = np.where(t >= qi, t - qi, t)  # CRed :325
                if index >= self.levels - 1:
                    break
            self.QP.op3("mulcoeffs_montgomery_and_sub", k1, sk_out, k0)
            evk[i, 0], evk[i, 1] = k0, k1
        return evk

    def gen_relin_key(self, sk, errors, uniforms):  # keygen.go:190-203
        return self.new_switching_key(self.QP.op3("mulcoeffs_montgomery", sk, sk), sk, errors, uniforms)

    def gen_rot_key(self, sk, gen, errors, uniforms):  # keygen.go:487-494
        idx = permute_ntt_index(gen, 1, self.N)
        return self.new_switching_key(permute_ntt_with_index(sk, idx), sk, errors, uniforms)

    def encrypt_pk(self, level, pt, pk, u, e0, e1, fast=False):  # encryptor.go:179-237
        nl = level + 1
        if fast:
            up = self.Q.ntt(self.Q.op2("mform_poly", signed_residues(self.Qm, u)))
            c0 = self.Q.op3("mulcoeffs_montgomery", up, np.ascontiguousarray(pk[0][: self.levels]))
            c1 = self.Q.op3("mulcoeffs_montgomery", up, np.ascontiguousarray(pk[1][: self.levels]))
            c0 = self.Q.op3("add", c0, self.Q.ntt(signed_residues(self.Qm, e0)))
            c1 = self.Q.op3("add", c1, self.Q.ntt(signed_residues(self.Qm, e1)))
            c0, c1 = c0[:nl].copy(), c1[:nl].copy()
        else:
            QP = self.Qm + self.Pm
            up = self.QP.ntt(self.QP.op2("mform_poly", signed_residues(QP, u)))
            p0 = self.QP.invntt(self.QP.op3("mulcoeffs_montgomery", up, pk[0]))
            p1 = self.QP.invntt(self.QP.op3("mulcoeffs_montgomery", up, pk[1]))
            p0 = self.QP.op3("add", p0, signed_residues(QP, e0))
            p1 = self.QP.op3("add", p1, signed_residues(QP, e1))
            c0 = self.Q.ntt(self.ext.moddown_pq(level, p0)[:nl].copy(), nl=nl)
            c1 = self.Q.ntt(self.ext.moddown_pq(level, p1)[:nl].copy(), nl=nl)
        return np.stack([self.Q.op3("add", c0, np.ascontiguousarray(pt[:nl]), nl=nl), c1])

    def encrypt_sk(self, level, pt, sk, crp, e, fast=False):  # encryptor.go:318-362
        nl = level + 1
        if fast:
            c0 = self.Q.op2("neg", self.Q.op3("mulcoeffs_montgomery", np.ascontiguousarray(crp), np.ascontiguousarray(sk[: self.levels])))
            c0 = self.Q.op3("add", c0, self.Q.ntt(signed_residues(self.Qm, e)))[:nl].copy()
            c1 = np.ascontiguousarray(crp[:nl]).copy()
        else:
            QP = self.Qm + self.Pm
            p0 = self.QP.invntt(self.QP.op2("neg", self.QP.op3("mulcoeffs_montgomery", np.ascontiguousarray(crp), sk)))
            p0 = self.QP.op3("add", p0, signed_residues(QP, e))
            c0 = self.Q.ntt(self.ext.moddown_pq(level, p0)[:nl].copy(), nl=nl)
            c1 = self.ext.moddown_ntt_pq(level, np.ascontiguousarray(crp).copy())[:nl].copy()
        return np.stack([self.Q.op3("add", c0, np.ascontiguousarray(pt[:nl]), nl=nl), c1])

    def decrypt(self, level, ct, sk):  # decryptor.go:53-78
        nl = level + 1
        skq = np.ascontiguousarray(sk[:nl])
        degree = len(ct) - 1
        pt = np.ascontiguousarray(ct[degree][:nl]).copy()
        for i in range(degree, 0, -1):
            pt = self.Q.op3("mulcoeffs_montgomery", pt, skq, nl=nl)
            pt = self.Q.op3("add", pt, np.ascontiguousarray(ct[i - 1][:nl]), nl=nl)
            if i & 7 == 7:
                pt = self.Q.op2("reduce", pt, nl=nl)
        if degree & 7 != 7:
            pt = self.Q.op2("reduce", pt, nl=nl)
        return pt


# ---------------------------------------------------------------------------
# dckks protocols CKS / RTG / RKG (dckks/keyswitching.go, rotkey_gen.go, relinkey_gen.go) restated as
# compositions of the oracle's ring ops; sampled values are inputs (signed coefficient vectors)
# ---------------------------------------------------------------------------
class DckksProtocols:
    def __init__(self, scheme):
        self.S = scheme
        self.K = scheme.QP
        self.mods = scheme.Qm + scheme.Pm

    def _add_digit(self, dst, src, i):
        S = self.S
        for j in range(S.alpha):
            index = i * S.alpha + j
            qi = np.uint64(self.mods[index])
            t = dst[index] + src[index]
            dst[index] = np.where(t >= qi, t - qi, t)
            if index >= S.levels - 1:
                break

    def cks_gen_share(self, level, sk_in, sk_out, ct1, e):  # keyswitching.go:62-96
        S = self.S
        nl, nQ = level + 1, S.levels
        delta = S.Q.op3("sub", np.ascontiguousarray(sk_in[:nQ]), np.ascontiguousarray(sk_out[:nQ]))
        share = S.Q.op3("mulcoeffs_montgomery", np.ascontiguousarray(ct1[:nl]), np.ascontiguousarray(delta[:nl]), nl=nl)
        share = S.Q.mul_scalar(share, [S.Pbig % q for q in S.Qm[:nl]], nl=nl)
        tmp = self.K.ntt(signed_residues(self.mods, e))
        share = S.Q.op3("add", share, np.ascontiguousarray(tmp[:nl]), nl=nl)
        return S.ext.moddown_splited_ntt_pq(level, share, np.ascontiguousarray(tmp[nQ:]))

    def bfv_cks_gen_share(self, sk_in, sk_out, ct1, e):  # dbfv/keyswitching.go:73-106 (coefficient-domain ciphertext)
        S = self.S
        nQ = S.levels
        delta = S.Q.op3("sub", np.ascontiguousarray(sk_in[:nQ]), np.ascontiguousarray(sk_out[:nQ]))
        share = S.Q.op3("mulcoeffs_montgomery", S.Q.ntt(np.ascontiguousarray(ct1)), delta)
        share = S.Q.invntt(S.Q.mul_scalar(share, [S.Pbig % q for q in S.Qm]))
        tmp = signed_residues(self.mods, e)
        share = S.Q.op3("add", share, np.ascontiguousarray(tmp[:nQ]))
        return S.ext.moddown_splited_pq(nQ - 1, share, np.ascontiguousarray(tmp[nQ:]))

    def rtg_gen_share(self, sk, gal_el, crp, errors):  # rotkey_gen.go:95-141
        S = self.S
        idx = permute_ntt_index(gal_el, 1, S.N)
        tmp = permute_ntt_with_index(sk, idx)
        tmp = self.K.op2("invmform_poly", self.K.mul_scalar(tmp, [S.Pbig % q for q in self.mods]))
        out = []
        for i in range(S.beta):
            ek = self.K.ntt(signed_residues(self.mods, errors[i]))
            self._add_digit(ek, tmp, i)
            self.K.op3("mulcoeffs_montgomery_and_sub", np.ascontiguousarray(crp[i]), sk, ek)
            out.append(self.K.op2("mform_poly", ek))
        return out

    def rtg_finalize(self, share, crp):  # rotkey_gen.go:164-174
        return np.ascontiguousarray(np.stack([np.stack([share[i], self.K.op2("mform_poly", np.ascontiguousarray(crp[i]))])
                                              for i in range(self.S.beta)]))

    def rkg_round1(self, u, sk, crp, errors):  # relinkey_gen.go:65-112
        S = self.S
        pool = self.K.op2("invmform_poly", self.K.mul_scalar(sk, [S.Pbig % q for q in self.mods]))
        out = []
        for i in range(S.beta):
            h = self.K.ntt(signed_residues(self.mods, errors[i]))
            self._add_digit(h, pool, i)
            self.K.op3("mulcoeffs_montgomery_and_sub", u, np.ascontiguousarray(crp[i]), h)
            out.append(h)
        return out

    def rkg_round2(self, round1, sk, crp, errors1, errors2):  # :135-163
        out = []
        for i in range(self.S.beta):
            s0 = self.K.op3("mulcoeffs_montgomery", round1[i], sk)
            s0 = self.K.op3("add", s0, self.K.ntt(signed_residues(self.mods, errors1[i])))
            s1 = self.K.ntt(signed_residues(self.mods, errors2[i]))
            self.K.op3("mulcoeffs_montgomery_and_add", sk, np.ascontiguousarray(crp[i]), s1)
            out.append((s0, s1))
        return out

    def rkg_round3(self, round2, u, sk, errors):  # :186-199
        pool = self.K.op3("sub", u, sk)
        out = []
        for i in range(self.S.beta):
            h = self.K.ntt(signed_residues(self.mods, errors[i]))
            self.K.op3("mulcoeffs_montgomery_and_add", pool, round2[i][1], h)
            out.append(h)
        return out

    def rkg_key(self, round2, round3):  # :210-223
        return np.ascontiguousarray(np.stack([np.stack([
            self.K.op2("mform_poly", self.K.op3("add", round2[i][0], round3[i])),
            self.K.op2("mform_poly", round2[i][1])]) for i in range(self.S.beta)]))

    # relinkey_gen_naive.go (dckks :53-200; dbfv identical except where round one's second sample lands, :76)
    def rkg_naive_round1(self, sk, pk, errors, us, second_error_into=0, share=None):
        S = self.S
        pool = self.K.op2("invmform_poly", self.K.mul_scalar(sk, [S.Pbig % q for q in self.mods]))
        zero = lambda: np.zeros((len(self.mods), S.N), dtype=np.uint64)
        out = [[zero(), zero()] for _ in range(S.beta)] if share is None else [[a.copy(), b.copy()] for a, b in share]
        for i in range(S.beta):
            out[i][0] = self.K.ntt(signed_residues(self.mods, errors[i][0]))
            out[i][second_error_into] = self.K.ntt(signed_residues(self.mods, errors[i][1]))
            self._add_digit(out[i][0], pool, i)
        for i in range(S.beta):
            u = self.K.ntt(np.ascontiguousarray(us[i]))
            self.K.op3("mulcoeffs_montgomery_and_add", pk[0], u, out[i][0])
            self.K.op3("mulcoeffs_montgomery_and_add", pk[1], u, out[i][1])
        return [(a, b) for a, b in out]

    def rkg_naive_round2(self, round1, sk, pk, us, errors):
        out = []
        for i in range(self.S.beta):
            s0 = self.K.op3("mulcoeffs_montgomery", round1[i][0], sk)
            s1 = self.K.op3("mulcoeffs_montgomery", round1[i][1], sk)
            u = self.K.ntt(np.ascontiguousarray(us[i]))
            self.K.op3("mulcoeffs_montgomery_and_add", pk[0], u, s0)
            self.K.op3("mulcoeffs_montgomery_and_add", pk[1], u, s1)
            s0 = self.K.op3("add", s0, self.K.ntt(signed_residues(self.mods, errors[i][0])))
            s1 = self.K.op3("add", s1, self.K.ntt(signed_residues(self.mods, errors[i][1])))
            out.append((s0, s1))
        return out

    def rkg_naive_key(self, round2):
        return np.ascontiguousarray(np.stack([np.stack([self.K.op2("mform_poly", a), self.K.op2("mform_poly", b)])
                                              for a, b in round2]))

    def add_lists(self, a, b):
        return [self.K.op3("add", x, y) for x, y in zip(a, b)]

    def add_pairs(self, a, b):
        return [(self.K.op3("add", x[0], y[0]), self.K.op3("add", x[1], y[1])) for x, y in zip(a, b)]


# ---------------------------------------------------------------------------
# CKKS constant ops, ckks/evaluator.go:373-833, restated whole (host scalars + coefficient loops)
# ---------------------------------------------------------------------------
def scale_up_exact(value, n, q):  # ckks/utils.go:22-49 (big.Float at 53 bits: x + 0.5 rounds like float64)
    x = -n * value if value < 0 else n * value
    res = int(x + 0.5) % q
    return q - res if value < 0 else res


def _mred_vec(a, s, q, qinv):
    """MRed(a[j], s) for a uint64 vector a (modular_reduction.go:70-79), exact via Python integers"""
    M = (1 << 64) - 1
    out = np.empty(a.shape, dtype=np.uint64)
    for j, x in enumerate(a.tolist()):
        t = x * s
        h = (((t & M) * qinv) & M) * q >> 64
        r = ((t >> 64) - h + q) & M
        out[j] = r - q if r >= q else r
    return out


def _cred_vec(v, q):
    return np.where(v >= np.uint64(q), v - np.uint64(q), v)


def ckks_const_op(ctx, op, level, polys_in, polys_out, c_real=0.0, c_imag=0.0, scale=1.0):
    """op in {"add", "mul", "mul_add", "mul_i", "div_i"}: AddConst :373-448 (value[0] only: pass one poly),
    MultByConst :622-730, MultByConstAndAdd inner loops :560-609, MultByi :746-784, DivByi :795-833.
    polys_*: lists of [nl][N] arrays (NTT domain); returns the new polys_out."""
    N = ctx.N
    h = N >> 1
    outs = [p.copy() for p in polys_out]
    for i in range(level + 1):
        q, qinv = int(ctx.moduli[i]), int(ctx.mred[i])
        bred = np.ascontiguousarray(ctx.bred[i])
        psi2 = int(ctx.tables(i)[0][1])
        if op in ("mul_i", "div_i"):
            first, second = (psi2, q - psi2) if op == "mul_i" else (q - psi2, psi2)
        else:
            re = im = sc = 0
            if c_real != 0:
                re = scale_up_exact(c_real, scale, q)
                sc = re
            if c_imag != 0:
                im = int(lib().orc_mred(scale_up_exact(c_imag, scale, q), psi2, q, qinv))
                sc = int(lib().orc_cred(sc + im, q))
            first = sc if op == "add" else int(lib().orc_mform(sc, q, ptr(bred)))
            second = first
            if c_imag != 0:
                t = int(lib().orc_cred(re + (q - im), q))
                second = t if op == "add" else int(lib().orc_mform(t, q, ptr(bred)))
        for u, pin in enumerate(polys_in):
            for sl, c in ((slice(0, h), first), (slice(h, N), second)):
                a = pin[i, sl]
                if op == "add":
                    outs[u][i, sl] = _cred_vec(a + np.uint64(c), q)
                elif op == "mul_add":
                    outs[u][i, sl] = _cred_vec(outs[u][i, sl] + _mred_vec(a, c, q, qinv), q)
                else:
                    outs[u][i, sl] = _mred_vec(a, c, q, qinv)
    return outs


# ---------------------------------------------------------------------------
# ring.Poly wire format (ring/ring_object.go:146-289)
# ---------------------------------------------------------------------------
def poly_marshal(p, with_metadata=True):
    """WriteTo :161-175: data[0] = log2(N), data[1] = #moduli, then binary.BigEndian.PutUint64 limb-major"""
    nl, N = p.shape
    head = bytes([N.bit_length() - 1, nl]) if with_metadata else b""
    return head + np.ascontiguousarray(p).astype(">u8").tobytes()


def poly_unmarshal(data):
    """UnmarshalBinary :257-274"""
    N, nl = 1 << data[0], data[1]
    if ((len(data) - 2) >> 3) != N * nl:
        raise ValueError("error : invalid polynomial encoding")
    return np.frombuffer(data, dtype=">u8", offset=2).astype(np.uint64).reshape(nl, N)


def _declare_bfv(L):
    def f(name, res, *args):
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = list(args)

    f("orc_bfv_eval_new", vp, vp, vp, vp, vp, u64, p64, p64)
    f("orc_bfv_eval_free", None, vp)
    f("orc_bfv_tensor_and_rescale", None, vp, p64, p64, p64)
    f("orc_bfv_switch_keys_core", None, vp, p64, p64, p64, p64)
    f("orc_bfv_relinearize", None, vp, p64, p64, p64)
    f("orc_bfv_switch_keys", None, vp, p64, p64, p64)
    f("orc_bfv_permute", None, vp, p64, u64, p64, p64)


class BfvEvaluator:
    """Hot ops of bfv.evaluator restated (bfv/evaluator.go:278-813)."""

    def __init__(self, ctxQ, ctxQMul, ctxP, t):
        L = lib()
        if not getattr(L, "_bfv_declared", False):
            _declare_bfv(L)
            L._bfv_declared = True
        self.Q, self.QMul, self.P = ctxQ, ctxQMul, ctxP
        self.QP = Context(ctxQ.N, ctxQ.moduli + ctxP.moduli)
        self.t = int(t)
        prod = 1
        for q in ctxQMul.moduli:
            prod *= q
        phalf = prod >> 1  # bfv/evaluator.go:98
        pm, pq = arr([phalf % q for q in ctxQMul.moduli]), arr([phalf % q for q in ctxQ.moduli])
        self.h = L.orc_bfv_eval_new(ctxQ.h, ctxQMul.h, ctxP.h, self.QP.h, self.t, ptr(pm), ptr(pq))
        self.nQ, self.nP, self.N = ctxQ.nl, ctxP.nl, ctxQ.N

    def __del__(self):
        try:
            lib().orc_bfv_eval_free(self.h)
        except Exception:
            pass

    def tensor_and_rescale(self, ct0, ct1):
        out = np.zeros((3, self.nQ, self.N), dtype=np.uint64)
        lib().orc_bfv_tensor_and_rescale(self.h, ptr(ct0), ptr(ct1), ptr(out))
        return out

    def switch_keys_core(self, cx, evk):
        p0 = np.zeros((self.nQ + self.nP, self.N), dtype=np.uint64)
        p1 = np.zeros((self.nQ + self.nP, self.N), dtype=np.uint64)
        lib().orc_bfv_switch_keys_core(self.h, ptr(cx), ptr(evk), ptr(p0), ptr(p1))
        return p0[: self.nQ].copy(), p1[: self.nQ].copy()

    def relinearize(self, ct, evk):
        out = np.zeros((2, self.nQ, self.N), dtype=np.uint64)
        lib().orc_bfv_relinearize(self.h, ptr(ct), ptr(evk), ptr(out))
        return out

    def switch_keys(self, ct, evk):
        out = np.zeros((2, self.nQ, self.N), dtype=np.uint64)
        lib().orc_bfv_switch_keys(self.h, ptr(ct), ptr(evk), ptr(out))
        return out

    def permute(self, ct, gen, evk):
        out = np.zeros((2, self.nQ, self.N), dtype=np.uint64)
        lib().orc_bfv_permute(self.h, ptr(ct), gen, ptr(evk), ptr(out))
        return out

    # ---- general-degree forms, restated over the oracle's ring ops (bfv/evaluator.go:278-464, :480-507, :578-690) ----
    def tensor_and_rescale_general(self, ct0, ct1, square=False):
        """tensorAndRescale for operands of any degree: ct = [deg+1][nQ][N] (a plaintext has one poly); the branch the
        reference takes when NOT both operands have degree 1 (:374-417), squaring when the operands are the same object"""
        Q, M = self.Q, self.QMul
        if not hasattr(self, "_q1q2"):
            self._q1q2 = Extender(Q, M)
        bc = self._q1q2
        levelQ, levelM = Q.nl - 1, M.nl - 1
        n0, n1 = ct0.shape[0], ct1.shape[0]
        nout = n0 + n1 - 1

        def extend(ct):  # :299-312
            q1 = [Q.ntt(np.ascontiguousarray(v)) for v in ct]
            q2 = [M.ntt(bc.modup_split_qp(levelQ, np.ascontiguousarray(v))) for v in ct]
            return q1, q2

        c0Q1, c0Q2 = extend(ct0)
        c1Q1, c1Q2 = (c0Q1, c0Q2) if square else extend(ct1)
        c2Q1 = [Q.new_poly() for _ in range(nout)]
        c2Q2 = [M.new_poly() for _ in range(nout)]
        if square:  # :382-404
            m1 = [Q.op2("mform_poly", x) for x in c0Q1]
            m2 = [M.op2("mform_poly", x) for x in c0Q2]
            for i in range(n0):
                for j in range(i + 1, n0):
                    c2Q1[i + j] = Q.op3("mulcoeffs_montgomery", m1[i], c0Q1[j])
                    c2Q2[i + j] = M.op3("mulcoeffs_montgomery", m2[i], c0Q2[j])
                    c2Q1[i + j] = Q.op3("add", c2Q1[i + j], c2Q1[i + j])
                    c2Q2[i + j] = M.op3("add", c2Q2[i + j], c2Q2[i + j])
            for i in range(n0):
                Q.op3("mulcoeffs_montgomery_and_add", m1[i], c0Q1[i], c2Q1[i << 1])
                M.op3("mulcoeffs_montgomery_and_add", m2[i], c0Q2[i], c2Q2[i << 1])
        else:  # :407-416
            for i in range(n0):
                a1, a2 = Q.op2("mform_poly", c0Q1[i]), M.op2("mform_poly", c0Q2[i])
                for j in range(n1):
                    Q.op3("mulcoeffs_montgomery_and_add", a1, c1Q1[j], c2Q1[i + j])
                    M.op3("mulcoeffs_montgomery_and_add", a2, c1Q2[j], c2Q2[i + j])
        prod = 1
        for q in M.moduli:
            prod *= q
        phalf = prod >> 1
        pm, pq = arr([phalf % q for q in M.moduli]), arr([phalf % q for q in Q.moduli])
        out = np.zeros((nout, Q.nl, self.N), dtype=np.uint64)
        for i in range(nout):  # :424-463
            x1, x2 = Q.invntt(c2Q1[i]), M.invntt(c2Q2[i])
            x2 = bc.moddown_splited_qp(levelQ, levelM, x1, x2)
            lib().orc_add_scalar(M.h, M.nl, ptr(x2), ptr(pm))
            y = bc.modup_split_pq(levelM, x2)
            lib().orc_sub_scalar(Q.h, Q.nl, ptr(y), ptr(pq))
            out[i] = Q.mul_scalar(y, [self.t] * Q.nl)
        return out

    def relinearize_general(self, ct, evks):
        """relinearize (:480-500) of a ciphertext of any degree >= 2; evks[deg-2] as EvaluationKey.evakey"""
        out = np.ascontiguousarray(ct[:2]).copy()
        for d in range(ct.shape[0] - 1, 1, -1):
            p0, p1 = self.switch_keys_core(np.ascontiguousarray(ct[d]), evks[d - 2])
            out[0] = self.Q.op3("add", np.ascontiguousarray(out[0]), p0)
            out[1] = self.Q.op3("add", np.ascontiguousarray(out[1]), p1)
        return out

    def rotate_columns(self, ct, k, left, right, galois_gen=5):
        """RotateColumns (:578-625) with rotateColumnsPow2 (:637-666); left / right = {k: evk}"""
        N = self.N
        k &= (N >> 1) - 1
        if k == 0:
            return np.ascontiguousarray(ct).copy()
        if k in left:
            return self.permute(np.ascontiguousarray(ct), pow(galois_gen, k, 2 * N), left[k])
        i = 1
        while i < (N >> 1):
            if i not in left or i not in right:
                raise ValueError("cannot RotateColumns: specific rotation and pow2 rotations have not been generated")
            i <<= 1
        if bin(k).count("1") <= bin((N >> 1) - k).count("1"):
            gen, keys = galois_gen, left
        else:
            gen, keys, k = pow(galois_gen, 2 * N - 1, 2 * N), right, (N >> 1) - k
        out = np.ascontiguousarray(ct).copy()
        mask, idx = (N << 1) - 1, 1
        while k > 0:
            if k & 1:
                out = self.permute(out, gen, keys[idx])
            gen = (gen * gen) & mask
            idx <<= 1
            k >>= 1
        return out


# ---------------------------------------------------------------------------
# SimpleScaler (ring/ring_scaling.go:166-300, ring/float128.go) and the BFV key generator /
# encryptor / decryptor / encoder ring sequences (bfv/keygen.go, encryptor.go, decryptor.go,
# encoder.go) restated as compositions of the oracle's ring ops; sampled values are inputs
# ---------------------------------------------------------------------------
def _declare_scaler(L):
    def f(name, res, *args):
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = list(args)

    pd = C.POINTER(C.c_double)
    f("orc_scaler_new", vp, u64, vp)
    f("orc_scaler_free", None, vp)
    f("orc_scaler_params", None, vp, p64, pd)
    f("orc_scaler_scale", None, vp, p64, p64, C.c_int)
    f("orc_f128_op", None, C.c_int, pd, pd, pd)
    f("orc_f128_to_u64", u64, pd)


def _scaler_lib():
    L = lib()
    if not getattr(L, "_scaler_declared", False):
        _declare_scaler(L)
        L._scaler_declared = True
    return L


def f128_op(op, a, b):
    """op 0/1/2 = Float128Add / Mul / Div on (hi, lo) pairs"""
    D = C.c_double * 2
    out = D()
    _scaler_lib().orc_f128_op(op, D(*a), D(*b), out)
    return (out[0], out[1])


def f128_to_u64(a):
    return int(_scaler_lib().orc_f128_to_u64((C.c_double * 2)(*a)))


class Scaler:
    """ring.SimpleScaler: t/Q scaling of an RNS polynomial, result mod t on every limb of the output"""

    def __init__(self, t, ctx):
        self.ctx, self.t = ctx, int(t)
        self.h = _scaler_lib().orc_scaler_new(self.t, ctx.h)

    def __del__(self):
        try:
            lib().orc_scaler_free(self.h)
        except Exception:
            pass

    def params(self):
        wi = np.zeros(self.ctx.nl, dtype=np.uint64)
        ti = np.zeros((self.ctx.nl, 2), dtype=np.float64)
        lib().orc_scaler_params(self.h, ptr(wi), ti.ctypes.data_as(C.POINTER(C.c_double)))
        return wi, ti

    def scale(self, p1, nl_out=1):
        p1 = np.ascontiguousarray(p1)
        out = np.zeros((nl_out, self.ctx.N), dtype=np.uint64)
        lib().orc_scaler_scale(self.h, ptr(p1), ptr(out), nl_out)
        return out


def bit_reverse(x, bits):
    return int(format(x, "0%db" % bits)[::-1], 2) if bits else 0


def bfv_index_matrix(N, galois_gen=5):  # bfv/encoder.go:36-58
    logN = N.bit_length() - 1
    row, m, pos = N >> 1, N << 1, 1
    idx = np.zeros(N, dtype=np.uint64)
    for i in range(row):
        idx[i] = bit_reverse((pos - 1) >> 1, logN)
        idx[i | row] = bit_reverse((m - pos - 1) >> 1, logN)
        pos = (pos * galois_gen) & (m - 1)
    return idx


class BfvScheme:
    def __init__(self, Q, P, N, t):
        self.Qm, self.Pm, self.N, self.t = list(Q), list(P), N, int(t)
        self.Q, self.P, self.QP = Context(N, self.Qm), Context(N, self.Pm), Context(N, self.Qm + self.Pm)
        self.T = Context(N, [self.t])  # contextT, bfv/bfv.go:47
        self.ext = Extender(self.Q, self.P)
        self.nQ, self.alpha = len(self.Qm), len(self.Pm)
        self.beta = -(-self.nQ // self.alpha)
        self.Pbig = 1
        for p in self.Pm:
            self.Pbig *= int(p)
        Qbig = 1
        for q in self.Qm:
            Qbig *= int(q)
        delta = Qbig // self.t  # GenLiftParams, bfv/utils.go:9-23
        self.delta_mont = [int(lib().orc_mform(delta % q, q, ptr(arr(bred_params(q))))) for q in self.Qm]
        self.index_matrix = bfv_index_matrix(N)
        self.scaler = Scaler(self.t, self.Q)

    # --- keygen.go
    def gen_secret_key(self, ternary):  # :82-96
        return self.QP.ntt(self.QP.op2("mform_poly", signed_residues(self.Qm + self.Pm, ternary)))

    def gen_public_key(self, sk, e, a):  # :120-135
        pk0 = self.QP.ntt(signed_residues(self.Qm + self.Pm, e))
        self.QP.op3("mulcoeffs_montgomery_and_add", sk, np.ascontiguousarray(a), pk0)
        return self.QP.op2("neg", pk0), np.ascontiguousarray(a).copy()

    def _times_p(self, p):
        return self.QP.mul_scalar(p, [self.Pbig % q for q in self.Qm + self.Pm])

    def new_switching_key(self, sk_in, sk_out, errors, uniforms):  # newswitchingkey :285-334 (sk_in already times P)
        QP = self.Qm + self.Pm
        evk = np.zeros((self.beta, 2, len(QP), self.N), dtype=np.uint64)
        for i in range(self.beta):
            k0 = self.QP.op2("mform_poly", self.QP.ntt(signed_residues(QP, errors[i])))
            k1 = np.ascontiguousarray(uniforms[i]).copy()
            for j in range(self.alpha):
                index = i * self.alpha + j
                qi = np.uint64(QP[index])
                t = k0[index] + sk_in[index]
                k0[index] = np.where(t >= qi, t - qi, t)  # CRed :319
                if index >= len(QP) - 1:  # :323 (the bound is the QP context's, not Q's)
                    break
            self.QP.op3("mulcoeffs_montgomery_and_sub", k1, sk_out, k0)
            evk[i, 0], evk[i, 1] = k0, k1
        return evk

    def gen_relin_key(self, sk, errors, uniforms):  # GenRelinKey :171-195, maxDegree 1
        pool = self._times_p(sk)
        pool = self.QP.op3("mulcoeffs_montgomery", pool, sk)
        return self.new_switching_key(pool, sk, errors, uniforms)

    def gen_switching_key(self, sk_in, sk_out, errors, uniforms):  # :248-262
        return self.new_switching_key(self._times_p(sk_in), sk_out, errors, uniforms)

    def gen_rot_key(self, sk, gen, errors, uniforms):  # genrotkey :429-441
        idx = permute_ntt_index(gen, 1, self.N)
        return self.new_switching_key(self._times_p(permute_ntt_with_index(sk, idx)), sk, errors, uniforms)

    # --- encryptor.go
    def encrypt_pk(self, pt, pk, u, e0, e1, fast=False, ct=None):  # pkEncryptor.encrypt :168-222
        nQ = self.nQ
        ct = np.zeros((2, nQ, self.N), dtype=np.uint64) if ct is None else ct.copy()
        if fast:
            # :174-192 computes pk*u + e into the pools and never copies them to the ciphertext: the
            # receiver only gets the plaintext added (:221).  Kept literal.
            pass
        else:
            QP = self.Qm + self.Pm
            up = self.QP.ntt(self.QP.op2("mform_poly", signed_residues(QP, u)))
            p0 = self.QP.invntt(self.QP.op3("mulcoeffs_montgomery", up, pk[0]))
            p1 = self.QP.invntt(self.QP.op3("mulcoeffs_montgomery", up, pk[1]))
            p0 = self.QP.op3("add", p0, signed_residues(QP, e0))
            p1 = self.QP.op3("add", p1, signed_residues(QP, e1))
            ct[0] = self.ext.moddown_pq(nQ - 1, p0)[:nQ]
            ct[1] = self.ext.moddown_pq(nQ - 1, p1)[:nQ]
        ct[0] = self.Q.op3("add", np.ascontiguousarray(ct[0]), np.ascontiguousarray(pt))
        return ct

    def encrypt_sk(self, pt, sk, crp, e, fast=False):  # skEncryptor.encrypt :296-345
        nQ = self.nQ
        ct = np.zeros((2, nQ, self.N), dtype=np.uint64)
        if fast:
            c0 = self.Q.op2("neg", self.Q.op3("mulcoeffs_montgomery", np.ascontiguousarray(crp), np.ascontiguousarray(sk[:nQ])))
            ct[0] = self.Q.op3("add", self.Q.invntt(c0), signed_residues(self.Qm, e))
            ct[1] = self.Q.invntt(np.ascontiguousarray(crp))
        else:
            QP = self.Qm + self.Pm
            p0 = self.QP.invntt(self.QP.op2("neg", self.QP.op3("mulcoeffs_montgomery", np.ascontiguousarray(crp), sk)))
            a = self.QP.invntt(np.ascontiguousarray(crp))
            p0 = self.QP.op3("add", p0, signed_residues(QP, e))
            ct[0] = self.ext.moddown_pq(nQ - 1, p0)[:nQ]
            ct[1] = self.ext.moddown_pq(nQ - 1, a)[:nQ]
        ct[0] = self.Q.op3("add", np.ascontiguousarray(ct[0]), np.ascontiguousarray(pt))
        return ct

    # --- decryptor.go:55-75
    def decrypt(self, ct, sk):
        nQ = self.nQ
        skq = np.ascontiguousarray(sk[:nQ])
        degree = len(ct) - 1
        pt = self.Q.ntt(np.ascontiguousarray(ct[degree]))
        for i in range(degree, 0, -1):
            pt = self.Q.op3("mulcoeffs_montgomery", pt, skq)
            pt = self.Q.op3("add", pt, self.Q.ntt(np.ascontiguousarray(ct[i - 1])))
            if i & 7 == 7:
                pt = self.Q.op2("reduce", pt)
        if degree & 7 != 7:
            pt = self.Q.op2("reduce", pt)
        return self.Q.invntt(pt)

    # --- encoder.go
    def encode_uint(self, coeffs):  # EncodeUint :69-90 + encodePlaintext :121-136
        c = np.asarray(coeffs, dtype=np.uint64)
        slots = np.zeros((1, self.N), dtype=np.uint64)
        slots[0, self.index_matrix[: len(c)].astype(np.int64)] = c
        m = self.T.invntt(slots)[0]
        pt = np.zeros((self.nQ, self.N), dtype=np.uint64)
        for i in range(self.nQ - 1, -1, -1):
            q = self.Qm[i]
            pt[i] = _mred_vec(m, self.delta_mont[i], q, int(lib().orc_mred_params(q)))
        return pt

    def encode_int(self, coeffs):  # EncodeInt :94-119
        c = np.asarray(coeffs, dtype=np.int64)
        return self.encode_uint(np.where(c < 0, np.int64(self.t) + c, c).astype(np.uint64))

    def decode_uint(self, pt):  # DecodeUint :139-153
        pool = self.T.ntt(self.scaler.scale(pt, 1))
        return pool[0, self.index_matrix.astype(np.int64)].copy()

    def decode_int(self, pt):  # DecodeInt :157-182
        v = self.decode_uint(pt).astype(np.int64)
        return np.where(v > (self.t >> 1), v - self.t, v)


# ---------------------------------------------------------------------------
# Refresh protocols (dckks/public_refresh.go:43-147, dbfv/public_refresh.go:105-205) restated over the oracle's ring
# ops.  Sampled values (masks, errors) are inputs; the big-integer steps (SetCoefficientsBigint, PolyToBigint,
# ring/ring_context.go:343-421) are exact Python integers.
# ---------------------------------------------------------------------------
def set_coefficients_bigint(moduli, coeffs):  # ring_context.go:343-367 (big.Int.Mod is Euclidean, like Python's %)
    return np.array([[int(c) % int(q) for c in coeffs] for q in moduli], dtype=np.uint64)


def poly_to_bigint(poly, moduli):  # ring_context.go:384-421: the CRT value in [0, prod(moduli of the poly's limbs))
    nl = poly.shape[0]
    mods = [int(q) for q in moduli[:nl]]
    big = 1
    for q in mods:
        big *= q
    rec = [(big // q) * pow(big // q, -1, q) for q in mods]
    cols = [poly[i].tolist() for i in range(nl)]
    return [sum(cols[i][x] * rec[i] for i in range(nl)) % big for x in range(poly.shape[1])]


class DckksRefresh:
    """dckks/public_refresh.go over contextQ of a CkksScheme"""

    def __init__(self, scheme):
        self.S = scheme

    def center_mask(self, level_start, n_parties, raw):  # :48-63, raw[i] = ring.RandInt(bound)
        bound = 1
        for q in self.S.Qm[: level_start + 1]:
            bound *= int(q)
        bound //= 2 * n_parties
        half = bound >> 1
        assert all(0 <= m < bound for m in raw)
        return [m - bound if m >= half else m for m in raw]

    def gen_shares(self, sk, level_start, n_parties, ct1, crs, raw_mask, e0, e1):  # :43-98
        S = self.S
        nl, nQ = level_start + 1, S.levels
        skq = np.ascontiguousarray(sk[:nQ])
        mask = self.center_mask(level_start, n_parties, raw_mask)
        h0 = S.Q.ntt(set_coefficients_bigint(S.Qm[:nl], mask), nl=nl)  # :66,:75
        h1 = S.Q.ntt(set_coefficients_bigint(S.Qm, mask))  # :68,:76
        S.Q.op3("mulcoeffs_montgomery_and_add", np.ascontiguousarray(skq[:nl]), np.ascontiguousarray(ct1[:nl]), h0, nl=nl)  # :79
        S.Q.op3("mulcoeffs_montgomery_and_add", skq, np.ascontiguousarray(crs), h1)  # :82
        tmp = S.Q.ntt(signed_residues(S.Qm, e0))  # SampleNTT :85
        h0 = S.Q.op3("add", h0, np.ascontiguousarray(tmp[:nl]), nl=nl)
        tmp = S.Q.ntt(signed_residues(S.Qm, e1))  # :89
        h1 = S.Q.op3("add", h1, tmp)
        return h0, S.Q.op2("neg", h1)  # :93

    def aggregate(self, a, b):  # :101-103
        return self.S.Q.op3("add", a, b, nl=a.shape[0])

    def decrypt(self, ct0, share_decrypt):  # :106-108
        return self.S.Q.op3("add", np.ascontiguousarray(ct0), share_decrypt, nl=ct0.shape[0])

    def recode(self, ct0):  # :111-139: value[0] at level len(ct0)-1 -> the same centred values over every limb of Q
        S = self.S
        nl = ct0.shape[0]
        vals = poly_to_bigint(S.Q.invntt(np.ascontiguousarray(ct0), nl=nl), S.Qm)
        qstart = 1
        for q in S.Qm[:nl]:
            qstart *= int(q)
        half = qstart >> 1
        vals = [v - qstart if v >= half else v for v in vals]
        return S.Q.ntt(set_coefficients_bigint(S.Qm, vals))

    def recrypt(self, ct0, crs, share_recrypt):  # :142-147
        return np.stack([self.S.Q.op3("add", ct0, share_recrypt), np.ascontiguousarray(crs)])


class DbfvRefresh:
    """dbfv/public_refresh.go over a BfvScheme.  hP is protocol state: GenShares adds the P limbs of the error sample
    into it without ever clearing it (:137-143), so a second call on the same object sees the first call's words."""

    def __init__(self, scheme):
        self.S = scheme
        self.hP = np.zeros((scheme.alpha, scheme.N), dtype=np.uint64)

    def lift(self, p0):  # :207-214: MRed(p0.Coeffs[0], deltaMont[i]) into every limb
        S = self.S
        out = np.zeros((S.nQ, S.N), dtype=np.uint64)
        for i in range(S.nQ - 1, -1, -1):
            q = S.Qm[i]
            out[i] = _mred_vec(np.ascontiguousarray(p0[0]), S.delta_mont[i], q, int(lib().orc_mred_params(q)))
        return out

    def gen_shares(self, sk, ct1, crs, e, e_prime, mask):  # :105-169
        S = self.S
        nQ = S.nQ
        QP = S.Qm + S.Pm
        skq = np.ascontiguousarray(sk[:nQ])
        h0 = S.Q.invntt(S.Q.op3("mulcoeffs_montgomery", skq, S.Q.ntt(np.ascontiguousarray(ct1))))  # :116-119
        h0 = S.Q.mul_scalar(h0, [S.Pbig % q for q in S.Qm])  # :122
        tmp1 = signed_residues(QP, e)  # sampler.Sample over QP :125
        h0 = S.Q.op3("add", h0, np.ascontiguousarray(tmp1[:nQ]))  # :126
        self.hP = self.hP + tmp1[nQ:]  # :128-134 (plain uint64 +=)
        h0 = S.ext.moddown_splited_pq(nQ - 1, h0, np.ascontiguousarray(self.hP))  # :137
        t2 = S.QP.op3("mulcoeffs_montgomery", np.ascontiguousarray(sk), S.QP.ntt(np.ascontiguousarray(crs)))  # :140-141
        t2 = S.QP.invntt(S.QP.op2("neg", t2))  # :142-143
        t2 = S.QP.op3("add", t2, signed_residues(QP, e_prime))  # SampleAndAdd :146
        h1 = S.ext.moddown_pq(nQ - 1, t2)[:nQ].copy()  # :149
        m = self.lift(np.asarray(mask, dtype=np.uint64)[None, :])  # :152-153
        return S.Q.op3("add", h0, m), S.Q.op3("sub", h1, m)  # :156,:159

    def aggregate(self, a, b):  # :172-175
        return self.S.Q.op3("add", a[0], b[0]), self.S.Q.op3("add", a[1], b[1])

    def finalize(self, ct, crs, share):  # Decrypt :178-180, Recode :183-188, Recrypt :191-199
        S = self.S
        nQ = S.nQ
        pt = S.Q.op3("add", np.ascontiguousarray(ct[0]), share[0])
        pt = self.lift(S.scaler.scale(pt, nQ))
        c0 = S.Q.op3("add", pt, share[1])
        c1 = S.ext.moddown_pq(nQ - 1, np.ascontiguousarray(crs))[:nQ].copy()
        return np.stack([c0, c1])


# ---------------------------------------------------------------------------
# wire formats of the scheme objects (ckks/marshaler.go, bfv/marshaler.go): byte strings built from numpy
# polynomials [nlimbs][N]; switching keys are [beta][2][nQP][N]
# ---------------------------------------------------------------------------
import struct as _struct


def ckks_ciphertext_marshal(value, scale, is_ntt=True):  # ckks/marshaler.go:24-52
    head = bytes([len(value)]) + _struct.pack("<d", scale) + bytes([0, 1 if is_ntt else 0])
    return head + b"".join(poly_marshal(p) for p in value)


def bfv_ciphertext_marshal(value, is_ntt=False):  # bfv/marshaler.go:9-33
    return bytes([len(value), 1 if is_ntt else 0]) + b"".join(poly_marshal(p) for p in value)


def _poly_at(data, pointer):  # DecodePolyNew ring_object.go:277-289
    N, nl = 1 << data[pointer], data[pointer + 1]
    inc = 2 + ((N * nl) << 3)
    return poly_unmarshal(data[pointer:pointer + inc]), inc


def ckks_ciphertext_unmarshal(data):  # ckks/marshaler.go:57-91
    scale = _struct.unpack("<d", data[1:9])[0]
    value, pointer = [], 11
    for _ in range(data[0]):
        p, inc = _poly_at(data, pointer)
        value.append(p)
        pointer += inc
    return value, scale, data[10] == 1


def bfv_ciphertext_unmarshal(data):  # bfv/marshaler.go:36-60
    value, pointer = [], 2
    for _ in range(data[0]):
        p, inc = _poly_at(data, pointer)
        value.append(p)
        pointer += inc
    return value, data[1] == 1


def public_key_marshal(pk):  # ckks/marshaler.go:133-143
    return poly_marshal(pk[0]) + poly_marshal(pk[1])


def swk_marshal(evk):  # SwitchingKey.encode ckks/marshaler.go:230-257
    return bytes([evk.shape[0]]) + b"".join(poly_marshal(evk[j, h]) for j in range(evk.shape[0]) for h in (0, 1))


def swk_unmarshal(data, pointer=0):  # decode :259-283
    beta, start = data[pointer], pointer
    pointer += 1
    polys = []
    for _ in range(2 * beta):
        p, inc = _poly_at(data, pointer)
        polys.append(p)
        pointer += inc
    return np.stack(polys).reshape(beta, 2, *polys[0].shape), pointer - start


def bfv_evaluation_key_marshal(keys):  # bfv/marshaler.go:166-183
    return bytes([len(keys)]) + b"".join(swk_marshal(k) for k in keys)


def rotation_keys_marshal(left, right, third=None):  # ckks/marshaler.go:312-355: type byte over the big-endian amount
    out = []
    for typ, keys in ((2, left), (1, right)):
        for i, k in keys.items():
            out.append(bytes([typ]) + _struct.pack(">I", i)[1:] + swk_marshal(k))
    if third is not None:
        out.append(bytes([3, 0, 0, 0]) + swk_marshal(third))
    return b"".join(out)
