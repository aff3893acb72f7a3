"""Measures the BASELINE.json configurations other than the headline one (which bench.py owns) on
one GPU, each beside the CPU oracle on a bounded sample.  One JSON line per configuration.

  C1  ring.Context NTT / InvNTT / MulCoeffsMontgomery, N=2^13, 4 x 60-bit limbs (ring_benchmark_test.go shapes)
  C2  CKKS PN14QP438: MulRelin + Rescale, batch of 1024 ciphertexts
  C3  BFV  PN15QP880: Mul + Relinearize + RotateColumns(1), batched
  C5  dckks PN15QP880: CKG.GenShare and PCKS.GenShare per party (N=2^15)

Usage: python profiles/tools/bench_configs.py [c1 c2 c3 c5]
"""
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "lattigo-fhe-by-go_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

import lattigpu
from lattigpu import bfv, ckks, dckks, ring
from oracle import ring_oracle as orc

QI60_TAIL = [1152921504066306049, 1152921504057917441, 1152921504053723137, 1152921504050839553]  # ring/params.go:12
DEV = torch.device("cuda", 0)


def uniform(shape_prefix, moduli, N, g):
    t = torch.empty(*shape_prefix, len(moduli), N, dtype=torch.int64, device=DEV)
    for i, q in enumerate(moduli):
        t[..., i, :] = torch.randint(0, q, (*shape_prefix, N), dtype=torch.int64, device=DEV, generator=g)
    return t


def wrap(t, N, nl, batch):
    return ring.Poly.wrap(t.data_ptr(), N, nl, batch, keep=t)


def gpu_time(fn, reps=10, warmup=3):
    sp = torch.cuda.current_stream()
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(sp)
    for _ in range(reps):
        fn()
    e1.record(sp)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3  # seconds per call


def cpu_parallel(make_worker, nops):
    """one oracle evaluator per host thread, ops pulled from a shared counter (psi.go:214-233 pattern)"""
    cores = os.cpu_count() or 1
    workers = [make_worker() for _ in range(cores)]
    nxt = {"i": 0}
    lock = threading.Lock()

    def run(w):
        while True:
            with lock:
                if nxt["i"] >= nops:
                    return
                nxt["i"] += 1
            w()

    t0 = time.perf_counter()
    ths = [threading.Thread(target=run, args=(w,)) for w in workers]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    return nops / (time.perf_counter() - t0), cores


def sp():
    return torch.cuda.current_stream().cuda_stream


def c1():
    N, Q, B = 1 << 13, QI60_TAIL, 4096
    g = torch.Generator(device=DEV)
    g.manual_seed(0x1A771C0 + 1)
    ctx = ring.NewContextWithParams(N, Q)
    a_t, b_t = uniform((B,), Q, N, g), uniform((B,), Q, N, g)
    o_t = torch.empty_like(a_t)
    a, b, o = wrap(a_t, N, 4, B), wrap(b_t, N, 4, B), wrap(o_t, N, 4, B)
    res = {}
    for name, fn in (("NTT", lambda: ctx.NTT(a, o, stream=sp())), ("InvNTT", lambda: ctx.InvNTT(a, o, stream=sp())),
                     ("MulCoeffsMontgomery", lambda: ctx.MulCoeffsMontgomery(a, b, o, stream=sp()))):
        s = gpu_time(fn, reps=20)
        res[name] = {"polys_per_s": B / s, "limb_ops_per_s": 4 * B / s, "us_per_batch": s * 1e6,
                     "hbm_GBps": (3 if name == "MulCoeffsMontgomery" else 2) * 4 * B * N * 8 / s / 1e9}
    oc = orc.Context(N, Q)
    x = a_t[0].cpu().numpy().astype(np.uint64)
    y = b_t[0].cpu().numpy().astype(np.uint64)
    cpu = {}
    for name, f in (("NTT", lambda: oc.ntt(x)), ("InvNTT", lambda: oc.invntt(x)),
                    ("MulCoeffsMontgomery", lambda: oc.op3("mulcoeffs_montgomery", x, y))):
        v, cores = cpu_parallel(lambda f=f: f, 2000)
        cpu[name] = {"polys_per_s": v, "cores": cores}
    return {"config": "C1 ring N=2^13 x 4 limbs (60-bit), batch 4096", "gpu": res, "cpu_oracle": cpu}


def c2():
    p = ckks.DefaultParams[ckks.PN14QP438]
    N = 1 << p["LogN"]
    Q, P = ckks.GenModuli(p)
    nQ, nP = len(Q), len(P)
    beta = -(-nQ // nP)
    B = 1024
    g = torch.Generator(device=DEV)
    g.manual_seed(0x1A771C0 + 2)
    cQ, cP = ring.NewContextWithParams(N, Q), ring.NewContextWithParams(N, P)
    ev = ckks.NewEvaluator(cQ, cP)
    evk_t = uniform((beta, 2), Q + P, N, g)
    rlk = ckks.SwitchingKey(N=N, device_ptr=evk_t.data_ptr(), beta=beta, nQP=nQ + nP, keep=evk_t)
    a_t = [uniform((B,), Q, N, g) for _ in range(2)]
    b_t = [uniform((B,), Q, N, g) for _ in range(2)]
    o_t = [torch.empty(B, nQ, N, dtype=torch.int64, device=DEV) for _ in range(2)]
    a, b, o = (tuple(wrap(t, N, nQ, B) for t in ts) for ts in (a_t, b_t, o_t))
    level = nQ - 1

    def step():
        ev.MulRelin(level, a, b, rlk, o, stream=sp())
        ev.Rescale(nQ, o, 1, stream=sp())

    s = gpu_time(step, reps=5)
    oQ, oP = orc.Context(N, Q), orc.Context(N, P)
    evk = evk_t.cpu().numpy().astype(np.uint64)
    x = np.ascontiguousarray(np.stack([a_t[0][0].cpu().numpy(), a_t[1][0].cpu().numpy()]).astype(np.uint64))
    y = np.ascontiguousarray(np.stack([b_t[0][0].cpu().numpy(), b_t[1][0].cpu().numpy()]).astype(np.uint64))

    def mk():
        e = orc.CkksEvaluator(oQ, oP)
        return lambda: e.rescale(e.mul_relin(level, x, y, evk))

    v, cores = cpu_parallel(mk, 8 * (os.cpu_count() or 1))
    out = {"config": "C2 CKKS PN14QP438 MulRelin+Rescale, batch 1024", "gpu_ops_per_s": B / s, "ms_per_batch": s * 1e3,
           "cpu_oracle_ops_per_s": v, "cpu_cores": cores}

    # ---- the whole config: encrypt (pk, ModDown path) x2 -> MulRelin -> Rescale -> decrypt, device-resident.
    # Sampling is host work in the reference (crypto/rand); here the small samples are drawn on the device with
    # torch and only expanded to residues, then every ring op of encryptor.go / decryptor.go runs through the ABI.
    from lattigpu import ckks_scheme

    kg = ckks_scheme.KeyGenerator(cQ, cP)
    K = kg.contextQP
    rs = np.random.default_rng(5)
    sk = kg.GenSecretKey(rs.integers(-1, 2, size=N))
    pk = kg.GenPublicKey(sk, np.rint(rs.normal(0, 3.2, size=N)).astype(np.int64),
                         np.stack([rs.integers(0, q, size=N, dtype=np.uint64) for q in Q + P]))
    rlk2, rlk2_host = kg.GenRelinKey(sk, [np.rint(rs.normal(0, 3.2, size=N)).astype(np.int64) for _ in range(beta)],
                             [np.stack([rs.integers(0, q, size=N, dtype=np.uint64) for q in Q + P]) for _ in range(beta)])
    enc = ckks_scheme.Encryptor(cQ, cP, K, pk=pk, sk=sk)
    dec = ckks_scheme.Decryptor(cQ, sk)
    Bf = 256  # the ModDown-path encryption holds 6 QP-sized pools per batch entry
    pts = [wrap(uniform((Bf,), Q, N, g), N, nQ, Bf) for _ in range(2)]
    cts = [(wrap(torch.empty(Bf, nQ, N, dtype=torch.int64, device=DEV), N, nQ, Bf),
            wrap(torch.empty(Bf, nQ, N, dtype=torch.int64, device=DEV), N, nQ, Bf)) for _ in range(3)]
    ptout = wrap(torch.empty(Bf, nQ, N, dtype=torch.int64, device=DEV), N, nQ, Bf)
    mods = torch.tensor(Q + P, dtype=torch.int64, device=DEV)

    def residues(c):  # [Bf][N] small signed -> [Bf][nQP][N] residues
        t = torch.where(c[:, None, :] < 0, c[:, None, :] + mods[None, :, None], c[:, None, :]).contiguous()
        return wrap(t, N, nQ + nP, Bf)

    def full():
        for k in range(2):
            u = residues(torch.randint(-1, 2, (Bf, N), dtype=torch.int64, device=DEV, generator=g))
            e0 = residues(torch.round(torch.randn(Bf, N, device=DEV, generator=g) * 3.2).to(torch.int64))
            e1 = residues(torch.round(torch.randn(Bf, N, device=DEV, generator=g) * 3.2).to(torch.int64))
            enc.EncryptPk(level, pts[k], cts[k], u, e0, e1, stream=sp())
        ev.MulRelin(level, cts[0], cts[1], rlk2, cts[2], stream=sp())
        ev.Rescale(nQ, cts[2], 1, stream=sp())
        dec.Decrypt(level - 1, cts[2], ptout, stream=sp())

    sf = gpu_time(full, reps=3, warmup=2)
    S = orc.CkksScheme(Q, P, N)
    osk = sk.numpy()
    opk = (pk[0].numpy(), pk[1].numpy())
    orlk = rlk2_host
    pt0 = pts[0].numpy(squeeze=False)[0]

    def mkfull():
        e = orc.CkksEvaluator(S.Q, S.P)
        r = np.random.default_rng(9)

        def one():
            c = [S.encrypt_pk(level, pt0, opk, r.integers(-1, 2, size=N), np.rint(r.normal(0, 3.2, size=N)).astype(np.int64),
                              np.rint(r.normal(0, 3.2, size=N)).astype(np.int64)) for _ in range(2)]
            w = e.rescale(e.mul_relin(level, np.ascontiguousarray(c[0]), np.ascontiguousarray(c[1]), orlk))
            return S.decrypt(level - 1, w, osk)

        return one

    vf, cores = cpu_parallel(mkfull, 2 * (os.cpu_count() or 1))
    out["full_pipeline"] = {"what": "encrypt(pk) x2 -> MulRelin -> Rescale -> decrypt, batch %d, device-resident" % Bf,
                            "gpu_ops_per_s": Bf / sf, "ms_per_batch": sf * 1e3, "cpu_oracle_ops_per_s": vf, "cpu_cores": cores}
    return out


def c3():
    p = bfv.DefaultParams[bfv.PN15QP880]
    N = 1 << p["LogN"]
    Q, P, QMul = bfv.GenModuli(p)
    nQ, nP = len(Q), len(P)
    beta = -(-nQ // nP)
    B = 64
    g = torch.Generator(device=DEV)
    g.manual_seed(0x1A771C0 + 3)
    cQ, cM, cP = (ring.NewContextWithParams(N, m) for m in (Q, QMul, P))
    ev = bfv.NewEvaluator(cQ, cM, cP, p["T"])
    evk_t = uniform((beta, 2), Q + P, N, g)
    key = ckks.SwitchingKey(N=N, device_ptr=evk_t.data_ptr(), beta=beta, nQP=nQ + nP, keep=evk_t)
    a_t = [uniform((B,), Q, N, g) for _ in range(2)]
    b_t = [uniform((B,), Q, N, g) for _ in range(2)]
    d2_t = [torch.empty(B, nQ, N, dtype=torch.int64, device=DEV) for _ in range(3)]
    d1_t = [torch.empty(B, nQ, N, dtype=torch.int64, device=DEV) for _ in range(2)]
    r_t = [torch.empty(B, nQ, N, dtype=torch.int64, device=DEV) for _ in range(2)]
    a, b, d2, d1, r = (tuple(wrap(t, N, nQ, B) for t in ts) for ts in (a_t, b_t, d2_t, d1_t, r_t))
    gen = pow(bfv.GaloisGen, 1, 2 * N)

    def step():
        ev.Mul(a, b, d2, stream=sp())
        ev.Relinearize(d2, key, d1, stream=sp())
        ev.permute(d1, gen, key, r, stream=sp())

    s = gpu_time(step, reps=5)
    oev_ctx = (orc.Context(N, Q), orc.Context(N, QMul), orc.Context(N, P))
    evk = evk_t.cpu().numpy().astype(np.uint64)
    x = np.ascontiguousarray(np.stack([a_t[0][0].cpu().numpy(), a_t[1][0].cpu().numpy()]).astype(np.uint64))
    y = np.ascontiguousarray(np.stack([b_t[0][0].cpu().numpy(), b_t[1][0].cpu().numpy()]).astype(np.uint64))

    def mk():
        e = orc.BfvEvaluator(*oev_ctx, p["T"])
        return lambda: e.permute(e.relinearize(e.tensor_and_rescale(x, y), evk), gen, evk)

    v, cores = cpu_parallel(mk, 2 * (os.cpu_count() or 1))
    return {"config": "C3 BFV PN15QP880 Mul+Relinearize+RotateColumns(1), batch 64", "gpu_ops_per_s": B / s,
            "ms_per_batch": s * 1e3, "cpu_oracle_ops_per_s": v, "cpu_cores": cores}


def c5():
    p = ckks.DefaultParams[ckks.PN15QP880]
    N = 1 << p["LogN"]
    Q, P = ckks.GenModuli(p)
    QP = Q + P
    nQ = len(Q)
    B = 8  # parties processed side by side on one GPU (one per GPU with lattigpu.dist.Comm)
    g = torch.Generator(device=DEV)
    g.manual_seed(0x1A771C0 + 5)
    cQ, cP, cK = (ring.NewContextWithParams(N, m) for m in (Q, P, QP))
    ckg, pcks = dckks.CKGProtocol(cK), dckks.PCKSProtocol(cQ, cP, cK)
    mk = lambda mods: uniform((B,), mods, N, g)
    sk_t, crs_t, e_t, sh_t = mk(QP), mk(QP), mk(QP), mk(QP)
    sk, crs, e, sh = (wrap(t, N, len(QP), B) for t in (sk_t, crs_t, e_t, sh_t))
    s_ckg = gpu_time(lambda: ckg.GenShare(sk, crs, sh, e, stream=sp()), reps=10)
    pk_t, u_t, e0_t, e1_t, ct1_t, skq_t = (mk(QP), mk(QP)), mk(QP), mk(QP), mk(QP), mk(Q), mk(Q)
    pk = tuple(wrap(t, N, len(QP), B) for t in pk_t)
    u, e0, e1 = (wrap(t, N, len(QP), B) for t in (u_t, e0_t, e1_t))
    ct1, skq = wrap(ct1_t, N, nQ, B), wrap(skq_t, N, nQ, B)
    share = pcks.AllocateShares(nQ - 1, B)
    s_pcks = gpu_time(lambda: pcks.GenShare(nQ - 1, skq, pk, ct1, share, u, e0, e1, stream=sp()), reps=5)
    # CPU: same sequences from the oracle's ring ops, one party per thread
    oQ, oP, oK = orc.Context(N, Q), orc.Context(N, P), orc.Context(N, QP)
    h = lambda t: np.ascontiguousarray(t[0].cpu().numpy().astype(np.uint64))
    hsk, hcrs, he, hu, he0, he1, hct1, hskq = (h(t) for t in (sk_t, crs_t, e_t, u_t, e0_t, e1_t, ct1_t, skq_t))
    hpk = (h(pk_t[0]), h(pk_t[1]))

    def mk_ckg():
        def f():
            w = oK.ntt(he)
            oK.op3("mulcoeffs_montgomery_and_sub", hsk, hcrs, w)
        return f

    def mk_pcks():
        ext = orc.Extender(oQ, oP)

        def f():
            t = oK.ntt(hu)
            s0 = oK.op3("add", oK.op3("mulcoeffs_montgomery", t, hpk[0]), oK.ntt(he0))
            s1 = oK.op3("add", oK.op3("mulcoeffs_montgomery", t, hpk[1]), oK.ntt(he1))
            w0 = ext.moddown_ntt_pq(nQ - 1, s0)
            ext.moddown_ntt_pq(nQ - 1, s1)
            oQ.op3("mulcoeffs_montgomery_and_add", hct1, hskq, w0)
        return f

    v_ckg, cores = cpu_parallel(mk_ckg, 4 * (os.cpu_count() or 1))
    v_pcks, _ = cpu_parallel(mk_pcks, 2 * (os.cpu_count() or 1))
    return {"config": "C5 dckks PN15QP880 (N=2^15, 18+3 limbs), 8 parties side by side on one GPU",
            "gpu_ckg_genshare_per_s": B / s_ckg, "gpu_pcks_genshare_per_s": B / s_pcks,
            "cpu_oracle_ckg_genshare_per_s": v_ckg, "cpu_oracle_pcks_genshare_per_s": v_pcks, "cpu_cores": cores}


def main():
    ring.set_device(0)
    which = [a.lower() for a in sys.argv[1:]] or ["c1", "c2", "c3", "c5"]
    for name in which:
        out = {"c1": c1, "c2": c2, "c3": c3, "c5": c5}[name]()
        print(json.dumps(out), flush=True)
        torch.cuda.empty_cache()


main()
