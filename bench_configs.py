"""BASELINE.json configurations other than the headline one, for `bench.py --config C1|C2|C3|C5`.

Each builder returns a dict that bench.py times with the same harness as the headline (CUDA events on the launching
stream, max over ranks, clock sampling, e2e through host buffers, CPU baseline on a bounded sample):

    workload        text for config.workload
    metric, unit    what `value` counts
    units_per_step  units one call of step() processes on this rank
    step            callable: the hot path on device-resident inputs (stream = torch's current stream)
    e2e             (inputs, outputs): lists of (device tensor, words to copy) moved host->device before and
                    device->host after step() inside the e2e timed region
    cpu             callable(nthreads) -> (units per second, sample description): the oracle on host threads
    roofline        callable(us_per_step) -> roofline dict (binding bound first, the other beside it)

  C1  ring.Context NTT, N=2^13, 4 x 60-bit limbs (ring_benchmark_test.go shapes), batch 4096; InvNTT and
      MulCoeffsMontgomery reported beside it
  C2  CKKS PN14QP438: MulRelin + Rescale, batch of 1024 ciphertexts (the encrypt -> ... -> decrypt pipeline beside it)
  C3  BFV  PN15QP880: Mul + Relinearize + RotateColumns(1), batch 64
  C5  dckks PN15QP880: CKG.GenShare + PCKS.GenShare for 8 parties (aggregation across GPUs: the `party` leg)
"""
import os
import threading
import time

import numpy as np
import torch

QI60_TAIL = [1152921504066306049, 1152921504057917441, 1152921504053723137, 1152921504050839553]  # ring/params.go:12
# register-resident butterfly rates on B200, butterflies per second (profiles/r01_butterfly_peaks.txt,
# profiles/r02_fp64_butterfly.txt): FP64-only (q < 3*2^44), 16-instruction Shoup (q < 2^56), [0,8q) (q < 2^61)
PEAK_BF = {"d64": 2.074e12, "free": 1.151e12, "lazy": 0.950e12}


def _bf_class(q):
    return "d64" if q < (3 << 44) else ("free" if q < (1 << 56) else "lazy")


def int_floor_us(limb_ntts_by_modulus, N):
    """time the butterflies alone need at the register-resident rates: {modulus: number of limb-NTTs}"""
    logN = N.bit_length() - 1
    return 1e6 * sum(n * (N // 2) * logN / PEAK_BF[_bf_class(q)] for q, n in limb_ntts_by_modulus.items())


def uniform(dev, shape_prefix, moduli, N, g):
    t = torch.empty(*shape_prefix, len(moduli), N, dtype=torch.int64, device=dev)
    for i, q in enumerate(moduli):
        t[..., i, :] = torch.randint(0, q, (*shape_prefix, N), dtype=torch.int64, device=dev, generator=g)
    return t


def cpu_parallel(make_worker, nops, nthreads):
    """one oracle evaluator per host thread, ops pulled from a shared counter (psi.go:214-233 pattern)"""
    workers = [make_worker() for _ in range(nthreads)]
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
    dt = time.perf_counter() - t0
    return nops / dt, dt


def _sp():
    return torch.cuda.current_stream().cuda_stream


def _roofline(kernel, int_us, bytes_, us, hbm_peak, peak_src, note):
    hbm_us = 1e6 * bytes_ / (hbm_peak * 1e9)
    bound = "int" if int_us >= hbm_us else "hbm"
    out = {"bound": bound, "kernel": kernel, "traffic": None, "note": note,
           "int": {"floor_us": int_us, "frac": int_us / us,
                   "peak_source": "register-resident butterfly rates on B200 (profiles/r02_fp64_butterfly.txt, r01_butterfly_peaks.txt)"},
           "hbm": {"algorithmic_bytes": bytes_, "achieved": bytes_ / (us * 1e-6) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                   "frac": hbm_us / us, "peak_source": peak_src}}
    if bound == "int":
        out.update({"achieved": 1e6 / us, "peak": 1e6 / int_us, "unit": "steps/s", "frac": int_us / us})
    else:
        out.update({"achieved": out["hbm"]["achieved"], "peak": hbm_peak, "unit": "GB/s", "frac": hbm_us / us})
    return out


def c1(lg, dev, seed, hbm_peak, peak_src, batch=4096):
    from oracle import ring_oracle as orc

    ring = lg.ring
    N, Q, B = 1 << 13, QI60_TAIL, batch
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    ctx = ring.NewContextWithParams(N, Q)
    a_t, b_t = uniform(dev, (B,), Q, N, g), uniform(dev, (B,), Q, N, g)
    o_t = torch.empty_like(a_t)
    W = lambda t: ring.Poly.wrap(t.data_ptr(), N, 4, B, keep=t)
    a, b, o = W(a_t), W(b_t), W(o_t)
    extra = {"InvNTT": lambda: ctx.InvNTT(a, o, stream=_sp()),
             "MulCoeffsMontgomery": lambda: ctx.MulCoeffsMontgomery(a, b, o, stream=_sp())}

    def cpu(nthreads):
        oc = orc.Context(N, Q)
        x = a_t[0].cpu().numpy().astype(np.uint64)
        v, dt = cpu_parallel(lambda: (lambda: oc.ntt(x)), 2000 * nthreads // 4 + 2000, nthreads)
        return v, "%d NTTs of one 4-limb polynomial, %.1f s wall" % (2000 * nthreads // 4 + 2000, dt)

    def roof(us):
        return _roofline("ntt_fwd (strided + contiguous phase), N = 2^13, 60-bit limbs", int_floor_us({q: B for q in Q}, N),
                         16.0 * N * 4 * B, us, hbm_peak, peak_src, "one launch pair over 4 x %d limb-NTTs" % B)

    return dict(workload="ring.Context NTT, N=2^13, 4 x 60-bit NTT-friendly moduli (ring_benchmark_test.go shapes), batch %d" % B,
                metric="ring.Context NTT polys/s at N=2^13 x 4 limbs", unit="polys/s", units_per_step=B,
                step=lambda: ctx.NTT(a, o, stream=_sp()), extra=extra, extra_unit="polys/s",
                e2e=([(a_t, a_t.numel())], [(o_t, o_t.numel())]), cpu=cpu, roofline=roof, dtype="u64")


def _ckks_modmuls(N, logN, nl, alpha):
    beta = -(-nl // alpha)
    xal = [min(alpha, nl - i * alpha) for i in range(beta)]
    ntts = (beta + 2) * (nl + alpha) + 2 * nl
    mm = (sum(x * (1 + nl - x + alpha) for x in xal) + 2 * beta * (nl + alpha) + 2 * alpha * (1 + nl) + 2 * nl + 6 * nl + 2 * (nl - 1))
    return ntts, ntts * (N // 2) * logN + mm * N, beta


def c2(lg, dev, seed, hbm_peak, peak_src, batch=1024):
    from oracle import ring_oracle as orc

    ring, ckks = lg.ring, lg.ckks
    p = ckks.DefaultParams[ckks.PN14QP438]
    N = 1 << p["LogN"]
    Q, P = ckks.GenModuli(p)
    nQ, nP = len(Q), len(P)
    beta = -(-nQ // nP)
    B = batch
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    cQ, cP = ring.NewContextWithParams(N, Q), ring.NewContextWithParams(N, P)
    ev = ckks.NewEvaluator(cQ, cP)
    evk_t = uniform(dev, (beta, 2), Q + P, N, g)
    rlk = ckks.SwitchingKey(N=N, device_ptr=evk_t.data_ptr(), beta=beta, nQP=nQ + nP, keep=evk_t)
    a_t = [uniform(dev, (B,), Q, N, g) for _ in range(2)]
    b_t = [uniform(dev, (B,), Q, N, g) for _ in range(2)]
    o_t = [torch.empty(B, nQ, N, dtype=torch.int64, device=dev) for _ in range(2)]
    W = lambda t: ring.Poly.wrap(t.data_ptr(), N, nQ, B, keep=t)
    a, b, o = tuple(W(t) for t in a_t), tuple(W(t) for t in b_t), tuple(W(t) for t in o_t)
    level = nQ - 1

    def step():
        ev.MulRelin(level, a, b, rlk, o, stream=_sp())
        ev.Rescale(nQ, o, 1, stream=_sp())

    def cpu(nthreads):
        oQ, oP = orc.Context(N, Q), orc.Context(N, P)
        evk = evk_t.cpu().numpy().astype(np.uint64)
        x = np.ascontiguousarray(np.stack([a_t[0][0].cpu().numpy(), a_t[1][0].cpu().numpy()]).astype(np.uint64))
        y = np.ascontiguousarray(np.stack([b_t[0][0].cpu().numpy(), b_t[1][0].cpu().numpy()]).astype(np.uint64))

        def mk():
            e = orc.CkksEvaluator(oQ, oP)
            return lambda: e.rescale(e.mul_relin(level, x, y, evk))

        v, dt = cpu_parallel(mk, 8 * nthreads, nthreads)
        return v, "%d MulRelin+Rescale ops, one oracle evaluator per host thread, %.1f s wall" % (8 * nthreads, dt)

    def roof(us):
        ntts, _, _ = _ckks_modmuls(N, p["LogN"], nQ, nP)
        per_mod = {}
        for q in Q + P:  # the limb-NTTs of one op spread evenly over the active moduli: a fair mix for the floor
            per_mod[q] = per_mod.get(q, 0) + B * ntts / (nQ + nP)
        bytes_ = B * 8.0 * N * (4 * nQ + 2 * (nQ - 1)) + 8.0 * N * 2 * beta * (nQ + nP)
        return _roofline("MulRelin+Rescale (tensor, key switch, ModDown, rescale), SURVEY.md 8(d) counts",
                         int_floor_us(per_mod, N), bytes_, us, hbm_peak, peak_src,
                         "%d limb-NTTs per op; compulsory bytes = operands + results + the key once per batch" % ntts)

    return dict(workload="CKKS PN14QP438 (N=2^14, 10+2 limbs, level 9): MulRelin+Rescale, batch of %d ciphertexts" % B,
                metric="CKKS MulRelin+Rescale ops/s at logN=14 (batched)", unit="ops/s", units_per_step=B, step=step,
                e2e=([(t, t.numel()) for t in a_t + b_t], [(t, B * (nQ - 1) * N) for t in o_t]), cpu=cpu, roofline=roof, dtype="u64")


def c3(lg, dev, seed, hbm_peak, peak_src, batch=64):
    from oracle import ring_oracle as orc

    ring, bfv, ckks = lg.ring, lg.bfv, lg.ckks
    p = bfv.DefaultParams[bfv.PN15QP880]
    N = 1 << p["LogN"]
    Q, P, QMul = bfv.GenModuli(p)
    nQ, nP, nM = len(Q), len(P), len(QMul)
    beta = -(-nQ // nP)
    B = batch
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    cQ, cM, cP = (ring.NewContextWithParams(N, m) for m in (Q, QMul, P))
    ev = bfv.NewEvaluator(cQ, cM, cP, p["T"])
    evk_t = uniform(dev, (beta, 2), Q + P, N, g)
    key = ckks.SwitchingKey(N=N, device_ptr=evk_t.data_ptr(), beta=beta, nQP=nQ + nP, keep=evk_t)
    a_t = [uniform(dev, (B,), Q, N, g) for _ in range(2)]
    b_t = [uniform(dev, (B,), Q, N, g) for _ in range(2)]
    mk_out = lambda n: [torch.empty(B, nQ, N, dtype=torch.int64, device=dev) for _ in range(n)]
    d2_t, d1_t, r_t = mk_out(3), mk_out(2), mk_out(2)
    W = lambda t: ring.Poly.wrap(t.data_ptr(), N, nQ, B, keep=t)
    a, b, d2, d1, r = (tuple(W(t) for t in ts) for ts in (a_t, b_t, d2_t, d1_t, r_t))
    gen = pow(bfv.GaloisGen, 1, 2 * N)

    def step():
        ev.Mul(a, b, d2, stream=_sp())
        ev.Relinearize(d2, key, d1, stream=_sp())
        ev.permute(d1, gen, key, r, stream=_sp())

    def cpu(nthreads):
        ctxs = (orc.Context(N, Q), orc.Context(N, QMul), orc.Context(N, P))
        evk = evk_t.cpu().numpy().astype(np.uint64)
        x = np.ascontiguousarray(np.stack([a_t[0][0].cpu().numpy(), a_t[1][0].cpu().numpy()]).astype(np.uint64))
        y = np.ascontiguousarray(np.stack([b_t[0][0].cpu().numpy(), b_t[1][0].cpu().numpy()]).astype(np.uint64))

        def mk():
            e = orc.BfvEvaluator(*ctxs, p["T"])
            return lambda: e.permute(e.relinearize(e.tensor_and_rescale(x, y), evk), gen, evk)

        v, dt = cpu_parallel(mk, 2 * nthreads, nthreads)
        return v, "%d Mul+Relinearize+RotateColumns ops, one oracle evaluator per host thread, %.1f s wall" % (2 * nthreads, dt)

    def roof(us):
        # limb-NTTs per op: Mul = 4 inputs x (nQ + nM) forward + 3 outputs x (nQ + nM) inverse; each of the two key switches
        # = nQ forward (c2) + beta*(nQ+nP) - nQ digit limbs + 2*(nQ+nP) inverse (bfv/evaluator.go:278-464, :736-813)
        mul = 7 * (nQ + nM)
        ks = nQ + beta * (nQ + nP) - nQ + 2 * (nQ + nP)
        ntts = mul + 2 * ks
        per_mod = {}
        for q in Q + P + QMul:
            per_mod[q] = per_mod.get(q, 0) + B * ntts / (nQ + nP + nM)
        bytes_ = B * 8.0 * N * nQ * (4 + 3 + 3 + 2 + 2 + 2) + 8.0 * N * 2 * beta * (nQ + nP)
        return _roofline("BFV Mul + Relinearize + RotateColumns", int_floor_us(per_mod, N), bytes_, us, hbm_peak, peak_src,
                         "%d limb-NTTs per op (58..61-bit limbs: the [0,8q) butterflies); bytes = operands and results of "
                         "the three calls + the key once per batch" % ntts)

    return dict(workload="BFV PN15QP880 (N=2^15, 12+3 limbs, QMul 12): Mul + Relinearize + RotateColumns(1), batch %d" % B,
                metric="BFV Mul+Relinearize+RotateColumns ops/s at logN=15 (batched)", unit="ops/s", units_per_step=B, step=step,
                e2e=([(t, t.numel()) for t in a_t + b_t], [(t, t.numel()) for t in r_t]), cpu=cpu, roofline=roof, dtype="u64")


def c5(lg, dev, seed, hbm_peak, peak_src, parties=8):
    from oracle import ring_oracle as orc

    ring, ckks, dckks = lg.ring, lg.ckks, lg.dckks
    p = ckks.DefaultParams[ckks.PN15QP880]
    N = 1 << p["LogN"]
    Q, P = ckks.GenModuli(p)
    QP = Q + P
    nQ, nK = len(Q), len(QP)
    B = parties
    g = torch.Generator(device=dev)
    g.manual_seed(seed)
    cQ, cP, cK = (ring.NewContextWithParams(N, m) for m in (Q, P, QP))
    ckg, pcks = dckks.CKGProtocol(cK), dckks.PCKSProtocol(cQ, cP, cK)
    mk = lambda mods: uniform(dev, (B,), mods, N, g)
    W = lambda t, nl: ring.Poly.wrap(t.data_ptr(), N, nl, B, keep=t)
    t = {k: mk(QP) for k in ("sk", "crs", "e", "ckg", "pk0", "pk1", "u", "e0", "e1")}
    t.update({k: mk(Q) for k in ("ct1", "skq", "s0", "s1")})
    level = nQ - 1

    def step():
        ckg.GenShare(W(t["sk"], nK), W(t["crs"], nK), W(t["ckg"], nK), W(t["e"], nK), stream=_sp())
        pcks.GenShare(level, W(t["skq"], nQ), (W(t["pk0"], nK), W(t["pk1"], nK)), W(t["ct1"], nQ), (W(t["s0"], nQ), W(t["s1"], nQ)),
                      W(t["u"], nK), W(t["e0"], nK), W(t["e1"], nK), stream=_sp())

    def cpu(nthreads):
        oQ, oP, oK = orc.Context(N, Q), orc.Context(N, P), orc.Context(N, QP)
        h = lambda x: np.ascontiguousarray(x[0].cpu().numpy().astype(np.uint64))
        hv = {k: h(v) for k, v in t.items()}

        def mkw():
            ext = orc.Extender(oQ, oP)

            def f():
                w = oK.ntt(hv["e"])
                oK.op3("mulcoeffs_montgomery_and_sub", hv["sk"], hv["crs"], w)
                tt = oK.ntt(hv["u"])
                s0 = oK.op3("add", oK.op3("mulcoeffs_montgomery", tt, hv["pk0"]), oK.ntt(hv["e0"]))
                s1 = oK.op3("add", oK.op3("mulcoeffs_montgomery", tt, hv["pk1"]), oK.ntt(hv["e1"]))
                w0 = ext.moddown_ntt_pq(level, s0)
                ext.moddown_ntt_pq(level, s1)
                oQ.op3("mulcoeffs_montgomery_and_add", hv["ct1"], hv["skq"], w0)
            return f

        v, dt = cpu_parallel(mkw, 2 * nthreads, nthreads)
        return v, "%d parties' CKG+PCKS GenShare, one oracle worker per host thread, %.1f s wall" % (2 * nthreads, dt)

    def roof(us):
        # per party: CKG 1 NTT over QP; PCKS 3 NTTs over QP + two ModDownNTTPQ (each: nP inverse + nQ forward)
        ntts = 4 * nK + 2 * (len(P) + nQ)
        per_mod = {}
        for q in QP:
            per_mod[q] = per_mod.get(q, 0) + B * ntts / nK
        bytes_ = B * 8.0 * N * (9 * nK + 4 * nQ)
        return _roofline("dckks CKG.GenShare + PCKS.GenShare", int_floor_us(per_mod, N), bytes_, us, hbm_peak, peak_src,
                         "%d limb-NTTs per party; bytes = every operand and share once" % ntts)

    ins = [(t[k], t[k].numel()) for k in ("sk", "e", "u", "e0", "e1", "skq")]
    outs = [(t[k], t[k].numel()) for k in ("ckg", "s0", "s1")]
    return dict(workload="dckks over CKKS PN15QP880 (N=2^15, 18+3 limbs): CKG.GenShare + PCKS.GenShare, %d parties side by side" % B,
                metric="dckks CKG+PCKS GenShare party-rounds/s at logN=15", unit="party-rounds/s", units_per_step=B, step=step,
                e2e=(ins, outs), cpu=cpu, roofline=roof, dtype="u64")


CONFIGS = {"C1": (c1, 1), "C2": (c2, 2), "C3": (c3, 3), "C5": (c5, 5)}
