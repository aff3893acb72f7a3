"""Sweep of the "ntt_l2_bytes" switch: forward / inverse batched limb-NTT time (N = 2^16, 34 limbs x 32 ciphertexts) and
the MulRelin+Rescale step for several L2 budgets of the two-phase transform (0 = one launch pair over the whole batch)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "lattigo-fhe-by-go_b200"))
import torch

import lattigpu
from lattigpu import ckks, ring

ring.set_device(0)
dev = torch.device("cuda", 0)
p = ckks.DefaultParams[ckks.PN16QP1761]
N = 1 << p["LogN"]
Q, P = ckks.GenModuli(p)
nQ, nP = len(Q), len(P)
beta = -(-nQ // nP)
B = 32
g = torch.Generator(device=dev)
g.manual_seed(1)


def uniform(prefix, moduli):
    t = torch.empty(*prefix, len(moduli), N, dtype=torch.int64, device=dev)
    for i, q in enumerate(moduli):
        t[..., i, :] = torch.randint(0, q, (*prefix, N), dtype=torch.int64, device=dev, generator=g)
    return t


W = lambda t: ring.Poly.wrap(t.data_ptr(), N, nQ, B, keep=t)
cQ, cP = ring.NewContextWithParams(N, Q), ring.NewContextWithParams(N, P)
ev = ckks.NewEvaluator(cQ, cP)
evk_t = uniform((beta, 2), Q + P)
rlk = ckks.SwitchingKey(N=N, device_ptr=evk_t.data_ptr(), beta=beta, nQP=nQ + nP, keep=evk_t)
a = tuple(W(uniform((B,), Q)) for _ in range(2))
b = tuple(W(uniform((B,), Q)) for _ in range(2))
o = tuple(W(torch.empty(B, nQ, N, dtype=torch.int64, device=dev)) for _ in range(2))
sp = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def step():
    ev.MulRelin(nQ - 1, a, b, rlk, o, stream=sp)
    ev.Rescale(nQ, o, 1, stream=sp)


if len(sys.argv) > 1 and sys.argv[1] == "time":  # one library (LATTIGPU_LIB): median / min step and transform times
    import statistics

    n = int(sys.argv[2]) if len(sys.argv) > 2 else 12
    st = [timed(step, reps=5, warm=1) for _ in range(n)]
    fw = [timed(lambda: cQ.NTT(a[0], o[0], stream=sp), reps=5, warm=1) for _ in range(n)]
    iv = [timed(lambda: cQ.InvNTT(a[0], o[0], stream=sp), reps=5, warm=1) for _ in range(n)]
    print(json.dumps({"lib": os.environ.get("LATTIGPU_LIB", "default"), "step_us": statistics.median(st), "step_min_us": min(st),
                      "fwd_us": statistics.median(fw), "inv_us": statistics.median(iv)}), flush=True)
    sys.exit(0)
if len(sys.argv) > 2 and sys.argv[1] == "ab":  # interleaved A/B of 0/1 switches: python ntt_l2_sweep.py ab <switch> [<switch> ...]
    import statistics

    for _ in range(3):
        step()
    for spec in sys.argv[2:]:  # <switch> (values 0 and 1) or <switch>:<a>:<b>
        name, *vals = spec.split(":")
        va, vb = (int(vals[0]), int(vals[1])) if vals else (0, 1)
        acc = {va: {"fwd": [], "inv": [], "step": []}, vb: {"fwd": [], "inv": [], "step": []}}
        rounds = int(os.environ.get("AB_ROUNDS", "8"))
        step_only = os.environ.get("AB_STEP_ONLY", "0") == "1"
        for rnd in range(rounds):
            for val in ((va, vb) if rnd % 2 == 0 else (vb, va)):
                ring.debug_set_switch(name, val)
                if not step_only:
                    acc[val]["fwd"].append(timed(lambda: cQ.NTT(a[0], o[0], stream=sp), reps=5, warm=1))
                    acc[val]["inv"].append(timed(lambda: cQ.InvNTT(a[0], o[0], stream=sp), reps=5, warm=1))
                acc[val]["step"].append(timed(step, reps=5 if step_only else 3, warm=1))
        ring.debug_set_switch(name, vb if vals else 0)
        diffs = [y / x - 1.0 for x, y in zip(acc[va]["step"], acc[vb]["step"])]  # per round: step(vb) / step(va) - 1
        print(json.dumps({"switch": spec, "rounds": rounds,
                          "median_us": {str(v): {k: statistics.median(x) for k, x in acc[v].items() if x} for v in (va, vb)},
                          "min_us": {str(v): {k: min(x) for k, x in acc[v].items() if x} for v in (va, vb)},
                          "step_rel_median": statistics.median(diffs), "step_rel_quartiles": statistics.quantiles(diffs, n=4)}), flush=True)
    sys.exit(0)
if len(sys.argv) > 1 and sys.argv[1] == "rev":  # A/B of the backward grid walk of the second phases (ABAB)
    for norev in (1, 0, 1, 0):
        ring.debug_set_switch("reverse_walk", 1 - norev)
        print(json.dumps({"no_reverse_walk": norev, "fwd_us": timed(lambda: cQ.NTT(a[0], o[0], stream=sp)),
                          "inv_us": timed(lambda: cQ.InvNTT(a[0], o[0], stream=sp)), "step_us": timed(step, reps=8)}), flush=True)
    sys.exit(0)
for mib in [int(x) for x in sys.argv[1:]] or [0, 16, 24, 32, 48, 64, 96]:
    ring.debug_set_switch("ntt_l2_bytes", mib << 20)
    res = {"ntt_l2_MiB": mib, "fwd_us": timed(lambda: cQ.NTT(a[0], o[0], stream=sp)),
           "inv_us": timed(lambda: cQ.InvNTT(a[0], o[0], stream=sp)), "step_us": timed(step, reps=8)}
    res["fwd_limb_ntt_per_s"] = B * nQ / (res["fwd_us"] * 1e-6)
    print(json.dumps(res), flush=True)
