#!/usr/bin/env python
"""bench.py -- headline benchmark of the ring hot path on B200.

Metric (BASELINE.json): batched NTT/s and CKKS MulRelin+Rescale ops/s at logN=16.
Workload at every N: CKKS PN16QP1761 (ckks/params.go:79-86: N=2^16, 34 Q limbs,
4 P limbs, alpha=4, beta=9), level 33 -> 32; one "step" = MulRelin (with a
relinearisation key) followed by Rescale on a batch of independent synthetic
ciphertexts (ckks/evaluator.go:1016 and :933).  `value` is MulRelin+Rescale ops
per second over all ranks with inputs resident in HBM; `ntt` carries the batched
limb-NTT rates; `e2e` is the same step through the C ABI with HOST buffers
(pinned upload of both operands, download of the result, inside the timed region).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU); ciphertexts are independent, so
ranks shard the batch with no data-path collective ("weak" scaling: B per GPU).
--impl reference times the CPU restatement of the reference (oracle/, no Go
toolchain exists in the image) with one evaluator per host thread.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "lattigo-fhe-by-go_b200"))

SEED = 0x1A771C0 + 4  # fixed seed + config id
PARAMS_ID = 4  # ckks.PN16QP1761
WORKLOAD = "CKKS PN16QP1761 (N=2^16, 34+4 limbs, level 33): MulRelin+Rescale, batch of independent ciphertexts"


def workload(params_id):
    """config.workload: the headline set by default; --params runs another ckks.DefaultParams entry and says so"""
    if params_id == PARAMS_ID:
        return WORKLOAD
    return "CKKS DefaultParams[%d] at the top level: MulRelin+Rescale, batch of independent ciphertexts (not the headline set)" % params_id


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=32, help="ciphertexts per GPU per step")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=0, help="ops in the cpu_baseline sample (0 = one per host thread)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-rotate", action="store_true")
    ap.add_argument("--hoisted-rotations", type=int, default=8, help="rotations per RotateHoisted call")
    ap.add_argument("--e2e-chunks", type=int, default=16, help="chunks the batch is cut into for the pipelined e2e path")
    ap.add_argument("--params", type=int, default=PARAMS_ID, help="index into ckks.DefaultParams (default PN16QP1761)")
    ap.add_argument("--config", default="C4", choices=["C1", "C2", "C3", "C4", "C5"],
                    help="BASELINE.json configuration (default C4 = the headline; the others print the same JSON contract)")
    ap.add_argument("--no-legs", action="store_true", help="skip the limb-axis / party-axis legs of a multi-GPU run")
    ap.add_argument("--no-parity-check", action="store_true", help="skip the oracle compare of one timed batch entry")
    return ap.parse_args()


# ----------------------------------------------------------------------------
# CPU arm: the oracle (a literal C restatement of the reference) on host threads,
# one evaluator per thread as in examples/dbfv/psi/psi.go:214-233
# ----------------------------------------------------------------------------
def cpu_mulrelin_rescale(params_id, nops, nthreads, seed=SEED):
    import numpy as np

    from lattigpu import ckks as gckks
    from oracle import ring_oracle as orc

    p = gckks.DefaultParams[params_id]
    N = 1 << p["LogN"]
    Q, P, _ = orc.gen_moduli(p["LogN"], p["LogQi"], p["LogPi"])
    nQ, nP = len(Q), len(P)
    beta = -(-nQ // nP)
    rng = np.random.default_rng(seed)
    oQ, oP = orc.Context(N, Q), orc.Context(N, P)
    evk = np.ascontiguousarray(
        np.stack([rng.integers(0, q, size=(beta, 2, N), dtype=np.uint64) for q in Q + P], axis=2))
    cts = [np.ascontiguousarray(np.stack([rng.integers(0, q, size=(2, N), dtype=np.uint64) for q in Q], axis=1))
           for _ in range(2)]
    evs = [orc.CkksEvaluator(oQ, oP) for _ in range(nthreads)]
    level = nQ - 1
    counter = {"next": 0}
    lock = threading.Lock()

    def worker(ev):
        while True:
            with lock:
                i = counter["next"]
                if i >= nops:
                    return
                counter["next"] = i + 1
            out = ev.mul_relin(level, cts[0], cts[1], evk)  # ctypes releases the GIL
            ev.rescale(out)

    t0 = time.perf_counter()
    ths = [threading.Thread(target=worker, args=(ev,)) for ev in evs]
    for t in ths:
        t.start()
    for t in ths:
        t.join()
    dt = time.perf_counter() - t0
    return nops / dt, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    nops = args.cpu_sample or 4 * cores
    vals = []
    for _ in range(max(args.warmup, 0) and 1):
        cpu_mulrelin_rescale(args.params, min(nops, cores), cores)
    t_all = time.perf_counter()
    for _ in range(args.steps):
        v, _ = cpu_mulrelin_rescale(args.params, nops, cores)
        vals.append(v)
        if time.perf_counter() - t_all > 240:
            break
    vals.sort()
    value = vals[len(vals) // 2]
    sample = "%d MulRelin+Rescale ops per step (one oracle evaluator per host thread), median of %d steps" % (nops, len(vals))
    line = {
        "impl": "reference", "metric": "CKKS MulRelin+Rescale ops/s at logN=16 (batched)", "value": value,
        "unit": "ops/s", "n_gpus": args.gpus, "steps": len(vals), "warmup": 1 if args.warmup > 0 else 0,
        "ms_per_step": 1e3 * nops / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": {"workload": workload(args.params), "batch_per_step": nops},
        "cpu_baseline": {"value": value, "unit": "ops/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "ops/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)



# ----------------------------------------------------------------------------
# Multi-GPU legs (world > 1): the two shard axes that have a real exchange step.
# Both check their result bit for bit inside the leg and abort the run on a mismatch.
# ----------------------------------------------------------------------------
def leg_limb_sharded(args, dev, world, rank, timed_ms):
    """BASELINE config 4, limb axis: ONE CKKS PN16QP1761 ciphertext pair (and a batch of 8), MulRelin + Rescale with
    the RNS limbs spread over the ranks, against the same call on this rank's GPU alone.  The sharded result is
    compared with the single-GPU result on every rank (ckks/evaluator.go:1016-1133, :933-968, :1475-1558)."""
    import torch
    import torch.distributed as dist

    import lattigpu
    from lattigpu import ckks as gckks
    from lattigpu import ring as gring

    p = gckks.DefaultParams[PARAMS_ID]
    N = 1 << p["LogN"]
    Q, P = gckks.GenModuli(p)
    nQ, nP = len(Q), len(P)
    beta = -(-nQ // nP)
    level = nQ - 1
    g = torch.Generator(device=dev)
    g.manual_seed(SEED + 1000)  # the SAME ciphertext on every rank
    sp = torch.cuda.current_stream().cuda_stream

    def uniform(prefix, moduli):
        t = torch.empty(*prefix, len(moduli), N, dtype=torch.int64, device=dev)
        for i, q in enumerate(moduli):
            t[..., i, :] = torch.randint(0, q, (*prefix, N), dtype=torch.int64, device=dev, generator=g)
        return t

    cQ, cP = gring.NewContextWithParams(N, Q), gring.NewContextWithParams(N, P)
    ev = gckks.NewEvaluator(cQ, cP)
    comm = lattigpu.dist.Comm(nccl=False)  # the limb axis exchanges through peer memory: no NCCL communicator
    comm.reserve(ev, 8)                    # exchange buffers for the largest batch of the leg, IPC-mapped by every peer
    evk_t = uniform((beta, 2), Q + P)
    rlk = gckks.SwitchingKey(N=N, device_ptr=evk_t.data_ptr(), beta=beta, nQP=nQ + nP, keep=evk_t)
    out = {"what": "CKKS PN16QP1761 MulRelin+Rescale at level 33, limbs of each ciphertext spread over %d GPUs" % world,
           "exchange": comm.exchange_description(), "cases": []}
    ok_all = True
    for batch in (1, 8):
        wrap = lambda t: gring.Poly.wrap(t.data_ptr(), N, nQ, batch, keep=t)
        a_t = [uniform((batch,), Q) for _ in range(2)]
        b_t = [uniform((batch,), Q) for _ in range(2)]
        o_t = [torch.zeros(batch, nQ, N, dtype=torch.int64, device=dev) for _ in range(2)]
        s_t = [torch.zeros(batch, nQ, N, dtype=torch.int64, device=dev) for _ in range(2)]
        a, b = tuple(wrap(t) for t in a_t), tuple(wrap(t) for t in b_t)
        o, so = tuple(wrap(t) for t in o_t), tuple(wrap(t) for t in s_t)

        def local():
            ev.MulRelin(level, a, b, rlk, o, stream=sp)
            ev.Rescale(nQ, o, 1, stream=sp)

        def sharded():
            comm.MulRelinRescale(ev, level, a, b, rlk, so, stream=sp)

        t_local = timed_ms(local, 20, 5)
        t_shard = timed_ms(sharded, 20, 5)
        comm.check(stream=sp)  # raises if a cross-rank barrier timed out
        # parity: gather the limb-resident result (outside the timed region) and compare with the single-GPU words
        comm.GatherLimbs(ev, nQ - 1, so, stream=sp)
        torch.cuda.synchronize()
        same = all(bool(torch.equal(x[:, :nQ - 1], y[:, :nQ - 1])) for x, y in zip(o_t, s_t))
        flag = torch.tensor([1 if same else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        same = bool(flag.item())
        ok_all &= same
        out["cases"].append({"batch": batch, "ms_per_op_one_gpu": t_local / batch, "ms_per_op_sharded": t_shard / batch,
                             "speedup": t_local / t_shard, "parity": same,
                             "nvlink_bytes_per_op_per_gpu": comm.exchange_bytes(ev, level, 1)})
    out["parity"] = ok_all
    if not ok_all:
        raise SystemExit("bench.py: limb-sharded MulRelin+Rescale differs from the single-GPU result: %r" % (out,))
    return out


def leg_party(args, dev, world, rank, timed_ms):
    """BASELINE config 5, party axis: dckks over CKKS PN15QP880 (N=2^15, 18+3 limbs), 8 parties spread over the ranks.
    Every rank runs CKG.GenShare (dckks/publickey_gen.go:39-42) and PCKS.GenShare (dckks/public_keyswitching.go:63-96)
    for its parties, adds its own shares, and the ranks aggregate with one all-reduce + Reduce
    (lg_comm_aggregate_shares).  Checked against the reference's sequential chain of AggregateShares
    (publickey_gen.go:45-47, public_keyswitching.go:99-103) over all 8 shares, recomputed on every rank."""
    import torch
    import torch.distributed as dist

    import lattigpu
    from lattigpu import ckks as gckks
    from lattigpu import dckks as gdckks
    from lattigpu import ring as gring

    parties = 8
    if parties % world:
        return {"skipped": "8 parties do not divide over %d ranks" % world}
    mine = parties // world
    p = gckks.DefaultParams[gckks.PN15QP880]
    N = 1 << p["LogN"]
    Q, P = gckks.GenModuli(p)
    QP = Q + P
    nQ, nK = len(Q), len(QP)
    level = nQ - 1
    sp = torch.cuda.current_stream().cuda_stream
    cQ, cP, cK = (gring.NewContextWithParams(N, m) for m in (Q, P, QP))
    ckg, pcks = gdckks.CKGProtocol(cK), gdckks.PCKSProtocol(cQ, cP, cK)
    comm = lattigpu.dist.Comm()

    def party_inputs(idx, mods, salt):
        """[len(idx)][len(mods)][N] uniform words, party i seeded by its index (so every rank can rebuild any party)"""
        t = torch.empty(len(idx), len(mods), N, dtype=torch.int64, device=dev)
        for k, i in enumerate(idx):
            g = torch.Generator(device=dev)
            g.manual_seed(SEED + 5000 + 64 * i + salt)
            for j, q in enumerate(mods):
                t[k, j] = torch.randint(0, q, (N,), dtype=torch.int64, device=dev, generator=g)
        return t

    def shared_inputs(mods, salt):
        g = torch.Generator(device=dev)
        g.manual_seed(SEED + 7000 + salt)
        t = torch.empty(1, len(mods), N, dtype=torch.int64, device=dev)
        for j, q in enumerate(mods):
            t[0, j] = torch.randint(0, q, (N,), dtype=torch.int64, device=dev, generator=g)
        return t

    W = lambda t, nl: gring.Poly.wrap(t.data_ptr(), N, nl, t.shape[0], keep=t)
    crs_1, pk0_1, pk1_1, ct1_1 = shared_inputs(QP, 1), shared_inputs(QP, 2), shared_inputs(QP, 3), shared_inputs(Q, 4)

    def build(idx):
        n = len(idx)
        t = {"sk": party_inputs(idx, QP, 0), "e": party_inputs(idx, QP, 1), "u": party_inputs(idx, QP, 2),
             "e0": party_inputs(idx, QP, 3), "e1": party_inputs(idx, QP, 4), "skq": party_inputs(idx, Q, 5),
             "crs": crs_1.expand(n, -1, -1).contiguous(), "pk0": pk0_1.expand(n, -1, -1).contiguous(),
             "pk1": pk1_1.expand(n, -1, -1).contiguous(), "ct1": ct1_1.expand(n, -1, -1).contiguous(),
             "ckg": torch.zeros(n, nK, N, dtype=torch.int64, device=dev),
             "s0": torch.zeros(n, nQ, N, dtype=torch.int64, device=dev), "s1": torch.zeros(n, nQ, N, dtype=torch.int64, device=dev)}
        return t

    def gen_shares(t):
        ckg.GenShare(W(t["sk"], nK), W(t["crs"], nK), W(t["ckg"], nK), W(t["e"], nK), stream=sp)
        pcks.GenShare(level, W(t["skq"], nQ), (W(t["pk0"], nK), W(t["pk1"], nK)), W(t["ct1"], nQ), (W(t["s0"], nQ), W(t["s1"], nQ)),
                      W(t["u"], nK), W(t["e0"], nK), W(t["e1"], nK), stream=sp)

    def chain(ctx, t, nl):
        """the reference's AggregateShares chain over the batch entries of t, into entry 0"""
        acc = W(t[0:1], nl)
        for k in range(1, t.shape[0]):
            ctx.Add(acc, W(t[k:k + 1], nl), acc, stream=sp)
        return t[0:1]

    my_idx = list(range(rank * mine, (rank + 1) * mine))
    tm = build(my_idx)

    def round_():
        gen_shares(tm)
        for key, ctx, nl in (("ckg", cK, nK), ("s0", cQ, nQ), ("s1", cQ, nQ)):
            chain(ctx, tm[key], nl)
            comm.AggregateShares(ctx, W(tm[key][0:1], nl), stream=sp)

    t_round = timed_ms(round_, 10, 3)
    # parity: all 8 parties on this rank, sequential chain in party order
    ta = build(list(range(parties)))
    gen_shares(ta)
    round_()  # leaves the aggregates in entry 0 of tm[*]
    ok = True
    for key, ctx, nl in (("ckg", cK, nK), ("s0", cQ, nQ), ("s1", cQ, nQ)):
        want = chain(ctx, ta[key], nl)
        torch.cuda.synchronize()
        ok &= bool(torch.equal(want, tm[key][0:1]))
    flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    ok = bool(flag.item())
    res = {"what": "dckks PN15QP880: CKG.GenShare + PCKS.GenShare for 8 parties (%d per GPU), AggregateShares over %d GPUs" % (mine, world),
           "collective": "ncclAllReduce(sum, uint64) + Reduce kernel, 3 polys per round", "parties": parties,
           "allreduce_bytes_per_round": (nK + 2 * nQ) * N * 8, "ms_per_round": t_round,
           "party_shares_per_s": parties * 1e3 / t_round, "parity": ok}
    if not ok:
        raise SystemExit("bench.py: aggregated shares differ from the sequential AggregateShares chain: %r" % (res,))
    return res

# ----------------------------------------------------------------------------
# stdout carries exactly ONE JSON line.  Libraries write there too (NCCL prints its version banner on stdout under
# NCCL_DEBUG=VERSION whatever NCCL_DEBUG_FILE says), so the process' fd 1 is pointed at stderr for the whole run and
# the JSON line goes to a duplicate of the original stdout.
# ----------------------------------------------------------------------------
_REAL_STDOUT = None


def capture_stdout():
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


# ----------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------
class ClockSampler:
    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + self.QUERY, "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:
            self.proc.kill()
            out = ""
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in out.strip().splitlines():
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}


def run_gpu(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    import lattigpu
    from lattigpu import ckks as gckks
    from lattigpu import ring as gring

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    gring.set_device(local_rank)
    if world > 1:
        capture_stdout()
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    p = gckks.DefaultParams[args.params]
    N = 1 << p["LogN"]
    Q, P = gckks.GenModuli(p)
    nQ, nP = len(Q), len(P)
    alpha = nP
    beta = -(-nQ // alpha)
    level = nQ - 1
    B = args.batch
    ctxQ, ctxP = gring.NewContextWithParams(N, Q), gring.NewContextWithParams(N, P)
    ev = gckks.NewEvaluator(ctxQ, ctxP)
    st = torch.cuda.current_stream()
    sp = st.cuda_stream

    g = torch.Generator(device=dev)
    g.manual_seed(SEED + rank)

    def uniform(shape_prefix, moduli):
        """uniform in [0, q_i) per limb, generated on the device: [*prefix][limb][N] int64"""
        t = torch.empty(*shape_prefix, len(moduli), N, dtype=torch.int64, device=dev)
        for i, q in enumerate(moduli):
            t[..., i, :] = torch.randint(0, q, (*shape_prefix, N), dtype=torch.int64, device=dev, generator=g)
        return t

    def wrap(t, nl, batch):
        return gring.Poly.wrap(t.data_ptr(), N, nl, batch, keep=t)

    evk_t = uniform((beta, 2), Q + P)
    rlk = gckks.SwitchingKey(N=N, device_ptr=evk_t.data_ptr(), beta=beta, nQP=nQ + nP, keep=evk_t)
    a_t = [uniform((B,), Q) for _ in range(2)]
    b_t = [uniform((B,), Q) for _ in range(2)]
    o_t = [torch.empty(B, nQ, N, dtype=torch.int64, device=dev) for _ in range(2)]
    ct_a = tuple(wrap(t, nQ, B) for t in a_t)
    ct_b = tuple(wrap(t, nQ, B) for t in b_t)
    ct_o = tuple(wrap(t, nQ, B) for t in o_t)

    def step():
        ev.MulRelin(level, ct_a, ct_b, rlk, ct_o, stream=sp)
        ev.Rescale(nQ, ct_o, 1, stream=sp)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = gring.launch_count()
        e0.record(st)
        for _ in range(steps):
            fn()
        e1.record(st)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = gring.launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    def timed_ms(fn, reps, warm):
        ms_, _ = timed(fn, reps, warm)
        return ms_ / reps

    # ---- headline: MulRelin + Rescale, inputs resident in HBM -------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, launches = timed(step, args.steps, args.warmup)
    clocks = sampler.stop() if sampler else None
    ms_per_step = ms / args.steps
    value = world * B * 1e3 / ms_per_step

    # ---- batched limb-NTT rates (the dominant kernel family, timed alone) -------
    ntt_in, ntt_out = ct_a[0], ct_o[0]
    nlimbs_launch = B * nQ

    def fwd():
        ctxQ.NTT(ntt_in, ntt_out, stream=sp)

    def inv():
        ctxQ.InvNTT(ntt_in, ntt_out, stream=sp)

    reps = max(20, args.steps)
    fwd_ms, _ = timed(fwd, reps, 3)
    inv_ms, _ = timed(inv, reps, 3)
    fwd_us = 1e3 * fwd_ms / reps
    inv_us = 1e3 * inv_ms / reps
    ntt_fwd_rate = world * nlimbs_launch / (fwd_us * 1e-6)
    ntt_inv_rate = world * nlimbs_launch / (inv_us * 1e-6)

    # ---- Rotate (BASELINE config 4 names MulRelin + Rescale + Rotate): RotateColumns with a direct key
    # (ckks/evaluator.go:1201-1248) and RotateHoisted over `nrot` rotations (:1252-1392), same batch, resident
    rotate = None
    if not args.no_rotate:
        GaloisGen = 5
        nrot = args.hoisted_rotations
        idxs = [gring.PermuteNTTIndex(GaloisGen, k + 1, N) for k in range(nrot)]
        rk = gckks.SwitchingKey(N=N, device_ptr=evk_t.data_ptr(), beta=beta, nQP=nQ + nP, keep=evk_t)  # any uniform key

        def rot():
            ev.permuteNTT(level, ct_a, idxs[0], rk, ct_o, stream=sp)

        def hoisted():
            ev.RotateHoisted(level, ct_a, [(ix, rk) for ix in idxs], [ct_o] * nrot, stream=sp)  # stream-ordered: one output buffer

        rsteps = max(3, min(args.steps, 10))
        rot_ms, _ = timed(rot, rsteps, 3)
        hst_ms, _ = timed(hoisted, rsteps, 3)
        rotate = {"rotate_columns_ops_per_s": world * B * rsteps * 1e3 / rot_ms,
                  "rotate_hoisted_rotations_per_s": world * B * nrot * rsteps * 1e3 / hst_ms,
                  "hoisted_rotations_per_call": nrot, "batch_per_gpu": B, "level": level}

    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    alg_bytes = 16.0 * N * nlimbs_launch  # one limb-NTT reads and writes N words once
    achieved = alg_bytes / (fwd_us * 1e-6) / 1e9
    butterflies = (N // 2) * p["LogN"] * nlimbs_launch
    # ncu --set full DRAM bytes of the same launch pair (profiles/r01_ncu_ntt_fwd.json, captured per round)
    traffic = None
    for name in ("r02_ncu_ntt_fwd.json", "r01_ncu_ntt_fwd.json", "r01_ncu_ntt_fwd_b16.json"):  # one capture per round / batch size
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                prof = json.load(f)
            if prof.get("limb_ntts_per_launch") == nlimbs_launch and prof.get("N") == N:
                traffic = prof["dram_bytes_per_launch"]
                break
        except Exception:
            pass
    # INT-side ceiling: register-resident butterfly rates measured on this GPU (profiles/microbench/
    # fast_butterfly.cu, profiles/r01_butterfly_peaks.txt): FP64-quotient butterfly for moduli below 3*2^44,
    # the 16-instruction Shoup butterfly below 2^56, the [0,8q) butterfly above
    # (round 2: the limbs below 3*2^44 run the 8-instruction FP64-only butterfly, 2.074e12/s register-resident,
    # profiles/r02_fp64_butterfly.txt -- the ceiling moves up with the shorter instruction sequence)
    peak_bf = {"f64": 2.074e12, "free": 1.151e12, "lazy": 0.950e12}
    mix = {"f64": sum(1 for q in Q if q < (3 << 44)), "free": sum(1 for q in Q if (3 << 44) <= q < (1 << 56)),
           "lazy": sum(1 for q in Q if q >= (1 << 56))}
    bf_per_limb = (N // 2) * p["LogN"]
    int_floor_us = 1e6 * B * sum(mix[k] * bf_per_limb / peak_bf[k] for k in mix)
    # quote the SLOWER bound.  The transform is INT-pipe bound long before it is HBM bound, so the
    # binding roof (bound/achieved/peak/frac) is the register-resident butterfly rate; the HBM side (algorithmic bytes
    # against the measured copy bandwidth, and the DRAM traffic ncu saw) stands beside it.
    bf_rate = butterflies / (fwd_us * 1e-6)
    bf_peak = butterflies / (int_floor_us * 1e-6)
    roofline = {
        "bound": "int", "kernel": "ntt_fwd (TMA-fed strided phase + pipelined contiguous phase, one batched limb-NTT launch pair)",
        "achieved": bf_rate / 1e9, "peak": bf_peak / 1e9, "unit": "Gbutterfly/s", "frac": bf_rate / bf_peak,
        "traffic": traffic, "limb_mix": mix,
        "peak_source": "profiles/r02_fp64_butterfly.txt + profiles/r01_butterfly_peaks.txt (register-resident butterfly "
                       "microbenchmarks on B200 of the instruction sequences shipped, weighted by the limb mix); tensor "
                       "cores unused: exact 64-bit modular integer arithmetic",
        "algorithmic_bytes_per_launch": alg_bytes, "limb_ntts_per_launch": nlimbs_launch, "launch_us": fwd_us,
        "hbm": {"achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak, "peak_source": peak_src,
                "traffic_over_algorithmic": (traffic / alg_bytes) if traffic else None,
                # the two-phase transform passes through HBM twice: on the DRAM bytes ncu saw it runs at this fraction of peak
                "dram_frac": (traffic / (fwd_us * 1e-6) / 1e9 / hbm_peak) if traffic else None},
    }

    # ---- op-level roofline, SURVEY.md 8(d): max(INT work / INT peak, compulsory bytes / HBM peak) over the measured
    # time of one MulRelin+Rescale.  Work counts are the reference algorithm's (every butterfly and every
    # coefficient product = one 64-bit modular multiplication = 11 32x32 multiplies, modular_reduction.go:70-79);
    # INT peak = the measured IMAD issue rate (profiles/r01_int_pipe.txt: 61 thread-instructions/clk/SM) at the
    # clock of this run.  The kernels need fewer multiplies than that per product (6 with the FP64 quotient), so this
    # fraction is an efficiency against the reference's arithmetic, not a hard ceiling; roofline.int_pipe above is
    # the ceiling of the instruction sequences actually issued.
    nl_top = level + 1
    xal = [min(alpha, nl_top - i * alpha) for i in range(beta)]
    ntts_per_op = (beta + 2) * (nl_top + alpha) + 2 * nl_top
    mm_coeff = (sum(x * (1 + nl_top - x + alpha) for x in xal) + 2 * beta * (nl_top + alpha)
                + 2 * alpha * (1 + nl_top) + 2 * nl_top + 6 * nl_top + 2 * (nl_top - 1))
    modmuls_per_op = ntts_per_op * (N // 2) * p["LogN"] + mm_coeff * N
    sm_mhz = float((clocks or {}).get("sm_mhz") or peaks.get("sm_max_mhz", 1965.0))
    imad_per_s = 61.0 * 148 * sm_mhz * 1e6
    int_us = 1e6 * modmuls_per_op * 11 / imad_per_s
    bytes_op = 8.0 * N * (4 * nl_top + 2 * (nl_top - 1)) + 8.0 * N * 2 * beta * (nl_top + alpha) / B
    hbm_us = 1e6 * bytes_op / (hbm_peak * 1e9)
    us_per_op = 1e3 * ms_per_step / B
    traffic_step = None  # total DRAM bytes of one step, summed over its launches in the ncu --set full capture
    for name in ("r02_step_traffic.json", "r01_step_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                prof = json.load(f)
            if prof.get("batch") == B and prof.get("N") == N:
                traffic_step = prof["dram_bytes_per_step"]
                break
        except Exception:
            pass
    roofline_op = {
        "what": "one MulRelin+Rescale (SURVEY.md 8(d) counts), key bytes amortised over the batch",
        "traffic_step": traffic_step, "compulsory_bytes_per_step": bytes_op * B,
        "limb_ntts_per_op": ntts_per_op, "modmuls_per_op": modmuls_per_op, "compulsory_bytes_per_op": bytes_op,
        "int_bound_us": int_us, "hbm_bound_us": hbm_us, "measured_us": us_per_op,
        "bound": "int" if int_us >= hbm_us else "hbm", "frac": max(int_us, hbm_us) / us_per_op,
        "int_peak": "61 IMAD/clk/SM x 148 SMs x %.0f MHz (profiles/r01_int_pipe.txt), 11 multiplies per modular product" % sm_mhz,
    }

    # ---- one entry of the timed batch against the oracle (outside the timed region) ------------------
    parity_check = None
    if rank == 0 and not args.no_parity_check:
        from oracle import ring_oracle as orc

        step()
        torch.cuda.synchronize()
        i = B - 1
        ha = np.ascontiguousarray(np.stack([a_t[0][i].cpu().numpy(), a_t[1][i].cpu().numpy()]).astype(np.uint64))
        hb = np.ascontiguousarray(np.stack([b_t[0][i].cpu().numpy(), b_t[1][i].cpu().numpy()]).astype(np.uint64))
        hk = np.ascontiguousarray(evk_t.cpu().numpy().astype(np.uint64))
        oev = orc.CkksEvaluator(orc.Context(N, Q), orc.Context(N, P))
        t0 = time.perf_counter()
        want = oev.rescale(oev.mul_relin(level, ha, hb, hk))
        got = np.stack([o_t[0][i, :nQ - 1].cpu().numpy(), o_t[1][i, :nQ - 1].cpu().numpy()]).astype(np.uint64)
        equal = bool(np.array_equal(got, want))
        parity_check = {"entry": i, "of_batch": B, "against": "oracle/ring_oracle.c (CPU restatement)", "equal": equal,
                        "oracle_s": time.perf_counter() - t0}
        if not equal:
            raise SystemExit("bench.py: entry %d of the timed batch differs from the oracle" % i)

    # ---- multi-GPU legs: limb axis (config 4) and party axis (config 5) -----------------------------
    legs = {"limb_sharded": None, "party": None}
    if world > 1 and not args.no_legs:
        legs["limb_sharded"] = leg_limb_sharded(args, dev, world, rank, timed_ms)
        torch.cuda.empty_cache()
        legs["party"] = leg_party(args, dev, world, rank, timed_ms)
        torch.cuda.empty_cache()

    # ---- e2e: host buffers through the C ABI ------------------------------------
    e2e = None
    if not args.no_e2e:
        h_a = [torch.empty(B, nQ, N, dtype=torch.int64).pin_memory() for _ in range(2)]
        h_b = [torch.empty(B, nQ, N, dtype=torch.int64).pin_memory() for _ in range(2)]
        h_o = [torch.empty(B, nQ - 1, N, dtype=torch.int64).pin_memory() for _ in range(2)]
        for h, d in zip(h_a + h_b, a_t + b_t):
            h.copy_(d)
        L = lattigpu.lib()
        u64p = ctypes.POINTER(ctypes.c_uint64)

        def hp(t):
            return ctypes.cast(t.data_ptr(), u64p)

        # The batch is cut into chunks that flow through `nstreams` streams: chunk c uploads both operands
        # (pinned -> device, stream-ordered), runs MulRelin + Rescale, and downloads the result, so the
        # copies of one chunk overlap the kernels of another.  All calls go through the C ABI.
        nchunks = max(1, min(args.e2e_chunks, B))
        while B % nchunks:
            nchunks -= 1
        cb = B // nchunks
        nstreams = min(3, nchunks)
        streams = [torch.cuda.Stream(device=dev) for _ in range(nstreams)]
        slots = []
        for _ in range(nstreams):
            ta = [torch.empty(cb, nQ, N, dtype=torch.int64, device=dev) for _ in range(4)]
            to = [torch.empty(cb, nQ, N, dtype=torch.int64, device=dev) for _ in range(2)]
            slots.append((tuple(wrap(t, nQ, cb) for t in ta[:2]), tuple(wrap(t, nQ, cb) for t in ta[2:]),
                          tuple(wrap(t, nQ, cb) for t in to)))

        def e2e_step():
            for c in range(nchunks):
                s_ = streams[c % nstreams]
                sa, sb, so = slots[c % nstreams]
                p_ = ctypes.c_void_p(s_.cuda_stream)
                for poly, h in zip(sa + sb, h_a + h_b):
                    lattigpu._lib.check(L.lg_poly_upload_async(poly.h, 0, cb, 0, nQ, hp(h[c * cb:]), p_))
                ev.MulRelin(level, sa, sb, rlk, so, stream=s_.cuda_stream)
                ev.Rescale(nQ, so, 1, stream=s_.cuda_stream)
                for poly, h in zip(so, h_o):
                    lattigpu._lib.check(L.lg_poly_download_async(poly.h, 0, cb, 0, nQ - 1, hp(h[c * cb:]), p_))
            for s_ in streams:
                s_.synchronize()

        e2e_steps = max(3, min(args.steps, 5))
        for _ in range(2):
            e2e_step()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        # the same bytes with no compute in between: what PCIe alone allows (both directions overlapped)
        def copy_only():
            for c in range(nchunks):
                s_ = streams[c % nstreams]
                sa, sb, so = slots[c % nstreams]
                p_ = ctypes.c_void_p(s_.cuda_stream)
                for poly, h in zip(sa + sb, h_a + h_b):
                    lattigpu._lib.check(L.lg_poly_upload_async(poly.h, 0, cb, 0, nQ, hp(h[c * cb:]), p_))
                for poly, h in zip(so, h_o):
                    lattigpu._lib.check(L.lg_poly_download_async(poly.h, 0, cb, 0, nQ - 1, hp(h[c * cb:]), p_))
            for s_ in streams:
                s_.synchronize()

        copy_only()
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            copy_only()
        barrier()
        copy_ms = 1e3 * (time.perf_counter() - t0) / e2e_steps
        h2d_b, d2h_b = 4 * B * nQ * N * 8, 2 * B * (nQ - 1) * N * 8
        e2e = {"value": world * B * e2e_steps / dt, "unit": "ops/s", "h2d_bytes_per_step": h2d_b,
               "copy_only_ms_per_step": copy_ms,
               # what this rank's host link delivered with all ranks copying at once (both directions overlapped)
               "per_rank_copy_GBps": {"h2d": h2d_b / copy_ms / 1e6, "d2h": d2h_b / copy_ms / 1e6,
                                      "aggregate_all_ranks": world * (h2d_b + d2h_b) / copy_ms / 1e6},
               "limiter": "host link: the step moves %.2f GB per GPU over PCIe; with N ranks copying concurrently the "
                          "aggregate host-memory / root-complex bandwidth of the box (all GPUs report one NUMA node) caps "
                          "the rate, not a kernel or a collective" % ((h2d_b + d2h_b) / 1e9),
               "d2h_bytes_per_step": 2 * B * (nQ - 1) * N * 8, "ms_per_step": 1e3 * dt / e2e_steps,
               "note": "pinned host buffers; per chunk of %d ciphertexts: lg_poly_upload_async x4, MulRelin, Rescale, "
                       "lg_poly_download_async x2; %d chunks over %d streams, host sync per step" % (cb, nchunks, nstreams)}

    # ---- cpu baseline (rank 0, N=1 only) -----------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        nops = args.cpu_sample or 8 * cores  # about 15 s of CPU work on the box
        v, dt = cpu_mulrelin_rescale(args.params, nops, cores)
        cpu = {"value": v, "unit": "ops/s", "cores": cores, "kind": "port",
               "sample": "%d MulRelin+Rescale ops, one oracle evaluator per host thread, %.1f s wall" % (nops, dt)}

    if rank == 0:
        line = {
            "metric": "CKKS MulRelin+Rescale ops/s at logN=16 (batched)", "value": value, "unit": "ops/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
            "config": {"workload": workload(args.params), "batch_per_gpu": B, "global_batch": B * world, "level": level,
                       "l2_policy": "inputs+key larger than L2 (%.0f MiB per step per GPU), no flush" %
                                    ((4 * B * nQ + 2 * beta * (nQ + nP)) * N * 8 / 2**20),
                       "parallelism": "batch-sharded x%d, no collective" % world, "seed": SEED},
            "ntt": {"fwd_limb_ntt_per_s": ntt_fwd_rate, "inv_limb_ntt_per_s": ntt_inv_rate, "N": N,
                    "limbs_per_launch": nlimbs_launch, "fwd_us_per_launch": fwd_us, "inv_us_per_launch": inv_us},
            "rotate": rotate, "roofline": roofline, "roofline_op": roofline_op, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks,
            "parity_check": parity_check, "limb_sharded": legs["limb_sharded"], "party": legs["party"],
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def run_config(args):
    """bench.py --config C1|C2|C3|C5: the same JSON contract on another BASELINE.json configuration (bench_configs.py)"""
    import numpy as np
    import torch
    import torch.distributed as dist

    import bench_configs
    import lattigpu
    from lattigpu import ring as gring

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    gring.set_device(local_rank)
    if world > 1:
        capture_stdout()
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "MEASURED_PEAKS.json (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"
    builder, cid = bench_configs.CONFIGS[args.config]
    cfg = builder(lattigpu, dev, 0x1A771C0 + cid + rank, hbm_peak, peak_src)
    st = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = gring.launch_count()
        e0.record(st)
        for _ in range(steps):
            fn()
        e1.record(st)
        barrier()
        ms = e0.elapsed_time(e1)
        launches = gring.launch_count() - l0
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, launches

    sampler = ClockSampler(local_rank) if rank == 0 else None
    ms, launches = timed(cfg["step"], args.steps, max(args.warmup, 3))
    clocks = sampler.stop() if sampler else None
    ms_per_step = ms / args.steps
    units = cfg["units_per_step"]
    value = world * units * 1e3 / ms_per_step
    extra = {}
    for name, fn in cfg.get("extra", {}).items():
        ems, _ = timed(fn, args.steps, 3)
        extra[name] = {"value": world * units * args.steps * 1e3 / ems, "unit": cfg.get("extra_unit", cfg["unit"])}

    # e2e: inputs from pinned host memory, results back to pinned host memory, inside the timed region
    e2e = None
    if not args.no_e2e:
        ins, outs = cfg["e2e"]
        h_in = [torch.empty(n, dtype=torch.int64).pin_memory() for _, n in ins]
        h_out = [torch.empty(n, dtype=torch.int64).pin_memory() for _, n in outs]
        for h, (t, n) in zip(h_in, ins):
            h.copy_(t.reshape(-1)[:n])

        def e2e_step():
            for h, (t, n) in zip(h_in, ins):
                t.reshape(-1)[:n].copy_(h, non_blocking=True)
            cfg["step"]()
            for h, (t, n) in zip(h_out, outs):
                h.copy_(t.reshape(-1)[:n], non_blocking=True)
            torch.cuda.current_stream().synchronize()

        for _ in range(2):
            e2e_step()
        barrier()
        nsteps = max(3, min(args.steps, 5))
        t0 = time.perf_counter()
        for _ in range(nsteps):
            e2e_step()
        barrier()
        dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * units * nsteps / dt, "unit": cfg["unit"], "h2d_bytes_per_step": 8 * sum(n for _, n in ins),
               "d2h_bytes_per_step": 8 * sum(n for _, n in outs), "ms_per_step": 1e3 * dt / nsteps,
               "note": "pinned host buffers -> device, the step through the C ABI, results -> pinned host buffers; host sync per step"}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        v, sample = cfg["cpu"](cores)
        cpu = {"value": v, "unit": cfg["unit"], "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": cfg["metric"], "value": value, "unit": cfg["unit"], "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": cfg["dtype"], "data": "synthetic",
            "config": {"workload": cfg["workload"], "baseline_config": args.config, "units_per_gpu_per_step": units,
                       "l2_policy": "inputs larger than L2, no flush", "parallelism": "batch-sharded x%d, no collective" % world},
            "also": extra, "roofline": cfg["roofline"](1e3 * ms_per_step), "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": launches, "clocks": clocks,
        }
        emit(line)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    elif args.config != "C4":
        run_config(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
