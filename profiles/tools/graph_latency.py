"""Single-ciphertext latency of CKKS PN16QP1761 MulRelin+Rescale (level 33): eager launches against a CUDA graph of the
same calls captured through the C ABI (torch.cuda.graph on the capture stream; scratch is stream-ordered, so the
allocations become graph nodes).  The replayed result is bit-compared with the eager one."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "lattigo-fhe-by-go_b200"))
import torch

from lattigpu import ckks, ring

ring.set_device(0)
dev = torch.device("cuda", 0)
p = ckks.DefaultParams[ckks.PN16QP1761]
N = 1 << p["LogN"]
Q, P = ckks.GenModuli(p)
nQ, nP = len(Q), len(P)
beta = -(-nQ // nP)
g = torch.Generator(device=dev)
g.manual_seed(1)


def uniform(prefix, moduli):
    t = torch.empty(*prefix, len(moduli), N, dtype=torch.int64, device=dev)
    for i, q in enumerate(moduli):
        t[..., i, :] = torch.randint(0, q, (*prefix, N), dtype=torch.int64, device=dev, generator=g)
    return t


cQ, cP = ring.NewContextWithParams(N, Q), ring.NewContextWithParams(N, P)
ev = ckks.NewEvaluator(cQ, cP)
evk_t = uniform((beta, 2), Q + P)
rlk = ckks.SwitchingKey(N=N, device_ptr=evk_t.data_ptr(), beta=beta, nQP=nQ + nP, keep=evk_t)


def timed(fn, reps=50, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for B in [int(x) for x in sys.argv[1:]] or [1, 2, 8]:
    W = lambda t: ring.Poly.wrap(t.data_ptr(), N, nQ, B, keep=t)
    at, bt = [uniform((B,), Q) for _ in range(2)], [uniform((B,), Q) for _ in range(2)]
    ot = [torch.empty(B, nQ, N, dtype=torch.int64, device=dev) for _ in range(2)]
    a, b, o = tuple(W(t) for t in at), tuple(W(t) for t in bt), tuple(W(t) for t in ot)

    def step(stream):
        ev.MulRelin(nQ - 1, a, b, rlk, o, stream=stream)
        ev.Rescale(nQ, o, 1, stream=stream)

    sp = torch.cuda.current_stream().cuda_stream
    step(sp)
    torch.cuda.synchronize()
    want = [t.clone() for t in ot]
    eager_ms = timed(lambda: step(sp))
    for t in ot:
        t.zero_()
    gr = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step(side.cuda_stream)  # warm-up on the capture stream (pool growth, attribute caches)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    with torch.cuda.graph(gr, stream=side):
        step(torch.cuda.current_stream().cuda_stream)
    for t in ot:
        t.zero_()
    gr.replay()
    torch.cuda.synchronize()
    equal = all(bool(torch.equal(x[:, : nQ - 1], y[:, : nQ - 1])) for x, y in zip(ot, want))
    graph_ms = timed(gr.replay)
    print(json.dumps({"batch": B, "eager_ms_per_call": eager_ms, "graph_ms_per_call": graph_ms, "equal": equal,
                      "eager_ms_per_op": eager_ms / B, "graph_ms_per_op": graph_ms / B}), flush=True)
