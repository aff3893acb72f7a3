"""The C-ABI entry points are CUDA-graph capturable: scratch is stream-ordered (cudaMallocAsync on the caller's stream),
nothing on the launch path synchronises or allocates outside the stream, and a key's derived forms are built on first use
(before the capture).  MulRelin + Rescale + RotateColumns captured once, replayed on new inputs, compared with the oracle."""
import numpy as np
import pytest

from oracle import ring_oracle as orc

pytestmark = pytest.mark.gpu


def test_mulrelin_rescale_rotate_in_a_cuda_graph():
    import torch

    import lattigpu
    from lattigpu import ckks, ring

    ring.set_device(0)
    p = ckks.DefaultParams[ckks.PN13QP218]
    N = 1 << p["LogN"]
    Q, P = ckks.GenModuli(p)
    nQ, nP = len(Q), len(P)
    beta = -(-nQ // nP)
    B = 3
    rng = np.random.default_rng(11)
    evk = np.ascontiguousarray(np.stack([rng.integers(0, q, size=(beta, 2, N), dtype=np.uint64) for q in Q + P], axis=2))
    dev = torch.device("cuda", 0)
    at = [torch.zeros(B, nQ, N, dtype=torch.int64, device=dev) for _ in range(2)]
    bt = [torch.zeros(B, nQ, N, dtype=torch.int64, device=dev) for _ in range(2)]
    ot = [torch.zeros(B, nQ, N, dtype=torch.int64, device=dev) for _ in range(2)]
    rt = [torch.zeros(B, nQ, N, dtype=torch.int64, device=dev) for _ in range(2)]
    W = lambda t: ring.Poly.wrap(t.data_ptr(), N, nQ, B, keep=t)
    a, b, o, r = (tuple(W(t) for t in ts) for ts in (at, bt, ot, rt))
    cQ, cP = ring.NewContextWithParams(N, Q), ring.NewContextWithParams(N, P)
    ev = ckks.NewEvaluator(cQ, cP)
    key = ckks.SwitchingKey(evk)
    idx = ring.PermuteNTTIndex(5, 1, N)
    level = nQ - 1

    def step(stream):
        ev.MulRelin(level, a, b, key, o, stream=stream)
        ev.Rescale(nQ, o, 1, stream=stream)
        ev.permuteNTT(level - 1, o, idx, key, r, stream=stream)

    def load(seed):
        g = np.random.default_rng(seed)
        va = np.stack([g.integers(0, q, size=(B, 2, N), dtype=np.uint64) for q in Q], axis=2)
        vb = np.stack([g.integers(0, q, size=(B, 2, N), dtype=np.uint64) for q in Q], axis=2)
        for h in range(2):
            at[h].copy_(torch.from_numpy(va[:, h].copy().view(np.int64)))
            bt[h].copy_(torch.from_numpy(vb[:, h].copy().view(np.int64)))
        return va, vb

    side = torch.cuda.Stream()
    load(1)
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        step(side.cuda_stream)  # first use: derived key forms, pool growth, kernel attributes
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    n0 = ring.launch_count()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph, stream=side):
        step(torch.cuda.current_stream().cuda_stream)
    assert ring.launch_count() > n0  # the library's own kernels were captured
    oQ, oP = orc.Context(N, Q), orc.Context(N, P)
    oev = orc.CkksEvaluator(oQ, oP)
    widx = orc.permute_ntt_index(5, 1, N)
    for seed in (2, 3):  # new inputs, same graph
        va, vb = load(seed)
        for t in rt:
            t.zero_()
        graph.replay()
        torch.cuda.synchronize()
        got = np.stack([rt[0].cpu().numpy().view(np.uint64), rt[1].cpu().numpy().view(np.uint64)], axis=1)
        for i in range(B):
            w = oev.rescale(oev.mul_relin(level, va[i].copy(), vb[i].copy(), evk))
            w = oev.permute_ntt(level - 1, w, widx, evk)
            assert np.array_equal(got[i][:, : nQ - 1], w), (seed, i)
