"""Two-stream L2-grouped transforms: forward / inverse batched limb-NTT time (N = 2^16, 34 limbs x 32 ciphertexts) for
ntt_l2_streams in {0 (one stream), 1 (groups of batch entries on two streams), 2 (groups of limbs on two streams)} and
several L2 budgets; budget 0 = the shipped single launch pair.  Every result is bit-compared with the budget-0 output."""
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
nQ = len(Q)
B = 32
g = torch.Generator(device=dev)
g.manual_seed(1)
t = torch.empty(B, nQ, N, dtype=torch.int64, device=dev)
for i, q in enumerate(Q):
    t[:, i, :] = torch.randint(0, q, (B, N), dtype=torch.int64, device=dev, generator=g)
a = ring.Poly.wrap(t.data_ptr(), N, nQ, B, keep=t)
ot = torch.empty_like(t)
o = ring.Poly.wrap(ot.data_ptr(), N, nQ, B, keep=ot)
cQ = ring.NewContextWithParams(N, Q)
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


ring.debug_set_switch("ntt_l2_bytes", 0)
cQ.NTT(a, o, stream=sp)
torch.cuda.synchronize()
ref_f = ot.clone()
cQ.InvNTT(a, o, stream=sp)
torch.cuda.synchronize()
ref_i = ot.clone()
for mode in (0, 1, 2):
    for mib in ([0] if mode == 0 else [int(x) for x in sys.argv[1:]] or [16, 32, 48, 64, 96]):
        ring.debug_set_switch("ntt_l2_streams", mode)
        ring.debug_set_switch("ntt_l2_bytes", mib << 20)
        cQ.NTT(a, o, stream=sp)
        torch.cuda.synchronize()
        okf = bool(torch.equal(ot, ref_f))
        cQ.InvNTT(a, o, stream=sp)
        torch.cuda.synchronize()
        oki = bool(torch.equal(ot, ref_i))
        print(json.dumps({"streams_mode": mode, "l2_MiB": mib, "fwd_us": timed(lambda: cQ.NTT(a, o, stream=sp)),
                          "inv_us": timed(lambda: cQ.InvNTT(a, o, stream=sp)), "equal": okf and oki}), flush=True)
ring.debug_set_switch("ntt_l2_bytes", 0)
ring.debug_set_switch("ntt_l2_streams", 0)
