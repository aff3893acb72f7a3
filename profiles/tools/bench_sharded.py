"""BASELINE config 4, limb axis: ONE CKKS PN16QP1761 ciphertext pair, MulRelin + Rescale with the RNS limbs spread
over the ranks (NCCL all-gathers where a basis extension needs every limb), against the same op on one GPU.
Run under torchrun:  python -m torch.distributed.run --nproc-per-node N profiles/tools/bench_sharded.py [batch]
Prints one JSON line on rank 0: latency per op sharded / replicated, device-timed, max over ranks."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "lattigo-fhe-by-go_b200"))
import torch
import torch.distributed as dist

import lattigpu
from lattigpu import ckks, ring


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    ring.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    dev = torch.device("cuda", lr)
    p = ckks.DefaultParams[ckks.PN16QP1761]
    N = 1 << p["LogN"]
    Q, P = ckks.GenModuli(p)
    nQ, nP = len(Q), len(P)
    beta = -(-nQ // nP)
    level = nQ - 1
    g = torch.Generator(device=dev)
    g.manual_seed(0x1A771C0 + 4)  # same ciphertext on every rank (replicated input)

    def uniform(prefix, moduli):
        t = torch.empty(*prefix, len(moduli), N, dtype=torch.int64, device=dev)
        for i, q in enumerate(moduli):
            t[..., i, :] = torch.randint(0, q, (*prefix, N), dtype=torch.int64, device=dev, generator=g)
        return t

    wrap = lambda t: ring.Poly.wrap(t.data_ptr(), N, nQ, batch, keep=t)
    evk_t = uniform((beta, 2), Q + P)
    rlk = ckks.SwitchingKey(N=N, device_ptr=evk_t.data_ptr(), beta=beta, nQP=nQ + nP, keep=evk_t)
    a = tuple(wrap(uniform((batch,), Q)) for _ in range(2))
    b = tuple(wrap(uniform((batch,), Q)) for _ in range(2))
    o = tuple(wrap(torch.empty(batch, nQ, N, dtype=torch.int64, device=dev)) for _ in range(2))
    cQ, cP = ring.NewContextWithParams(N, Q), ring.NewContextWithParams(N, P)
    ev = ckks.NewEvaluator(cQ, cP)
    comm = lattigpu.dist.Comm()
    sp = torch.cuda.current_stream().cuda_stream

    def sharded():
        comm.MulRelin(ev, level, a, b, rlk, o, stream=sp)
        comm.Rescale(ev, nQ, o, stream=sp)

    def local():
        ev.MulRelin(level, a, b, rlk, o, stream=sp)
        ev.Rescale(nQ, o, stream=sp)

    def timed(fn, reps=20, warm=5):
        for _ in range(warm):
            fn()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return lattigpu.dist.max_over_ranks(e0.elapsed_time(e1) / reps)

    t_local = timed(local)
    t_shard = timed(sharded)
    if rank == 0:
        print(json.dumps({"config": "C4 CKKS PN16QP1761 MulRelin+Rescale, one ciphertext batch of %d, limbs sharded" % batch,
                          "n_gpus": world, "ms_per_op_one_gpu": t_local / batch, "ms_per_op_sharded": t_shard / batch,
                          "speedup": t_local / t_shard}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
