"""Multi-GPU data paths on real GPUs (needs >= 2 devices; skipped otherwise):
  * limb axis: CKKS switchKeys / MulRelin / Rescale with the RNS limbs of one ciphertext spread over
    the ranks and NCCL all-gathers where a basis extension needs every limb -- must be bit-identical
    to the single-GPU path and to the oracle;
  * party axis: AggregateShares = all-reduce(sum) + Reduce equals the reference's chain of
    context.Add over the parties' shares (dckks/publickey_gen.go:45-47)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, q):
    for p in (ROOT, os.path.join(ROOT, "lattigo-fhe-by-go_b200")):
        sys.path.insert(0, p)
    import torch.distributed as dist

    import lattigpu
    from lattigpu import ckks, ring
    from oracle import ring_oracle as orc

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    ring.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        comm = lattigpu.dist.Comm()
        cases = (dict(LogN=13, LogQi=[33, 30, 30, 30, 30, 30], LogPi=[35]),  # PN13QP218
                 dict(LogN=12, LogQi=[50, 40, 40, 40, 40, 40, 40], LogPi=[50, 50, 50]),
                 dict(LogN=12, LogQi=[55] + [45] * 33, LogPi=[55] * 4),  # the headline digit shape (alpha 4, beta 9)
                 dict(LogN=14, LogQi=[45] + [34] * 9, LogPi=[43, 43]))  # PN14QP438
        # exchange buffers of the limb axis: mapped between the processes with CUDA IPC, sized for the largest case
        comm.reserve_words(max(comm.words_needed(1 << c["LogN"], len(c["LogQi"]), len(c["LogPi"]), 2) for c in cases))
        for params in cases:
            N = 1 << params["LogN"]
            Q, P = ckks.GenModuli(params)
            nQ, nP = len(Q), len(P)
            beta = -(-nQ // nP)
            rng = np.random.default_rng(77)  # same inputs on every rank (replicated ciphertext)
            evk = np.ascontiguousarray(np.stack([rng.integers(0, m, size=(beta, 2, N), dtype=np.uint64) for m in Q + P], axis=2))
            batch = 2
            a = np.ascontiguousarray(np.stack([rng.integers(0, m, size=(batch, 2, N), dtype=np.uint64) for m in Q], axis=2))
            b = np.ascontiguousarray(np.stack([rng.integers(0, m, size=(batch, 2, N), dtype=np.uint64) for m in Q], axis=2))
            cQ, cP = ring.NewContextWithParams(N, Q), ring.NewContextWithParams(N, P)
            ev = ckks.NewEvaluator(cQ, cP)
            key = ckks.SwitchingKey(evk)

            def polys(ct):
                return (ring.Poly.from_numpy(np.ascontiguousarray(ct[:, 0])), ring.Poly.from_numpy(np.ascontiguousarray(ct[:, 1])))

            def host(ct, nl):
                return np.stack([ct[0].numpy(nl=nl, squeeze=False), ct[1].numpy(nl=nl, squeeze=False)], axis=1)

            for level in (nQ - 1, nQ - 2):
                nl = level + 1
                ref = (ring.Poly(N, nQ, batch), ring.Poly(N, nQ, batch))
                ev.MulRelin(level, polys(a), polys(b), key, ref)
                want_mr = host(ref, nl)
                ev.Rescale(nl, ref)
                want_rs = host(ref, nl - 1)
                out = (ring.Poly(N, nQ, batch), ring.Poly(N, nQ, batch))
                comm.MulRelin(ev, level, polys(a), polys(b), key, out)
                assert np.array_equal(host(out, nl), want_mr), ("MulRelin", params["LogN"], level)
                comm.Rescale(ev, nl, out)
                assert np.array_equal(host(out, nl - 1), want_rs), ("Rescale", params["LogN"], level)
                p0, p1 = ring.Poly(N, nQ, batch), ring.Poly(N, nQ, batch)
                r0, r1 = ring.Poly(N, nQ, batch), ring.Poly(N, nQ, batch)
                cx = polys(a)[1]
                ev.switchKeysInPlace(level, cx, key, r0, r1)
                comm.switchKeysInPlace(ev, level, cx, key, p0, p1)
                assert np.array_equal(p0.numpy(nl=nl, squeeze=False), r0.numpy(nl=nl, squeeze=False))
                assert np.array_equal(p1.numpy(nl=nl, squeeze=False), r1.numpy(nl=nl, squeeze=False))
            if rank == 0 and params["LogN"] == 13:  # and against the oracle once
                oev = orc.CkksEvaluator(orc.Context(N, Q), orc.Context(N, P))
                w = oev.rescale(oev.mul_relin(nQ - 1, np.ascontiguousarray(a[0]), np.ascontiguousarray(b[0]), evk))
                out = (ring.Poly(N, nQ, batch), ring.Poly(N, nQ, batch))
                # all ranks must enter the collective: done below, outside the rank test
            out = (ring.Poly(N, nQ, batch), ring.Poly(N, nQ, batch))
            comm.MulRelin(ev, nQ - 1, polys(a), polys(b), key, out)
            comm.Rescale(ev, nQ, out)
            if rank == 0 and params["LogN"] == 13:
                assert np.array_equal(host(out, nQ - 1)[0], w)
            # limb-resident forms: own limbs after MulRelin + Rescale, then the gathered result
            ref = (ring.Poly(N, nQ, batch), ring.Poly(N, nQ, batch))
            ev.MulRelin(nQ - 1, polys(a), polys(b), key, ref)
            ev.Rescale(nQ, ref)
            want = host(ref, nQ - 1)
            out = (ring.Poly(N, nQ, batch), ring.Poly(N, nQ, batch))
            comm.MulRelinRescale(ev, nQ - 1, polys(a), polys(b), key, out)
            comm.check()
            own = lattigpu.dist.own_limbs(nQ - 1, world, rank)
            assert np.array_equal(host(out, nQ - 1)[:, :, own], want[:, :, own]), ("resident", params["LogN"])
            comm.GatherLimbs(ev, nQ - 1, out)
            comm.check()
            assert np.array_equal(host(out, nQ - 1), want), ("gathered", params["LogN"])

            # party axis: every rank holds one party's share over QP
            cQP = ring.NewContextWithParams(N, Q + P)
            srng = np.random.default_rng(1000 + rank)
            share = np.ascontiguousarray(np.stack([srng.integers(0, m, size=(N,), dtype=np.uint64) for m in Q + P]))
            ps = ring.Poly.from_numpy(share)
            comm.AggregateShares(cQP, ps)
            oQP = orc.Context(N, Q + P)
            agg = None
            for r in range(world):
                g = np.random.default_rng(1000 + r)
                s = np.ascontiguousarray(np.stack([g.integers(0, m, size=(N,), dtype=np.uint64) for m in Q + P]))
                agg = s if agg is None else oQP.op3("add", agg, s)
            assert np.array_equal(ps.numpy(), agg)
        dist.barrier()
        q.put((rank, "ok"))
    except Exception as exc:  # pragma: no cover
        import traceback

        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_paths(world):
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29700 + world
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=600) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert all(v == "ok" for v in res.values()), res
