"""Limb axis of the multi-GPU path (BASELINE config 4) on ONE GPU: `world` ranks live in one process
on the same device (lattigpu.dist.Comm.inproc_group: exchange buffers attached by pointer instead of CUDA IPC), each
issues its limb-resident ops on its own stream, and the cross-rank barrier kernels of the ranks meet on the device.
This runs the production code of the limb axis -- cyclic limb ownership, peer-buffer source pointers of the basis
extensions, strided limb maps of the NTT / tensor / tail kernels, the barrier kernel -- where the driver's single-GPU box
can see it; tests/test_gpu_multi.py runs the same entry points over real peers (IPC + NVLink) when GPUs are available.

Checked bit for bit: every rank's own limbs of MulRelin + Rescale (ckks/evaluator.go:1016-1133, :933-968) and of
switchKeysInPlace (:1475-1558) against the single-GPU evaluator (itself pinned to the oracle), the gathered result,
and one case directly against the oracle.  Runs in a subprocess with CUDA_DEVICE_MAX_CONNECTIONS=32 so that the ranks'
streams never share a hardware queue (a rank waiting in a barrier must not hold back another rank's kernels) and with
eager module loading (a lazy first launch waits for running kernels, i.e. for a barrier whose peer is not issued yet)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import os, sys
ROOT = %r
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "lattigo-fhe-by-go_b200"))
import numpy as np
import torch
import lattigpu
from lattigpu import ckks, ring
from lattigpu import dist as ld
from oracle import ring_oracle as orc

world = int(sys.argv[1])
ring.set_device(0)
CASES = [dict(LogN=12, LogQi=[55] + [45] * 33, LogPi=[55] * 4),          # the headline digit shape: alpha 4, beta 9
         dict(LogN=12, LogQi=[50, 40, 40, 40, 40, 40, 40], LogPi=[50, 50, 50]),  # alpha 3, partial last digit
         dict(LogN=13, LogQi=[33, 30, 30, 30, 30, 30], LogPi=[35])]      # PN13QP218: alpha 1
for ci, params in enumerate(CASES):
    N = 1 << params["LogN"]
    Q, P = ckks.GenModuli(params)
    nQ, nP = len(Q), len(P)
    beta = -(-nQ // nP)
    rng = np.random.default_rng(770 + ci)
    evk = np.ascontiguousarray(np.stack([rng.integers(0, m, size=(beta, 2, N), dtype=np.uint64) for m in Q + P], axis=2))
    batch = 2
    a = np.ascontiguousarray(np.stack([rng.integers(0, m, size=(batch, 2, N), dtype=np.uint64) for m in Q], axis=2))
    b = np.ascontiguousarray(np.stack([rng.integers(0, m, size=(batch, 2, N), dtype=np.uint64) for m in Q], axis=2))
    cQ, cP = ring.NewContextWithParams(N, Q), ring.NewContextWithParams(N, P)
    ev = ckks.NewEvaluator(cQ, cP)
    key = ckks.SwitchingKey(evk)
    comms = ld.Comm.inproc_group(world, ev, batch)
    streams = [ring.Stream() for _ in range(world)]

    def polys(ct):
        return (ring.Poly.from_numpy(np.ascontiguousarray(ct[:, 0])), ring.Poly.from_numpy(np.ascontiguousarray(ct[:, 1])))

    def host(ct, nl):
        return np.stack([ct[0].numpy(nl=nl, squeeze=False), ct[1].numpy(nl=nl, squeeze=False)], axis=1)

    def fresh():
        return (ring.Poly(N, nQ, batch), ring.Poly(N, nQ, batch))

    pa, pb = polys(a), polys(b)
    levels = sorted({nQ - 1, nQ - 2, max(1, nQ // 2), 1}, reverse=True)
    for level in levels:
        nl = level + 1
        ref = fresh()
        ev.MulRelin(level, pa, pb, key, ref)
        want_mr = host(ref, nl)
        ev.Rescale(nl, ref)
        want = host(ref, nl - 1)
        torch.cuda.synchronize()
        # (1) limb-resident MulRelin + Rescale: all ranks enqueued before anyone is waited for
        outs = [fresh() for _ in range(world)]
        torch.cuda.synchronize()
        for r in range(world):
            comms[r].MulRelinRescale(ev, level, pa, pb, key, outs[r], nrescale=1, stream=streams[r])
        for r in range(world):
            comms[r].check(stream=streams[r])
        for r in range(world):
            got = host(outs[r], nl - 1)
            own = ld.own_limbs(nl - 1, world, r)
            assert np.array_equal(got[:, :, own], want[:, :, own]), ("MulRelinRescale own limbs", ci, level, r)
        # (2) gather: every rank ends with the complete result
        for r in range(world):
            comms[r].GatherLimbs(ev, nl - 1, outs[r], stream=streams[r])
        for r in range(world):
            comms[r].check(stream=streams[r])
        for r in range(world):
            assert np.array_equal(host(outs[r], nl - 1), want), ("gathered", ci, level, r)
        # (3) the replicated forms (round 1 interface): MulRelin, then Rescale
        outs = [fresh() for _ in range(world)]
        torch.cuda.synchronize()
        for r in range(world):
            comms[r].MulRelin(ev, level, pa, pb, key, outs[r], stream=streams[r])
        for r in range(world):
            comms[r].check(stream=streams[r])
        for r in range(world):
            assert np.array_equal(host(outs[r], nl), want_mr), ("MulRelin replicated", ci, level, r)
        for r in range(world):
            comms[r].Rescale(ev, nl, outs[r], stream=streams[r])
        for r in range(world):
            comms[r].check(stream=streams[r])
        for r in range(world):
            assert np.array_equal(host(outs[r], nl - 1), want), ("Rescale replicated", ci, level, r)
        # (4) switchKeysInPlace on a user-layout input (arbitrary 64-bit words: the range-flag path of the inverse NTT)
        cxw = rng.integers(0, 1 << 64, size=(batch, nQ, N), dtype=np.uint64)
        pcx = ring.Poly.from_numpy(cxw)
        r0, r1 = ring.Poly(N, nQ, batch), ring.Poly(N, nQ, batch)
        ev.switchKeysInPlace(level, pcx, key, r0, r1)
        w0, w1 = r0.numpy(nl=nl, squeeze=False), r1.numpy(nl=nl, squeeze=False)
        torch.cuda.synchronize()
        ps = [(ring.Poly(N, nQ, batch), ring.Poly(N, nQ, batch)) for _ in range(world)]
        torch.cuda.synchronize()
        for r in range(world):
            comms[r].switchKeysInPlaceResident(ev, level, pcx, key, ps[r][0], ps[r][1], stream=streams[r])
        for r in range(world):
            comms[r].check(stream=streams[r])
        for r in range(world):
            own = ld.own_limbs(nl, world, r)
            assert np.array_equal(ps[r][0].numpy(nl=nl, squeeze=False)[:, own], w0[:, own]), ("switchKeys", ci, level, r)
            assert np.array_equal(ps[r][1].numpy(nl=nl, squeeze=False)[:, own], w1[:, own]), ("switchKeys", ci, level, r)
    if ci == 2:  # and once directly against the oracle
        oev = orc.CkksEvaluator(orc.Context(N, Q), orc.Context(N, P))
        w = oev.rescale(oev.mul_relin(nQ - 1, np.ascontiguousarray(a[0]), np.ascontiguousarray(b[0]), evk))
        outs = [fresh() for _ in range(world)]
        torch.cuda.synchronize()
        for r in range(world):
            comms[r].MulRelinRescale(ev, nQ - 1, pa, pb, key, outs[r], stream=streams[r])
        for r in range(world):
            comms[r].GatherLimbs(ev, nQ - 1, outs[r], stream=streams[r])
        for r in range(world):
            comms[r].check(stream=streams[r])
        assert np.array_equal(host(outs[0], nQ - 1)[0], w)
    del comms
print("ok")
"""


@pytest.mark.parametrize("world", [2, 3, 4])
def test_limb_axis_ranks_in_one_process(world):
    env = dict(os.environ)
    env["CUDA_DEVICE_MAX_CONNECTIONS"] = "32"
    # a kernel's first launch loads its module, which waits for running kernels: with the ranks in ONE process the host
    # thread would block behind a rank's barrier kernel before the peer's work is issued (separate processes only stall)
    env["CUDA_MODULE_LOADING"] = "EAGER"
    res = subprocess.run([sys.executable, "-c", SCRIPT % ROOT, str(world)], env=env, capture_output=True, text=True, timeout=900)
    assert res.returncode == 0 and res.stdout.strip().endswith("ok"), (res.stdout[-2000:], res.stderr[-4000:])
