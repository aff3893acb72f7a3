"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: the limb-axis ownership rule, batch
sharding, the unique-id exchange and the max-over-ranks timing reduction used by bench.py.  The
data-path collectives themselves (NCCL) are covered on GPUs by tests/test_gpu_multi.py."""
import os
import sys

import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, os.path.join(ROOT, "lattigo-fhe-by-go_b200"))
    import torch.distributed as dist

    from lattigpu import dist as ld

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # every rank derives the same (cyclic) partition of the 38 table limbs of CKKS PN16 (34 Q + 4 P)
        mine = ld.own_limbs(38, world, rank)
        all_sets = [None] * world
        dist.all_gather_object(all_sets, mine)
        assert all_sets == [ld.own_limbs(38, world, r) for r in range(world)]
        assert sorted(j for s in all_sets for j in s) == list(range(38))
        # batch axis
        blocks = [ld.shard_batch(1024, world, r) for r in range(world)]
        assert blocks[0][0] == 0 and blocks[-1][1] == 1024 and all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
        # the 128-byte id made on rank 0 reaches everyone unchanged
        uid = ld.exchange_unique_id(lambda: bytes(range(128)))
        assert uid == bytes(range(128))
        # timing reduction: max over ranks
        assert ld.max_over_ranks(1.0 + rank) == float(world)
        q.put((rank, "ok"))
    except Exception as exc:  # pragma: no cover
        q.put((rank, repr(exc)))
    finally:
        dist.destroy_process_group()


def test_gloo_world2_host_logic():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29650 + os.getpid() % 200
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert res == {0: "ok", 1: "ok"}, res


@pytest.mark.parametrize("n,world", [(38, 8), (38, 3), (5, 8), (34, 1), (12, 4)])
def test_limb_ownership_partition(n, world):
    """cyclic ownership: a partition, balanced within one limb at EVERY level (prefix of the limb list), and stable when
    the last limb is dropped (a rescale never moves a limb between ranks)"""
    sys.path.insert(0, os.path.join(ROOT, "lattigo-fhe-by-go_b200"))
    from lattigpu import dist as ld

    for nl in range(1, n + 1):
        sets = [ld.own_limbs(nl, world, r) for r in range(world)]
        assert sorted(j for s in sets for j in s) == list(range(nl))
        sizes = [len(s) for s in sets]
        assert max(sizes) - min(sizes) <= 1
        if nl > 1:
            prev = [ld.own_limbs(nl - 1, world, r) for r in range(world)]
            assert all(set(p) <= set(s) for p, s in zip(prev, sets))
