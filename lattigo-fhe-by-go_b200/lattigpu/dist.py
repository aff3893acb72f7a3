"""Multi-GPU plumbing: one process per GPU, torch.distributed for rendezvous, NCCL (through the
C ABI's lg_comm) for the data path.

Shard axes:
  * batch axis   -- independent ciphertexts, no collective: `shard_batch`
  * limb axis    -- one ciphertext, RNS limbs spread over the ranks: `Comm.MulRelin/Rescale/...`
                    (all-gather where a basis extension needs every limb)
  * party axis   -- dckks/dbfv shares, one party per GPU: `Comm.AggregateShares` (all-reduce + Reduce)
"""
import ctypes as C

from ._lib import check, lib, vp
from .ring import _s


def shard_batch(total, world, rank):
    """contiguous block of `total` independent items owned by `rank` (sizes differ by at most one)"""
    return (rank * total) // world, ((rank + 1) * total) // world


def limb_range(nlimbs, world, rank):
    """ownership rule of the limb axis (lg_comm_limb_range): [begin, end)"""
    b, e = C.c_int(0), C.c_int(0)
    check(lib().lg_comm_limb_range(nlimbs, world, rank, C.byref(b), C.byref(e)))
    return b.value, e.value


def max_over_ranks(value, group=None):
    """max of a python float over the ranks of a torch.distributed group (any backend)"""
    import torch
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return float(value)
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.tensor([float(value)], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def exchange_unique_id(make_id, group=None):
    """rank 0 creates the 128-byte NCCL id, every rank returns the same bytes"""
    import torch.distributed as dist

    rank = dist.get_rank(group) if dist.is_initialized() else 0
    box = [make_id() if rank == 0 else None]
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast_object_list(box, src=0, group=group)
    return box[0]


def _make_id():
    buf = (C.c_uint8 * 128)()
    check(lib().lg_comm_get_unique_id(buf))
    return bytes(buf)


class Comm:
    """lg_comm over the ranks of the current torch.distributed world (call after ring.set_device)"""

    def __init__(self, world=None, rank=None, unique_id=None):
        import torch.distributed as dist

        if world is None:
            world = dist.get_world_size() if dist.is_initialized() else 1
            rank = dist.get_rank() if dist.is_initialized() else 0
        self.world, self.rank = world, rank
        if world > 1 and unique_id is None:
            unique_id = exchange_unique_id(_make_id)
        h = vp()
        idbuf = (C.c_uint8 * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        check(lib().lg_comm_create(world, rank, idbuf, C.byref(h)))
        self.h = h

    def __del__(self):
        try:
            lib().lg_comm_destroy(self.h)
        except Exception:
            pass

    def AggregateShares(self, context, share, nl=None, stream=None):
        """dckks/dbfv AggregateShares across ranks: share <- Reduce(sum over ranks of share)"""
        check(lib().lg_comm_aggregate_shares(self.h, context.h, context.nl if nl is None else nl, share.h, _s(stream)))

    def switchKeysInPlace(self, evaluator, level, cx, evakey, p0, p1, stream=None):
        check(lib().lg_ckks_switch_keys_in_place_sharded(evaluator.h, self.h, level, cx.h, evakey.h, p0.h, p1.h, _s(stream)))

    def MulRelin(self, evaluator, level, ct0, ct1, evakey, ctOut, stream=None):
        check(lib().lg_ckks_mul_relin_sharded(evaluator.h, self.h, level, ct0[0].h, ct0[1].h, ct1[0].h, ct1[1].h, evakey.h,
                                              ctOut[0].h, ctOut[1].h, _s(stream)))

    def Rescale(self, evaluator, nl, ct, stream=None):
        check(lib().lg_ckks_rescale_sharded(evaluator.h, self.h, nl, ct[0].h, ct[1].h, _s(stream)))

    # ---- limb-resident forms: a ciphertext stays spread over the ranks between ops -------------------------
    def MulRelinRescale(self, evaluator, level, ct0, ct1, evakey, ctOut, stream=None):
        """MulRelin at `level` followed by one Rescale; ctOut holds this rank's own limbs of the level-1 result
        (GatherLimbs replicates them)."""
        self.MulRelin(evaluator, level, ct0, ct1, evakey, ctOut, stream=stream)
        self.Rescale(evaluator, level + 1, ctOut, stream=stream)

    def GatherLimbs(self, evaluator, nl, ct, stream=None):
        """replicate the first nl limbs of a limb-resident ciphertext on every rank (no-op while outputs are replicated)"""
        return None

    def exchange_description(self):
        return ("ncclBroadcast groups (one per owner rank and batch entry) of: coefficient-domain c2 before "
                "DecomposeAndSplit, special-prime accumulators before ModDown, result limbs, rescaled limbs")

    def exchange_bytes(self, evaluator, level, batch):
        """bytes one rank receives per MulRelin+Rescale"""
        N, nl, nP, w = evaluator.contextQ.N, level + 1, evaluator.contextP.nl, self.world
        words = batch * N * (nl + 2 * nP + 2 * nl + 2 * (nl - 1))
        return int(words * 8 * (w - 1) / w)
