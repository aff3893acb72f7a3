"""Multi-GPU plumbing: one process per GPU, torch.distributed for rendezvous; the data path goes through the
C ABI's lg_comm.

Shard axes:
  * batch axis   -- independent ciphertexts, no collective: `shard_batch`
  * limb axis    -- one ciphertext, RNS limbs spread cyclically over the ranks (limb t of Q || P on rank t mod world).
                    Ciphertexts stay limb-resident between ops; limbs cross NVLink through peer memory exactly where
                    a basis extension needs every source limb (no library collective): `Comm.MulRelinRescale`,
                    `Comm.switchKeysInPlaceResident`, `Comm.RescaleResident`, `Comm.GatherLimbs`; the replicated
                    forms `Comm.MulRelin` / `Comm.Rescale` / `Comm.switchKeysInPlace` gather their results
  * party axis   -- dckks/dbfv shares, one group of parties per GPU: `Comm.AggregateShares` (NCCL all-reduce + Reduce)
"""
import ctypes as C

from ._lib import check, lib, vp
from .ring import _s


def shard_batch(total, world, rank):
    """contiguous block of `total` independent items owned by `rank` (sizes differ by at most one)"""
    return (rank * total) // world, ((rank + 1) * total) // world


def limb_owner(limb, world):
    """ownership rule of the limb axis (lg_comm_limb_owner): limb t of Q || P belongs to rank t mod world"""
    return int(lib().lg_comm_limb_owner(limb, world))


def own_limbs(nlimbs, world, rank):
    """the limbs among [0, nlimbs) that `rank` owns"""
    return [j for j in range(nlimbs) if limb_owner(j, world) == rank]


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
    """lg_comm over the ranks of the current torch.distributed world (call after ring.set_device).
    nccl=False creates a handle for the peer-memory limb axis only (no NCCL communicator)."""

    def __init__(self, world=None, rank=None, unique_id=None, nccl=True):
        import torch.distributed as dist

        if world is None:
            world = dist.get_world_size() if dist.is_initialized() else 1
            rank = dist.get_rank() if dist.is_initialized() else 0
        self.world, self.rank = world, rank
        if world > 1 and unique_id is None and nccl:
            unique_id = exchange_unique_id(_make_id)
        h = vp()
        idbuf = (C.c_uint8 * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        check(lib().lg_comm_create(world, rank, idbuf, C.byref(h)))
        self.h = h
        self._peers = []  # keeps in-process peers alive

    def __del__(self):
        try:
            lib().lg_comm_destroy(self.h)
        except Exception:
            pass

    # ---- exchange buffers of the limb axis ---------------------------------------------------------------------
    def xbuf_words(self):
        return int(lib().lg_comm_xbuf_words(self.h))

    @staticmethod
    def words_needed(N, nQ, nP, batch):
        """exchange words MulRelin + Rescale + the gathers of both polynomials need (lg_comm_xbuf_words_needed)"""
        return int(lib().lg_comm_xbuf_words_needed(N, nQ, nP, batch))

    def reserve(self, evaluator, batch):
        """Collective: size the exchange buffer for `batch` ciphertexts of `evaluator`'s parameters (see reserve_words)"""
        self.reserve_words(self.words_needed(evaluator.contextQ.N, evaluator.contextQ.nl, evaluator.contextP.nl, batch))

    def reserve_words(self, words):
        """Collective: allocate this rank's exchange buffer and map every peer's (CUDA IPC, the handles travel through
        torch.distributed).  The buffer is allocated once: reserve the largest shape first."""
        import torch.distributed as dist

        if self.xbuf_words() >= words:
            return
        if self.xbuf_words():
            raise RuntimeError("Comm.reserve: the exchange buffer is already allocated with %d words, %d needed; reserve the "
                               "largest shape first" % (self.xbuf_words(), words))
        buf = (C.c_uint8 * 128)()
        check(lib().lg_comm_xbuf_alloc(self.h, words, buf))
        if self.world > 1:
            handles = [None] * self.world
            dist.all_gather_object(handles, bytes(buf))
            for r, hb in enumerate(handles):
                if r != self.rank:
                    check(lib().lg_comm_xbuf_open(self.h, r, (C.c_uint8 * 128).from_buffer_copy(hb)))
            dist.barrier()

    @staticmethod
    def inproc_group(world, evaluator, batch):
        """`world` ranks living in ONE process on the current device (tests: the limb-axis logic without several GPUs).
        Their ops must be issued on distinct non-blocking streams, every rank's call before any is synchronised."""
        comms = [Comm(world, r, nccl=False) for r in range(world)]
        words = int(lib().lg_comm_xbuf_words_needed(evaluator.contextQ.N, evaluator.contextQ.nl, evaluator.contextP.nl, batch))
        for c in comms:
            check(lib().lg_comm_xbuf_alloc(c.h, words, None))
        for c in comms:
            for p in comms:
                if p is not c:
                    check(lib().lg_comm_xbuf_attach(c.h, p.rank, p.h))
                    c._peers.append(p)
        return comms

    def check(self, stream=None):
        """synchronise `stream` and raise if a limb-axis barrier timed out"""
        check(lib().lg_comm_check(self.h, _s(stream)))

    # ---- party axis -------------------------------------------------------------------------------------------------
    def AggregateShares(self, context, share, nl=None, stream=None):
        """dckks/dbfv AggregateShares across ranks: share <- Reduce(sum over ranks of share)"""
        check(lib().lg_comm_aggregate_shares(self.h, context.h, context.nl if nl is None else nl, share.h, _s(stream)))

    # ---- limb axis, replicated outputs ---------------------------------------------------------------------------
    def switchKeysInPlace(self, evaluator, level, cx, evakey, p0, p1, stream=None):
        check(lib().lg_ckks_switch_keys_in_place_sharded(evaluator.h, self.h, level, cx.h, evakey.h, p0.h, p1.h, _s(stream)))

    def MulRelin(self, evaluator, level, ct0, ct1, evakey, ctOut, stream=None):
        check(lib().lg_ckks_mul_relin_sharded(evaluator.h, self.h, level, ct0[0].h, ct0[1].h, ct1[0].h, ct1[1].h, evakey.h,
                                              ctOut[0].h, ctOut[1].h, _s(stream)))

    def Rescale(self, evaluator, nl, ct, stream=None):
        check(lib().lg_ckks_rescale_sharded(evaluator.h, self.h, nl, ct[0].h, ct[1].h, _s(stream)))

    # ---- limb axis, limb-resident forms: a ciphertext stays spread over the ranks between ops ---------------------
    def switchKeysInPlaceResident(self, evaluator, level, cx, evakey, p0, p1, stream=None):
        check(lib().lg_ckks_switch_keys_in_place_resident(evaluator.h, self.h, level, cx.h, evakey.h, p0.h, p1.h, _s(stream)))

    def MulRelinRescale(self, evaluator, level, ct0, ct1, evakey, ctOut, nrescale=1, stream=None):
        """MulRelin at `level` followed by `nrescale` Rescale steps; ctOut holds this rank's own limbs of the result
        (GatherLimbs replicates them)."""
        check(lib().lg_ckks_mul_relin_rescale_resident(evaluator.h, self.h, level, ct0[0].h, ct0[1].h, ct1[0].h, ct1[1].h,
                                                       evakey.h, ctOut[0].h, ctOut[1].h, nrescale, _s(stream)))

    def RescaleResident(self, evaluator, nl, ct, stream=None):
        check(lib().lg_ckks_rescale_resident(evaluator.h, self.h, nl, ct[0].h, ct[1].h, _s(stream)))

    def GatherLimbs(self, evaluator, nl, ct, stream=None):
        """replicate the first nl limbs of a limb-resident ciphertext on every rank"""
        for p in ct:
            check(lib().lg_comm_gather_limbs(self.h, evaluator.contextQ.h, nl, p.h, _s(stream)))

    def exchange_description(self):
        return ("peer memory over NVLink, no library collective: DecomposeAndSplit loads the coefficient-domain c2 limbs, the "
                "ModDown basis extension loads the special-prime accumulator limbs, the rescale copies the last limb -- each "
                "from the owner's IPC-mapped exchange buffer, ordered by a one-CTA barrier kernel in stream order "
                "(3 barriers per MulRelin+Rescale); limbs owned cyclically (limb t on rank t mod world)")

    def exchange_bytes(self, evaluator, level, batch):
        """bytes one rank loads over NVLink per MulRelin+Rescale of `batch` ciphertexts (upper bound: rank != owner)"""
        N, nl, nP, w = evaluator.contextQ.N, level + 1, evaluator.contextP.nl, self.world
        words = batch * N * ((nl + 2 * nP) * (w - 1) / w + 2)
        return int(words * 8)
