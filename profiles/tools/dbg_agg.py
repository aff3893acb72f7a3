import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
for p in (ROOT, os.path.join(ROOT, "lattigo-fhe-by-go_b200")):
    sys.path.insert(0, p)
import numpy as np, torch, torch.distributed as dist
import lattigpu
from lattigpu import ring
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank); ring.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
comm = lattigpu.dist.Comm()
N = 4096
Q = ring.GenerateNTTPrimes(45, 12, 3)
c = ring.NewContextWithParams(N, Q)
share = np.ascontiguousarray(np.stack([np.random.default_rng(1000 + rank).integers(0, m, size=(N,), dtype=np.uint64) for m in Q]))
ps = ring.Poly.from_numpy(share)
comm.AggregateShares(c, ps)
got = ps.numpy()
s = [np.ascontiguousarray(np.stack([np.random.default_rng(1000 + r).integers(0, m, size=(N,), dtype=np.uint64) for m in Q])) for r in range(world)]
want = sum(x.astype(object) for x in s)
want = np.array([[int(v) % Q[i] for v in want[i]] for i in range(3)], dtype=np.uint64)
print(rank, "match", np.array_equal(got, want), "mismatches", int((got != want).sum()), got[0, :3], want[0, :3], share[0, :3], flush=True)
dist.barrier(); dist.destroy_process_group()
