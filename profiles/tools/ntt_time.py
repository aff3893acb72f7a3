"""Times the batched forward / inverse NTT (and optionally MulRelin+Rescale) of one library build.
Usage: LATTIGPU_LIB=path/to/lib.so python profiles/tools/ntt_time.py [logN] [limbs] [batch]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "lattigo-fhe-by-go_b200"))
import torch

from lattigpu import ring


def main():
    logN = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    nl = int(sys.argv[2]) if len(sys.argv) > 2 else 34
    batch = int(sys.argv[3]) if len(sys.argv) > 3 else 8
    N = 1 << logN
    ring.set_device(0)
    moduli = ring.GenerateNTTPrimes(45, logN, nl)
    ctx = ring.NewContextWithParams(N, moduli)
    t = torch.randint(0, moduli[0], (batch, nl, N), dtype=torch.int64, device="cuda")
    o = torch.empty_like(t)
    a = ring.Poly.wrap(t.data_ptr(), N, nl, batch, keep=t)
    b = ring.Poly.wrap(o.data_ptr(), N, nl, batch, keep=o)
    sp = torch.cuda.current_stream().cuda_stream
    for name, fn in (("fwd", ctx.NTT), ("inv", ctx.InvNTT)):
        for _ in range(5):
            fn(a, b, stream=sp)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 50
        e0.record()
        for _ in range(reps):
            fn(a, b, stream=sp)
        e1.record()
        torch.cuda.synchronize()
        us = 1e3 * e0.elapsed_time(e1) / reps
        print("%s %s logN=%d limbs=%d: %.1f us/launch-pair, %.3f M limb-NTT/s" %
              (os.path.basename(os.environ.get("LATTIGPU_LIB", "default")), name, logN, nl * batch, us, nl * batch / us))


main()
