"""Does data written by one kernel stay in L2 for the next kernel?  In-place read-modify-write passes over a buffer of
X MiB (torch x.add_(1)): bytes moved per second against X.  Above the HBM rate = the passes hit L2."""
import json
import torch

dev = torch.device("cuda", 0)
for mib in (8, 16, 32, 48, 64, 80, 96, 128, 192, 256, 512, 1024):
    x = torch.zeros(mib * (1 << 20) // 8, dtype=torch.int64, device=dev)
    for _ in range(5):
        x.add_(1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 200 if mib <= 128 else 50
    e0.record()
    for _ in range(reps):
        x.add_(1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(json.dumps({"MiB": mib, "us_per_pass": ms * 1e3, "GBps_read_plus_write": 2 * mib * (1 << 20) / (ms * 1e-3) / 1e9}), flush=True)
