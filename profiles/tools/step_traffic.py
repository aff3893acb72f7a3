"""Total DRAM bytes of one MulRelin+Rescale step from the per-launch table of profiles/tools/ncu_table.py (columns rdMB, wrMB)
-> profiles/r02_step_traffic.json, read by bench.py for roofline_op.traffic_step.
Usage: python profiles/tools/step_traffic.py summary.txt batch N out.json"""
import json
import sys


def main(path, batch, N, out):
    rd = wr = us = 0.0
    n = 0
    per = {}
    for line in open(path):
        f = line.split()
        if line.startswith("#") or line.startswith("kernel") or len(f) < 6:
            continue
        # name may contain spaces ("ntt_fwd_strided<8, 0>"): the grid is the first field that starts with "("
        gi = next(i for i, x in enumerate(f) if x.startswith("("))
        name = " ".join(f[:gi])
        t, r, w = float(f[gi + 1]), float(f[gi + 2]), float(f[gi + 3])
        us += t
        rd += r
        wr += w
        n += 1
        k = per.setdefault(name, {"launches": 0, "us": 0.0, "dram_MB": 0.0})
        k["launches"] += 1
        k["us"] += t
        k["dram_MB"] += r + w
    json.dump({"batch": batch, "N": N, "launches": n, "us_serialised": us, "dram_read_bytes": rd * 1e6, "dram_write_bytes": wr * 1e6,
               "dram_bytes_per_step": (rd + wr) * 1e6, "per_kernel": per}, open(out, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4])
