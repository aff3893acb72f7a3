"""Stall-sample and opcode mix summary of one kernel from `ncu --page source --csv --print-source sass`."""
import collections
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]
    data = rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    print(rows[0][1])
    tot = collections.Counter()
    n = samples = 0
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    byop = collections.defaultdict(lambda: [0, 0])
    for d in data:
        if len(d) < len(hdr):
            continue
        if not d[ix["Instructions Executed"]].isdigit():
            continue
        ex = int(d[ix["Instructions Executed"]] or 0)
        if ex == 0:
            continue
        n += 1
        s = int(d[ix["# Samples"]] or 0)
        samples += s
        for st in stalls:
            tot[st] += int(d[ix[st]] or 0)
        f = d[ix["Source"]].split()
        op = f[1] if f[0].startswith("@") else f[0]
        byop[op][0] += ex
        byop[op][1] += s
    print("executed sass lines", n, "samples", samples)
    for k, v in tot.most_common(8):
        print("%-24s %6d %5.1f%%" % (k, v, 100 * v / max(samples, 1)))
    te = sum(v[0] for v in byop.values())
    print("warp-instructions executed: %d" % te)
    for k, v in sorted(byop.items(), key=lambda kv: -kv[1][0])[:14]:
        print("%-22s exec %9d (%4.1f%%) samples %6d (%4.1f%%)" % (k, v[0], 100 * v[0] / te, v[1], 100 * v[1] / max(samples, 1)))


if __name__ == "__main__":
    main(sys.argv[1])
