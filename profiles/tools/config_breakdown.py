"""Per-kernel time and DRAM bytes of the last step in an ncu launch list taken with
`--metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` (profiles/tools/config_launches.sh).
Usage: python profiles/tools/config_breakdown.py launches.csv marker-kernel-substring [--period] [-v]"""
import collections
import csv
import sys

UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6, "nsecond": 1.0, "usecond": 1e3, "msecond": 1e6}


def main(path, first, verbose):
    launches = collections.OrderedDict()
    for r in csv.reader(open(path)):
        if len(r) < 15 or not r[0].isdigit():
            continue
        name = r[4].replace("<unnamed>::", "").replace("void ", "").split("(")[0]
        d = launches.setdefault(int(r[0]), {"name": name, "grid": r[8]})
        d[r[12]] = float(r[14].replace(",", "")) * UNIT.get(r[13], 1.0)
    ls = list(launches.values())
    starts = [i for i, l in enumerate(ls) if l["name"].startswith(first)]
    # the launches between the last two marker kernels = one step of the periodic sequence (rotated); a list that ends
    # with the step (marker first) is cut from the last marker to the end instead
    step = ls[starts[-2]:starts[-1]] if "--period" in sys.argv else ls[starts[-1]:]
    t = sum(l["gpu__time_duration.sum"] for l in step) / 1e3
    rd = sum(l["dram__bytes_read.sum"] for l in step)
    wr = sum(l["dram__bytes_write.sum"] for l in step)
    print("last step: %d launches, %.1f us serialised (cold cache), DRAM read %.2f GB + write %.2f GB = %.2f TB/s average"
          % (len(step), t, rd / 1e9, wr / 1e9, (rd + wr) / t / 1e6))
    agg = collections.OrderedDict()
    for l in step:
        a = agg.setdefault(l["name"], [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += l["gpu__time_duration.sum"] / 1e3
        a[2] += l["dram__bytes_read.sum"]
        a[3] += l["dram__bytes_write.sum"]
    print("%-40s %4s %10s %6s %9s %9s %7s" % ("kernel", "n", "us", "share", "rd MB", "wr MB", "TB/s"))
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-40s %4d %10.1f %5.1f%% %9.1f %9.1f %7.2f" % (k[:40], a[0], a[1], 100 * a[1] / t, a[2] / 1e6, a[3] / 1e6, (a[2] + a[3]) / a[1] / 1e6))
    if verbose:
        for l in step:
            print(l["name"], l["grid"], "%.1f us" % (l["gpu__time_duration.sum"] / 1e3), "%.1f MB" % ((l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"]) / 1e6))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2], "-v" in sys.argv)
