"""Aggregates one bench step out of an `ncu --metrics gpu__time_duration.sum --csv` launch list.
Usage: python profiles/tools/step_breakdown.py launches.csv [first-kernel-substring]"""
import collections
import csv
import sys


def main(path, first="tensor_kernel"):
    rows = [r for r in csv.reader(open(path)) if len(r) > 14 and r[0].isdigit()]
    names = [(r[4].replace("<unnamed>::", "").replace("void ", "").split("(")[0], r[8], int(r[14])) for r in rows]
    idx = [i for i, n in enumerate(names) if n[0].startswith(first)]
    s = idx[-1]
    e = len(names)
    # One MulRelin+Rescale = tensor ... key switch ... ModDown ... two rescalings.  Since the tails ride on the forward
    # transform, the step ends with the third forward contiguous-phase launch after the tensor product (ModDown pair,
    # rescale c0, rescale c1); older lists end with the last ew_kernel before the NTT-rate loop.
    fwd = [i for i, n in enumerate(names) if i > s and n[0].startswith("ntt_contig_pipe<1")]
    tails = [i for i, n in enumerate(names) if i > s and n[0].startswith("ew_kernel<3")]
    if len(fwd) >= 3 and (not tails or tails[0] > fwd[2]):
        last = fwd[2]
    else:
        last = max(i for i, n in enumerate(names) if i >= s and n[0].startswith("ew_kernel"))
    step = names[s:last + 1]
    tot = sum(n[2] for n in step)
    print("step: %d launches, %.1f us (serialised, cold cache)" % (len(step), tot / 1e3))
    agg = collections.OrderedDict()
    for n in step:
        agg.setdefault(n[0], [0, 0])
        agg[n[0]][0] += 1
        agg[n[0]][1] += n[2]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-44s %3d %9.1f us %5.1f%%" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))
    if "-v" in sys.argv:
        for n in step:
            print(n)


if __name__ == "__main__":
    main(sys.argv[1], *(a for a in sys.argv[2:] if a != "-v"))
