"""DRAM bytes of the forward NTT launch pair (strided + pipelined contiguous phase) from an
`ncu --set full --page raw --csv` export -> profiles/r0N_ncu_ntt_fwd.json, read by bench.py for roofline.traffic.
Usage: python profiles/tools/ntt_fwd_traffic.py raw.csv N limbs batch out.json"""
import csv
import json
import sys


def main(path, N, limbs, batch, out):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}

    def bytes_of(d, key):
        v = float(d[ix[key]].replace(",", ""))
        u = units[ix[key]]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]

    found = {}
    for d in data:
        name = d[ix["Kernel Name"]]
        grid = d[ix["Grid Size"]].replace(" ", "")
        g = [int(x) for x in grid.strip("()").split(",")]
        if "ntt_fwd_strided" in name and g[2] == limbs:  # ntt_fwd_strided (batch, tiles, limbs) / _tma (tiles, batch groups, limbs)
            found["strided"] = d
        if "ntt_contig_pipe<(bool)1" in name.replace(" ", "") or "ntt_contig_pipe<1" in name.replace(" ", ""):
            if g[2] == limbs and g[1] == N // 2048:
                found["contig"] = d
    assert len(found) == 2, "launch pair not found: %s" % list(found)
    res = {"N": N, "limb_ntts_per_launch": limbs * batch, "kernels": {}}
    total = 0.0
    for k, d in found.items():
        rd, wr = bytes_of(d, "dram__bytes_read.sum"), bytes_of(d, "dram__bytes_write.sum")
        dur = float(d[ix["gpu__time_duration.sum"]].replace(",", ""))
        res["kernels"][k] = {"name": d[ix["Kernel Name"]], "grid": d[ix["Grid Size"]], "dram_read_bytes": rd,
                             "dram_write_bytes": wr, "duration": dur, "duration_unit": units[ix["gpu__time_duration.sum"]]}
        total += rd + wr
    res["dram_bytes_per_launch"] = total
    res["algorithmic_bytes_per_launch"] = 16.0 * N * limbs * batch
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res)[:400])


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), sys.argv[5])
