"""One line per profiled launch from an `ncu --page raw --csv` export (ncu --set full)."""
import csv
import sys

COLS = [("gpu__time_duration.sum", "us", 1.0), ("dram__bytes_read.sum", "rdMB", 1.0), ("dram__bytes_write.sum", "wrMB", 1.0),
        ("launch__registers_per_thread", "regs", 1.0), ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps%", 1.0),
        ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", 1.0),
        ("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "alu%", 1.0),
        ("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "fma%", 1.0),
        ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%", 1.0),
        ("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "fp64%", 1.0),
        ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "l2%", 1.0),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "shortsb", 1.0),
        ("smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio", "dispat", 1.0),
        ("smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "barrier", 1.0),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "longsb", 1.0),
        ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "mathpt", 1.0),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "wait", 1.0),
        ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "notsel", 1.0),
        ("smsp__inst_executed.sum", "Minst", 1e-6)]


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print("%-34s %-14s " % ("kernel", "grid") + " ".join("%7s" % c[1] for c in COLS))
    for d in data:
        name = d[idx["Kernel Name"]].replace("void <unnamed>::", "").replace("<unnamed>::", "").split("(")[0]
        vals = []
        for key, _, sc in COLS:
            v = d[idx[key]].replace(",", "") if key in idx else "nan"
            u = units[idx[key]] if key in idx else ""
            try:
                f = float(v) * sc
                if u == "Gbyte":
                    f *= 1e3
                elif u == "Kbyte":
                    f *= 1e-3
                elif u == "byte":
                    f *= 1e-6
                elif u == "ms":
                    f *= 1e3
                elif u == "ns":
                    f *= 1e-3
                vals.append("%7.1f" % f)
            except ValueError:
                vals.append("%7s" % v[:7])
        print("%-34s %-14s " % (name[:34], d[idx["Grid Size"]].replace(" ", "")) + " ".join(vals))


if __name__ == "__main__":
    main(sys.argv[1])
