"""Top sampled SASS lines of an `ncu --set full --import-source on` report (warp-stall sampling), plus headline metrics.

usage: python tools/ncu_hot.py report.ncu-rep [n_lines]
"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
        "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_l1tex2xbar_write_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__grid_size",
        "launch__cluster_size", "sm__cycles_elapsed.avg", "smsp__cycles_active.avg"]


def main():
    rep = sys.argv[1]
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    for h, u, v in zip(rows[0], rows[1], rows[2]):
        if h in WANT:
            print("%s = %s %s" % (h, v, u))
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(src.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    tot = sum(int(r[ix["# Samples"]] or 0) for r in data)
    print("total samples", tot)
    for r in sorted(data, key=lambda r: -int(r[ix["# Samples"]] or 0))[:n]:
        print("%6s %9s %s %s" % (r[ix["# Samples"]], r[ix["Instructions Executed"]], r[ix["Address"]][-5:], r[ix["Source"]][:100]))


if __name__ == "__main__":
    main()
