"""Markdown summary of an `ncu --set full` report: one block of headline metrics per captured kernel.

usage: python tools/ncu_summary.py report.ncu-rep [label ...] >> profiles/rNN_ncu_*.md
"""
import csv
import subprocess
import sys

METRICS = [
    "gpu__time_duration.sum",
    "dram__bytes_read.sum",
    "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__grid_size",
    "launch__block_size",
    "launch__registers_per_thread",
    "launch__shared_mem_per_block_dynamic",
]


def main():
    rep = sys.argv[1]
    labels = sys.argv[2:]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    ix = {h: i for i, h in enumerate(hdr)}
    for n, r in enumerate(data):
        name = r[ix["Kernel Name"]]
        short = name.split("(")[0].split("::")[-1]
        label = labels[n] if n < len(labels) else ""
        print("### %s%s\n" % (short, " -- " + label if label else ""))
        for m in METRICS:
            if m in ix:
                print("- %s = %s %s" % (m, r[ix[m]], units[ix[m]]))
        print()


if __name__ == "__main__":
    main()
