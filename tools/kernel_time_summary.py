"""Aggregate an ncu `--metrics gpu__time_duration.sum[,dram__bytes_*]` csv launch list by kernel name."""
import collections
import csv
import sys


def main(path, skip=0):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    per = collections.OrderedDict()
    ids = collections.OrderedDict()
    for r in rows:
        ids.setdefault(r["ID"], {})[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Kernel Name"], r.get("Metric Unit", ""))
    for i, (k, m) in enumerate(ids.items()):
        if i < skip:
            continue
        t, name, unit = m["gpu__time_duration.sum"]
        us = t / 1e3 if unit in ("ns", "nsecond") else t
        name = name.split("(")[0].split("<")[0][-40:]
        rd = m.get("dram__bytes_read.sum", (0, "", ""))
        wr = m.get("dram__bytes_write.sum", (0, "", ""))
        scale = lambda v: v[0] * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(v[2], 1)
        d = per.setdefault(name, [0, 0.0, 0.0])
        d[0] += 1
        d[1] += us
        d[2] += scale(rd) + scale(wr)
    tot = sum(v[1] for v in per.values())
    for name, (n, us, by) in sorted(per.items(), key=lambda kv: -kv[1][1]):
        print("%-42s %5d launches %10.1f us %5.1f%%  %8.1f MB dram  %6.0f GB/s" % (name, n, us, 100 * us / tot, by / 1e6, by / us / 1e3 if us else 0))
    print("total %.1f us" % tot)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
