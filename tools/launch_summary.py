"""Summarise an ncu gpu__time_duration launch list (csv) of one bench step: per-launch us next to the layer it is."""
import csv
import sys

LAYERS = ["prep_s2d", "conv1(s2d 4x1,K256,N64)", "maxpool"]
blocks = (("b1", 64, 3), ("b2", 128, 4), ("b3", 256, 6), ("b4", 512, 3))
cin = 64
for name, base, units in blocks:
    for u in range(units):
        if cin != base * 4:
            LAYERS.append("%su%d shortcut %d->%d" % (name, u + 1, cin, base * 4))
        LAYERS.append("%su%d conv1 %d->%d" % (name, u + 1, cin, base))
        LAYERS.append("%su%d conv2 3x3 %d" % (name, u + 1, base))
        LAYERS.append("%su%d conv3 %d->%d +res" % (name, u + 1, base, base * 4))
        cin = base * 4
LAYERS += ["head GEMM", "col2im", "softargmax_partial", "softargmax_finalize", "(torch copy)", "potentials"]


def main(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    tot = 0.0
    for i, r in enumerate(rows):
        us = float(r["Metric Value"].replace(",", "")) / 1e3
        tot += us
        print("%3d %-34s %-28s %9.1f us" % (i, LAYERS[i] if i < len(LAYERS) else "", r["Kernel Name"].split("(")[0][-28:], us))
    print("total %.1f us" % tot)


if __name__ == "__main__":
    main(sys.argv[1])
