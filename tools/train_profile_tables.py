"""profiles/rNN_train_*.md from an ncu launch list of tools/bench_train.py (last step of the capture):
per-kernel-family summary and the per-layer wgrad / dgrad tables."""
import collections
import csv
import sys

SCALE = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1.0}


def load(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    ids = collections.OrderedDict()
    for r in csv.DictReader(lines):
        ids.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r.get("Grid Size", "")})[r["Metric Name"]] = (
            float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    out = []
    for m in ids.values():
        t, u = m["gpu__time_duration.sum"]
        us = t / 1e3 if u in ("ns", "nsecond") else t
        by = sum(m[k][0] * SCALE.get(m[k][1], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in m)
        out.append((m["name"].split("(")[0].split("<")[0].split("::")[-1], us, by))
    return out


def layer_list(H=747, W=832, nt=10, nj=4):
    c2 = lambda v: -(-v // 2)
    hh, ww = c2(c2(H)), c2(c2(W))
    units, cin = [], 64
    for name, base, n, bs in (("b1", 64, 3, 2), ("b2", 128, 4, 2), ("b3", 256, 6, 1), ("b4", 512, 3, 1)):
        for u in range(n):
            s = bs if u == n - 1 else 1
            ho, wo = (c2(hh), c2(ww)) if s == 2 else (hh, ww)
            units.append((name + "u%d" % (u + 1), cin, base, 4 * base, hh, ww, ho, wo, cin != 4 * base, s))
            hh, ww, cin = ho, wo, 4 * base
    order = [("heads 2048->%d" % (9 * 3 * nj), 2 * hh * ww * 2048 * 9 * 3 * nj)]
    for (nm, cin, base, depth, h, w, ho, wo, proj, s) in reversed(units):
        order.append((nm + " conv3 %d->%d" % (base, depth), 2 * ho * wo * base * depth))
        order.append((nm + " conv2 3x3 %d%s" % (base, " s2" if s == 2 else ""), 2 * ho * wo * 9 * base * base))
        order.append((nm + " conv1 %d->%d" % (cin, base), 2 * h * w * cin * base))
        if proj:
            order.append((nm + " shortcut %d->%d" % (cin, depth), 2 * h * w * cin * depth))
    order.append(("root conv1 (s2d K=256)", 2 * c2(H) * c2(W) * 256 * 64))
    return [(n, f * nt) for n, f in order]


def main(path):
    L = load(path)
    starts = [i for i, (n, _, _) in enumerate(L) if n == "prep_s2d_kernel"]
    # the capture may end in the middle of a step: take the last COMPLETE one (between two consecutive prep launches)
    step = L[starts[-2]:starts[-1]] if len(starts) >= 2 else L[starts[-1]:]
    # the profiled pass is the last one; cut at its end (everything after the last build_head_dgrad_w_kernel belongs to teardown)
    fam = collections.OrderedDict()
    for n, us, by in step:
        d = fam.setdefault(n, [0, 0.0, 0.0])
        d[0] += 1; d[1] += us; d[2] += by
    tot = sum(v[1] for v in fam.values())
    print("## Kernels of one training step (ncu gpu__time_duration, serialised, cold cache; 747x832, nt = 10, nj = 4 + locref)\n")
    print("| kernel | launches | us | share | DRAM MB (r+w) | GB/s |\n|---|---|---|---|---|---|")
    for n, (k, us, by) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
        print("| %s | %d | %.1f | %.1f %% | %.1f | %.0f |" % (n, k, us, 100 * us / tot, by / 1e6, by / us / 1e3 if us else 0))
    print("| **total** | %d | **%.1f** | | | |\n" % (sum(v[0] for v in fam.values()), tot))
    wg = [x for x in step if x[0] == "wgrad_gemm_kernel"]
    rd = [x for x in step if x[0] == "wgrad_reduce_kernel"]
    order = layer_list()
    print("## Weight-gradient GEMM per layer (backward order)\n")
    print("| layer | wgrad GEMM us | TFLOP/s (algorithmic) | DRAM MB | DRAM GB/s | reduce us |\n|---|---|---|---|---|---|")
    tg = tf = 0.0
    for (nm, fl), g, r in zip(order, wg, rd):
        print("| %s | %.1f | %.0f | %.0f | %.0f | %.1f |" % (nm, g[1], fl / g[1] / 1e6, g[2] / 1e6, g[2] / g[1] / 1e3, r[1]))
        tg += g[1]; tf += fl
    print("| **all %d layers** | **%.1f** | **%.0f** | | | %.1f |" % (len(wg), tg, tf / tg / 1e6, sum(r[1] for r in rd)))


if __name__ == "__main__":
    main(sys.argv[1])
