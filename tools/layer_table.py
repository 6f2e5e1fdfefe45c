"""Per-layer table for profiles/: ncu launch time of each GEMM layer of one bench step next to its roofline bounds.

usage: python tools/layer_table.py gpurun_out/launches_rXX.csv [batch] > profiles/rNN_layers.md
bounds per layer (bf16 activations, fp32 logits): t_hbm = (A read + residual read + output write) / 6464 GB/s,
t_tc = algorithmic FLOPs / 1414 TFLOP/s (MEASURED_PEAKS.json), bound = max of the two.
"""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
H, W, NJ = 747, 832, 4


def layers(B):
    c2 = lambda v: -(-v // 2)
    out = []
    h1, w1 = c2(H), c2(W)
    m1 = B * h1 * w1
    out.append(("conv1 7x7s2 (s2d 4x1, K=256)", 2 * m1 * 147 * 64, B * (h1 + 3) * (w1 + 3) * 32 + m1 * 64 * 2))
    hh, ww = c2(h1), c2(w1)
    cin = 64
    for name, base, units, bstride in (("b1", 64, 3, 2), ("b2", 128, 4, 2), ("b3", 256, 6, 1), ("b4", 512, 3, 1)):
        for u in range(units):
            s = bstride if u == units - 1 else 1
            depth = 4 * base
            ho, wo = (c2(hh), c2(ww)) if s == 2 else (hh, ww)
            m, mo = B * hh * ww, B * ho * wo
            if cin != depth:
                out.append(("%su%d shortcut 1x1 %d->%d" % (name, u + 1, cin, depth), 2 * m * cin * depth, m * (cin + depth) * 2))
            out.append(("%su%d conv1 1x1 %d->%d" % (name, u + 1, cin, base), 2 * m * cin * base, m * (cin + base) * 2))
            out.append(("%su%d conv2 3x3 %d s%d" % (name, u + 1, base, s), 2 * mo * 9 * base * base, (m + mo) * base * 2))
            out.append(("%su%d conv3 1x1 %d->%d +res" % (name, u + 1, base, depth), 2 * mo * base * depth,
                        mo * (base + 2 * depth) * 2))
            hh, ww, cin = ho, wo, depth
    npad = 48 if NJ == 4 else 9 * NJ
    out.append(("heads deconv-as-GEMM 2048->%d" % (9 * NJ), 2 * B * hh * ww * 2048 * 9 * NJ, B * hh * ww * (2048 * 2 + npad * 4)))
    return out


def read_step(path):
    """One full step from an ncu launch list (csv, any of gpu__time_duration.sum / dram__bytes_*.sum per launch):
    launches are grouped by ID and rotated so that the list starts at prep_s2d_kernel."""
    launches = {}
    order = []
    for r in csv.DictReader(l for l in open(path) if not l.startswith("==")):
        if r["ID"] not in launches:
            launches[r["ID"]] = {"name": r["Kernel Name"]}
            order.append(r["ID"])
        launches[r["ID"]][r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
    seq = [launches[i] for i in order]
    starts = [i for i, l in enumerate(seq) if "prep_s2d" in l["name"]]
    if not starts:
        return seq
    s0 = starts[0]
    step_len = (starts[1] - s0) if len(starts) > 1 else None
    if step_len is None:
        # single prep in the window: the step is [s0, end) + [first launches before s0 that complete it]
        tail = seq[s0:]
        need = 60 - len(tail)
        head = seq[max(0, s0 - 60 + len(tail) + (60 - len(tail)) - need):][:0]
        return tail + seq[len(tail) - 60 + s0: s0] if need > 0 else tail[:60]
    return seq[s0:s0 + step_len]


def main():
    path = sys.argv[1]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm, tc = peaks["hbm_gbs"] * 1e9, peaks["bf16_tflops_sustained"] * 1e12
    step = read_step(path)
    T = "gpu__time_duration.sum"
    gem = [l for l in step if "conv_gemm" in l["name"]][:54]
    other = [(l["name"].split("(")[0].split("::")[-1], l[T] / 1e3) for l in step if "conv_gemm" not in l["name"]]
    have_dram = all("dram__bytes_read.sum" in l for l in gem)
    print("| layer | us (ncu, B=%d) | GFLOP | t_tc us | MB (algorithmic) | MB (ncu dram r+w) | t_hbm us | bound | x over bound | TFLOP/s |" % B)
    print("|---|---|---|---|---|---|---|---|---|---|")
    tot, totb, totd = 0.0, 0.0, 0.0
    for (name, fl, by), l in zip(layers(B), gem):
        us = l[T] / 1e3
        ttc, thbm = fl / tc * 1e6, by / hbm * 1e6
        bound = max(ttc, thbm)
        tot += us
        totb += bound
        dram = (l["dram__bytes_read.sum"] + l["dram__bytes_write.sum"]) if have_dram else float("nan")
        totd += dram
        print("| %s | %.1f | %.1f | %.1f | %.0f | %.0f | %.1f | %s | %.2f | %.0f |" % (
            name, us, fl / 1e9, ttc, by / 1e6, dram / 1e6, thbm, "TC" if ttc >= thbm else "HBM", us / bound, fl / us / 1e6))
    print("| **all GEMM layers** | **%.0f** | | | | **%.0f** | | | **%.2f** (sum of bounds %.0f us) | |" % (tot, totd / 1e6, tot / totb, totb))
    print()
    print("Other kernels of the step (ncu, us): " + ", ".join("%s %.1f" % o for o in other))
    print()
    print("Step total (ncu, serialised, cold cache): %.0f us; GEMM share %.1f %%." % (sum(l[T] for l in step) / 1e3, 100 * tot * 1e3 / sum(l[T] for l in step)))
    if have_dram:
        json.dump({"kernel": "conv_gemm_kernel", "launches": len(gem), "batch": B,
                   "traffic_bytes_per_launch": totd / len(gem), "traffic_bytes_per_frame": totd / B,
                   "source": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none "
                             "-s 183 -c 61 python bench.py --steps 1 --warmup 3 (tools/profile_round.sh)"},
                  open(os.path.join(ROOT, "gpurun_out", "traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
