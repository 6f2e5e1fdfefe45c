"""Stand-alone roofline microbenchmarks of the bandwidth-class kernels at sizes that fill the GPU (inputs >> L2).

usage: python tools/bench_kernels.py   -> one JSON line per kernel on stdout
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from deepgraphpose_b200.engine import Engine  # noqa: E402


def timeit(fn, iters=10, warmup=3):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    hbm = peaks["hbm_gbs"]
    eng = Engine(4)
    out = []
    for (B, H, W, nj, tag) in [(4096, 94, 104, 4, "configs[1] scoremaps, 4096 frames"), (1024, 128, 160, 16, "configs[2] scoremaps, 1024 frames"),
                               (2048, 60, 80, 20, "configs[4] scoremaps, 2048 frames")]:
        x = torch.randn(B, H, W, nj, device="cuda") * 3
        gb = x.numel() * 4 / 1e9
        for wants, name in ((("mu", "peak", "lik"), "softargmax, estimate_pose read-out (mu, peak, lik)"),
                            (("mu", "peak", "lik", "dlc_peak", "dlc_pose"), "softargmax + DLC global peak")):
            ms = timeit(lambda: eng.softargmax(x, None, 1.0, 1.0, want=wants))
            out.append({"kernel": name, "workload": tag, "bytes": x.numel() * 4, "ms": ms,
                        "achieved_gbs": gb / (ms / 1e3), "peak_gbs": hbm, "frac": gb / (ms / 1e3) / hbm})
        del x
    T, nj = 4_000_000, 16
    mu = torch.rand(T, nj, 2, device="cuda") * 100
    edges = [(i, i + 1) for i in range(nj - 1)]
    buf = eng.potentials(mu, edges)
    ms = timeit(lambda: eng.potentials(mu, edges, out=buf))
    by = T * (8 * nj + 4 * (len(edges) + nj) + 4)
    out.append({"kernel": "potentials", "workload": "4M frames, 16 joints, chain", "bytes": by, "ms": ms,
                "achieved_gbs": by / 1e9 / (ms / 1e3), "peak_gbs": hbm, "frac": by / 1e9 / (ms / 1e3) / hbm})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
