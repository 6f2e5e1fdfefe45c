"""In-step duration of every launch of the inference step (CUDA events recorded by the handle on the launching stream while
the step runs back to back at its real clocks / power state), next to the SM clock and board power sampled during the run.

    python tools/instep_layers.py [--steps 200] [--precision fp16|bf16] [--config b|c|e] [--ncu launches.csv]

Prints a markdown table (mean us per launch index over the steps) and one JSON line with the totals.  With --ncu the
isolated (cold-cache, serialised, boost-clock) duration of the same launch from an ncu launch list is printed beside it.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
import bench  # noqa: E402
import layer_table  # noqa: E402


class SmiSampler:
    def __init__(self):
        self.rows, self.stop_flag = [], False

    def _loop(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(0)
            while not self.stop_flag:
                self.rows.append((time.perf_counter(), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                                  pynvml.nvmlDeviceGetPowerUsage(h) / 1e3))
                time.sleep(0.005)
        except Exception as ex:  # noqa: BLE001
            self.rows.append((time.perf_counter(), -1, -1.0))
            print("nvml unavailable: %r" % (ex,), file=sys.stderr)

    def start(self):
        self.th = threading.Thread(target=self._loop, daemon=True)
        self.th.start()

    def stop(self):
        self.stop_flag = True
        self.th.join(timeout=5)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--precision", default="fp16")
    ap.add_argument("--config", default="b")
    ap.add_argument("--batch", type=int, default=0)
    ap.add_argument("--ncu", default=None)
    args = ap.parse_args()
    bench.H, bench.W, bench.NJ, _, bench.WORKLOAD = bench.CONFIGS[args.config]
    layer_table.H, layer_table.W, layer_table.NJ = bench.H, bench.W, bench.NJ
    wl = bench.Workload(args.config, 0, 0, args.precision, args.batch)
    for i in range(5):
        wl.step(i, 1)
    torch.cuda.synchronize()
    # 1. plain timed run (PDL chaining intact), with clock / power samples
    smp = SmiSampler()
    smp.start()
    time.sleep(0.05)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_begin = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        wl.step(i, 1)
    e1.record()
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    ms_plain = e0.elapsed_time(e1) / args.steps
    # 2. the same with per-launch events
    wl.eng.get_profile()
    wl.eng.set_profiling(True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(args.steps):
        wl.step(i, 1)
    p1.record()
    torch.cuda.synchronize()
    ms_prof = p0.elapsed_time(p1) / args.steps
    wl.eng.set_profiling(False)
    recs = wl.eng.get_profile_records(args.steps * 80)
    smp.stop()
    per_step = len(recs) // args.steps
    kinds = [recs[i][0] for i in range(per_step)]
    ms = np.array([r[1] for r in recs[:per_step * args.steps]]).reshape(args.steps, per_step)
    mean_us, min_us = 1e3 * ms.mean(0), 1e3 * ms.min(0)
    iso = None
    if args.ncu:
        step = layer_table.read_step(args.ncu)
        iso = [l["gpu__time_duration.sum"] / 1e3 for l in step if "conv_gemm" in l["name"]][:54]
    names = iter(n for n, _, _ in layer_table.layers(wl.B))
    flops = iter(f for _, f, _ in layer_table.layers(wl.B))
    print("| # | kernel / layer | in-step us (mean) | min | isolated us (ncu) | in-step TFLOP/s |")
    print("|---|---|---|---|---|---|")
    gi = 0
    tot_gemm = 0.0
    for i, k in enumerate(kinds):
        if k == "conv_gemm":
            nm, fl = next(names), next(flops)
            tot_gemm += mean_us[i]
            print("| %d | %s | %.1f | %.1f | %s | %.0f |" % (i, nm, mean_us[i], min_us[i], ("%.1f" % iso[gi]) if iso and gi < len(iso) else "", fl / mean_us[i] / 1e6))
            gi += 1
        else:
            print("| %d | %s | %.1f | %.1f | | |" % (i, k, mean_us[i], min_us[i]))
    rows = [r for r in smp.rows if t_begin <= r[0] <= t_end and r[1] > 0]
    line = {"config": args.config, "precision": args.precision, "batch": wl.B, "steps": args.steps, "ms_per_step": ms_plain,
            "ms_per_step_with_events": ms_prof, "gemm_us_in_step": tot_gemm, "launches_per_step": per_step,
            "gemm_tflops_in_step": wl.flops_frame * wl.B / tot_gemm / 1e6,
            "sm_mhz_median": float(np.median([r[1] for r in rows])) if rows else None,
            "sm_mhz_min": float(np.min([r[1] for r in rows])) if rows else None,
            "power_w_mean": float(np.mean([r[2] for r in rows])) if rows else None,
            "power_w_max": float(np.max([r[2] for r in rows])) if rows else None, "samples": len(rows)}
    print()
    print(json.dumps(line))
    wl.close()


if __name__ == "__main__":
    main()
