#!/usr/bin/env python
"""bench.py -- frames/sec of the DGP scoremap-and-graph hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path (one rank per GPU under torchrun)
  python bench.py --impl reference --steps K --warmup W    # the reference's CPU path (oracle port) on the host cores

A "step" is one pass of the hot path over one batch of synthetic frames of BASELINE.json configs[1]
(747x832 RGB, 4 bodyparts, chain skeleton): ResNet-50 OS16 scoremap net + part_pred deconv head + fused
soft-argmax / peak / likelihood + skeleton & temporal potentials (+ the one-frame halo exchange when N > 1).
Prints ONE JSON line (rank 0).
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

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

H, W, NJ = 747, 832, 4
METRIC = "frames/sec (ResNet-50 scoremap + DGP soft-argmax/potentials)"
WORKLOAD = "configs[1]: synthetic 4-bodypart reaching video 747x832, estimate_pose inference (part_pred head, chain skeleton)"


CONFIGS = {  # BASELINE.json configs -> (H, W, num_joints, skeleton, label)
    "b": (747, 832, 4, "chain", WORKLOAD),
    "c": (1024, 1280, 16, "chain", "configs[2]: synthetic 1280x1024 video, 16 bodyparts with chain skeleton, frame-sharded estimate_pose inference"),
    "e": (480, 640, 20, "dense", "configs[4]: synthetic 640x480 videos x 20 bodyparts with dense skeleton (190 edges), estimate_pose inference"),
}


def conv_flops_per_frame(h, w, nj, locref=False):
    """Algorithmic forward FLOPs (2*MACs) of the conv + deconv layers (SURVEY.md 8d / Appendix B closed form)."""
    c2 = lambda v: -(-v // 2)
    total = 0
    hh, ww = c2(h), c2(w)
    total += 2 * hh * ww * 147 * 64
    hh, ww = c2(hh), c2(ww)
    cin = 64
    blocks = ((64, 3, 2), (128, 4, 2), (256, 6, 1), (512, 3, 1))  # stride as executed at output_stride 16
    for base, units, bstride in blocks:
        for u in range(units):
            s = bstride if u == units - 1 else 1
            depth = 4 * base
            ho, wo = c2(hh) if s == 2 else hh, c2(ww) if s == 2 else ww
            if cin != depth:
                total += 2 * hh * ww * cin * depth
            total += 2 * hh * ww * cin * base
            total += 2 * ho * wo * 9 * base * base
            total += 2 * ho * wo * base * depth
            hh, ww, cin = ho, wo, depth
    heads = nj * (3 if locref else 1)
    total += 2 * hh * ww * 2048 * 9 * heads
    return total, (2 * hh, 2 * ww)


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clocks / throttle reasons sampled DURING the timed region (NVML every 5 ms; nvidia-smi as fallback)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index):
        self.index = index
        self.sm, self.bits = [], 0
        self.stop_flag = False
        self.active = False     # samples are kept only while the timed region runs
        self.ready = threading.Event()
        self.th = None
        self.max_mhz = None

    def _loop(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.ready.set()
            while not self.stop_flag:
                if self.active:
                    self.sm.append(float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)))
                    try:
                        self.bits |= int(pynvml.nvmlDeviceGetCurrentClocksEventReasons(h))
                    except Exception:
                        self.bits |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
                time.sleep(0.002)
        except Exception:
            self.ready.set()
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=10).stdout
                a, b = [float(x) for x in out.strip().split(",")]
                self.sm.append(a)
                self.max_mhz = b
            except Exception:
                pass

    def begin(self):
        """Call right before the timed region (the thread was started earlier so that NVML is already initialised)."""
        self.ready.wait(timeout=10)
        self.active = True

    def start(self):
        self.th = threading.Thread(target=self._loop, daemon=True)
        self.th.start()

    def stop(self):
        self.active = False
        self.stop_flag = True
        if self.th:
            self.th.join(timeout=5)
        reasons = sorted(n for b, n in self.REASONS.items() if self.bits & b)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "reasons": reasons, "samples": len(self.sm)}


def make_frame_pool(n_frames, seed=1234):
    """n_frames distinct synthetic frames (uint8, T,H,W,3): a short seeded clip, extended by cyclic shifts."""
    from deepgraphpose_b200 import synthetic
    base_n = min(n_frames, 16)
    base, _ = synthetic.make_video(base_n, H, W, NJ, seed=seed)
    out = np.empty((n_frames, H, W, 3), np.uint8)
    for i in range(n_frames):
        out[i] = np.roll(base[i % base_n], shift=(3 * (i // base_n), 5 * (i // base_n)), axis=(0, 1))
    return out


def cpu_reference_fps(n_frames, warmup=1, threads=None):
    """The reference's CPU path (oracle restatement; TF1.15 is not installable here) on the host cores, batch 1 like
    eval.py:328.  Returns (fps, cores, seconds)."""
    from deepgraphpose_b200 import synthetic
    from oracle import dgp_ops, pose_net
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    Wn = synthetic.make_weights(NJ, seed=0, location_refinement=False)
    Wt = {k: torch.from_numpy(v) for k, v in Wn.items()}
    frames = make_frame_pool(max(2, min(n_frames, 4)))
    edges = synthetic.chain_skeleton(NJ)
    S0 = dgp_ops.skeleton_matrix(edges, NJ)

    def one(i):
        x = torch.from_numpy(frames[i % len(frames)][None].astype(np.float32))
        with torch.no_grad():
            net = pose_net.extract_features(x, Wt)
            pred = pose_net.prediction_layer(net, Wt, "part_pred")
            mu, _ = dgp_ops.argmax_2d_from_cm(pred, NJ, 1, 1)
        dgp_ops.estimate_pose_readout(mu.numpy(), pred.numpy())
        dgp_ops.skeleton_distances(mu, S0)
        return mu

    for i in range(warmup):
        one(i)
    t0 = time.perf_counter()
    for i in range(n_frames):
        one(i)
    dt = time.perf_counter() - t0
    return n_frames / dt, threads, dt


def run_reference(args, rank):
    if rank != 0:
        return
    per_step = 2
    for _ in range(max(args.warmup, 1)):
        cpu_reference_fps(1, warmup=0)
    fps, cores, dt = cpu_reference_fps(per_step * args.steps, warmup=0)
    line = {
        "impl": "reference", "metric": METRIC, "value": fps, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": per_step, "batch": 1},
        "cpu_baseline": {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                         "sample": "%d frames, batch 1 as eval.py:328, CPU restatement of the reference TF1 path "
                                   "(TF1.15 not installable: py3.12, no network)" % (per_step * args.steps)},
        "e2e": {"value": fps, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=0, help="frames per step per GPU (0 = engine.suggest_batch)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the auxiliary training-step measurement (configs[3])")
    ap.add_argument("--config", default="b", choices=sorted(CONFIGS),
                    help="BASELINE.json workload: b = configs[1] (the headline, default), c = configs[2], e = configs[4]")
    ap.add_argument("--cpu-frames", type=int, default=12)
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    global H, W, NJ, WORKLOAD
    H, W, NJ, skeleton_kind, WORKLOAD = CONFIGS[args.config]
    if args.config != "b":
        args.no_train = True

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch.distributed as dist
    from deepgraphpose_b200 import sharding, synthetic
    from deepgraphpose_b200.engine import Engine
    from deepgraphpose_b200.eval import estimate_pose_frames

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    from deepgraphpose_b200.engine import suggest_batch
    B = args.batch if args.batch > 0 else suggest_batch(H, W, *((12, 24) if args.config == "c" else (24, 48)))
    eng = Engine(NJ, location_refinement=False, device=local_rank)
    eng.load_weights(synthetic.make_weights(NJ, seed=0, location_refinement=False))
    edges = synthetic.dense_skeleton(NJ) if skeleton_kind == "dense" else synthetic.chain_skeleton(NJ)
    flops_frame, (hs, ws) = conv_flops_per_frame(H, W, NJ, locref=False)

    # Distinct input batches, together larger than the 126 MB L2 (and every layer's activations are far larger still).
    n_pool = 4
    pool_host = make_frame_pool(n_pool * B, seed=1234 + rank)
    pool = [torch.from_numpy(pool_host[i * B:(i + 1) * B]).to(dev) for i in range(n_pool)]
    ws_vec = np.full(len(edges), 1000.0 / 60.0, np.float32)
    ws_max = np.full(len(edges), 1.2 * 80.0, np.float32)

    def step(i):
        logits, _ = eng.forward(pool[i % n_pool], want_locref=False)
        out = eng.softargmax(logits, None, 1.0, 1.0, want=("mu", "peak", "lik"))
        halo = sharding.exchange_halo(out["mu"][0]) if world > 1 else None
        pot = eng.potentials(out["mu"], edges, halo_next=halo, ws=ws_vec, ws_max=ws_max, wt_max=0.0)
        return out, pot

    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.begin()
    launches0 = eng.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for i in range(args.steps):
        step(i)
    e1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    launches = eng.launch_count() - launches0
    clocks = sampler.stop()
    if world > 1:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    # Second pass over the SAME K steps with a CUDA-event pair around every kernel launch (recorded by the handle on the
    # launching stream) for the per-kernel-family roofline numbers.  It is kept out of the headline pass because an
    # event record between two layers defeats their programmatic-dependent-launch overlap.
    eng.get_profile()
    eng.set_profiling(True)
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for i in range(args.steps):
        step(i)
    p1.record()
    torch.cuda.synchronize()
    ms_prof = p0.elapsed_time(p1)
    eng.set_profiling(False)
    prof = eng.get_profile()
    value = world * B * args.steps / (ms / 1e3)

    # ---- end to end through the public API with HOST buffers (H2D of every batch + D2H of its results inside)
    e2e_steps = max(2, min(args.steps, 10))
    host = torch.from_numpy(np.concatenate([pool_host] * ((e2e_steps * B + len(pool_host) - 1) // len(pool_host)))[: e2e_steps * B]).pin_memory()
    estimate_pose_frames(eng, host[: 2 * B], batch=B)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    res = estimate_pose_frames(eng, host, batch=B)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if world > 1:
        t = torch.tensor([dt], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dt = float(t.item())
    e2e_value = world * e2e_steps * B / dt

    # ---- configs[3] alongside: one data-parallel DGP training step per rank (fwd + bwd + all-reduce + clip/Momentum)
    train_line = None
    if not args.no_train:
        try:
            sys.path.insert(0, os.path.join(ROOT, "tools"))
            import bench_train
            train_line = bench_train.measure(rank, local_rank, world, 5, 3, 10, H, W, profile=False)
        except Exception as ex:  # the headline line must survive a failure of the auxiliary measurement
            train_line = {"error": repr(ex)}
    if rank == 0:
        peaks, peak_src = load_peaks()
        gemm_ms, gemm_n = prof["conv_gemm"]
        sa_ms, sa_n = prof["softargmax"]
        frames_timed = B * args.steps
        achieved_tf = flops_frame * frames_timed / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else None
        peak_tf = peaks.get("bf16_tflops_sustained", peaks.get("bf16_tflops"))
        sa_bytes = 4 * hs * ws * NJ * frames_timed
        traffic, traffic_note = None, "no ncu capture committed"
        tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
        if os.path.exists(tpath) and args.config == "b":
            with open(tpath) as f:
                tj = json.load(f)
            traffic = tj["traffic_bytes_per_frame"] * B / 54
            traffic_note = ("dram__bytes_read.sum + dram__bytes_write.sum per conv_gemm launch (mean of the 54 layers) scaled from the "
                            "committed ncu capture profiles/r01_launches.csv (B=%d there); algorithmic minimum %.0f MB/launch" %
                            (tj["batch"], tj["minimum_bytes_per_frame_bf16"] * B / 54 / 1e6))
        line = {
            "metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step_per_gpu": B, "frame": [H, W, 3], "num_joints": NJ,
                       "skeleton": skeleton_kind, "l2": "4 rotating input batches (%d MB) and per-layer activations (>600 MB/step) exceed the 126 MB L2" % (n_pool * B * H * W * 3 // 2 ** 20),
                       "batch_choice": "engine.suggest_batch: tile counts of the persistent GEMM grid land on multiples of the SM count",
                       "parallelism": "frame shards x%d, 1-frame halo all_gather" % world if world > 1 else "single GPU"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": "frames/s", "h2d_bytes_per_step": B * H * W * 3,
                    "d2h_bytes_per_step": B * NJ * (8 + 8 + 4), "timing": "wall clock around estimate_pose_frames (synchronous API), max over ranks",
                    "frames": e2e_steps * B * world},
            "gpu_launches": launches,
            "roofline": {"bound": "tensor", "kernel": "conv_gemm_kernel (tcgen05 implicit GEMM, %d launches)" % gemm_n,
                         "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s",
                         "frac": (achieved_tf / peak_tf) if achieved_tf else None, "traffic": traffic,
                         "traffic_note": traffic_note,
                         "peak_source": peak_src + ", bf16 sustained", "share_of_step": gemm_ms / ms_prof,
                         "timing": "CUDA events around each of the %d launches in a second pass over the same steps (%.3f ms/step with events, %.3f without)" % (gemm_n, ms_prof / args.steps, ms / args.steps),
                         "algorithmic_gflop_per_frame": flops_frame / 1e9},
            "roofline_aux": {"softargmax": {"bound": "hbm", "achieved": sa_bytes / (sa_ms / 1e3) / 1e9 if sa_ms > 0 else None,
                                            "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "launches": sa_n,
                                            "note": "partial+finalize pair timed together inside the step"},
                             "ms_by_kernel_family": {k: v[0] for k, v in prof.items()}},
            "finite": bool(np.isfinite(res["x"]).all()),
        }
        if train_line is not None:
            keep = ("metric", "value", "unit", "ms_per_step", "config", "gpu_launches", "loss_after", "finite", "e2e", "error")
            line["train_step"] = {k: train_line[k] for k in keep if k in train_line}
        if world == 1 and not args.no_cpu_baseline and args.config == "b":
            fps, cores, cdt = cpu_reference_fps(args.cpu_frames, warmup=1)
            line["cpu_baseline"] = {"value": fps, "unit": "frames/s", "cores": cores, "kind": "port",
                                    "sample": "%d frames of the same workload, batch 1 (eval.py:328), %.1f s; CPU restatement of "
                                              "the reference TF1 path (TF1.15 not installable)" % (args.cpu_frames, cdt)}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
