"""Parity of the two 16-bit storage modes against the fp32 oracle at every BASELINE.json inference shape, on both synthetic
weight sets (flat random-init, trained-like), plus the training-step loss.  Prints one JSON line per case and writes
gpurun_out/precision.jsonl.  north_star tolerances: sigmoid <= 1e-2 max-abs, soft-argmax <= 0.5 image px, loss <= 1e-3 rel.

    python tools/diag_precision.py [--quick]
"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from deepgraphpose_b200 import fitdgp, synthetic  # noqa: E402
from deepgraphpose_b200.engine import Engine  # noqa: E402
from oracle import dgp_loss as oracle_loss  # noqa: E402
from oracle import dgp_ops, pose_net  # noqa: E402

SHAPES = [("a", 747, 832, 5, True), ("b", 747, 832, 4, False), ("c", 1024, 1280, 16, False), ("e", 480, 640, 20, True)]


def inference_case(tag, H, W, nj, locref, trained, out):
    Wn = synthetic.make_weights(nj, seed=0, location_refinement=locref, trained_like=trained)
    Wt = {k: torch.from_numpy(v) for k, v in Wn.items()}
    frames, _ = synthetic.make_video(1, H, W, nj, seed=7)
    t0 = time.time()
    with torch.no_grad():
        net = pose_net.extract_features(torch.from_numpy(frames.astype(np.float32)), Wt)
        pred = pose_net.prediction_layer(net, Wt, "part_pred")
        loc = pose_net.prediction_layer(net, Wt, "locref_pred") if locref else None
        mu_ref, _ = dgp_ops.argmax_2d_from_cm(pred, nj, 1.0, 1.0)
    _, pk_ref, lik_ref = dgp_ops.estimate_pose_readout(mu_ref.numpy(), pred.numpy())
    cpu_s = time.time() - t0
    for prec in ("fp16", "bf16"):
        eng = Engine(nj, location_refinement=locref, precision=prec)
        eng.load_weights(Wn)
        logits, lr = eng.forward(torch.from_numpy(frames).cuda())
        o = eng.softargmax(logits, lr)
        torch.cuda.synchronize()
        row = {"case": tag, "frame": [H, W], "nj": nj, "weights": "trained_like" if trained else "flat", "precision": prec,
               "logit_std": float(pred.std()), "logit_max": float(pred.abs().max()),
               "logit_rel": float((logits.cpu() - pred).abs().max() / pred.abs().max()),
               "sigmoid_maxabs": float((torch.sigmoid(logits.cpu()) - torch.sigmoid(pred)).abs().max()),
               "mu_err_image_px": float((o["mu"].cpu() - mu_ref).abs().max() * 8.0),
               "peaks_equal": float((o["peak"][0].cpu().numpy() == pk_ref).all(axis=-1).mean()),
               "lik_maxabs": float(np.abs(o["lik"][0].cpu().numpy() - lik_ref).max()),
               "cpu_oracle_s": cpu_s}
        if locref:
            row["locref_rel"] = float((lr.cpu() - loc).abs().max() / loc.abs().max())
        print(json.dumps(row), flush=True)
        out.append(row)
        eng.close()


def training_case(nt, Hin, Win, nj, trained, out):
    rng = np.random.default_rng(3)
    from test_gpu_loss import make_batch
    Wn = synthetic.make_weights(nj, seed=3, trained_like=trained)
    frames, _ = synthetic.make_video(nt, Hin, Win, nj, seed=11)
    H, Wd = 2 * -(-Hin // 16), 2 * -(-Win // 16)
    labels, batch = make_batch(rng, nt, H, Wd, nj, [0, 2], ((0, 1),))
    edges = synthetic.chain_skeleton(nj)
    S0 = dgp_ops.skeleton_matrix(edges, nj)
    cfg = oracle_loss.default_dgp_cfg(gm2=1, gm3=3, wt=0.0)
    ws, ws_max = oracle_loss.spatial_clique_params(labels, S0, cfg)
    ws_max = ws_max * 0.3
    Wt = {k: torch.from_numpy(v) for k, v in Wn.items()}
    with torch.no_grad():
        heads = pose_net.get_net(torch.from_numpy(frames.astype(np.float32)), Wt, True)
        ref, _, _ = oracle_loss.dgp_loss_from_heads(heads["part_pred"], heads["locref"], batch, cfg, S0, ws, ws_max, 200, 20)
    for prec in ("fp16", "bf16"):
        eng = Engine(nj, precision=prec)
        eng.load_weights(Wn)
        got = fitdgp.train_forward_backward(eng, torch.from_numpy(frames).cuda(), batch, cfg, edges, ws, ws_max, 200, 20)
        row = {"case": "train", "frame": [Hin, Win], "nt": nt, "weights": "trained_like" if trained else "flat",
               "precision": prec}
        for k, v in ref.items():
            row[k + "_rel"] = abs(float(got[k]) - float(v)) / max(abs(float(v)), 1e-12)
        print(json.dumps(row), flush=True)
        out.append(row)
        eng.close()


def main():
    quick = "--quick" in sys.argv
    torch.set_num_threads(os.cpu_count())
    out = []
    for trained in (False, True):
        for tag, H, W, nj, locref in SHAPES:
            if quick and tag in ("a", "c"):
                continue
            inference_case(tag, H, W, nj, locref, trained, out)
        training_case(3, 64, 96, 4, trained, out)
        training_case(4, 235, 301, 4, trained, out)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "precision.jsonl"), "w") as f:
        for r in out:
            f.write(json.dumps(r) + "\n")


if __name__ == "__main__":
    main()
