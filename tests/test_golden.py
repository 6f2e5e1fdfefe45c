"""The oracle vs golden vectors produced by the reference's OWN code (tests/golden/make_golden.py runs the unmodified
reference modules from /root/reference on oracle/tf1_shim.py).  This pins the oracle's composition of the TF ops."""
import os

import numpy as np
import torch

from deepgraphpose_b200 import synthetic
from oracle import dgp_loss, dgp_ops, pose_net

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with np.load(os.path.join(G, name)) as z:
        return {k: z[k] for k in z.files}


def test_softargmax_golden():
    g = load("softargmax.npz")
    for tag in "abc":
        nj, gamma, gl = g[tag + "_par"]
        mu, sm = dgp_ops.argmax_2d_from_cm(torch.from_numpy(g[tag + "_x"]), int(nj), float(gamma), float(gl))
        assert np.abs(mu.numpy() - g[tag + "_mu"]).max() < 1e-5
        assert np.abs(sm.numpy() - g[tag + "_sm"]).max() < 1e-6


def test_posenet_golden():
    g = load("posenet.npz")
    nj, wseed, vseed, H, W = [int(v) for v in g["meta"]]
    Wt = {k: torch.from_numpy(v) for k, v in synthetic.make_weights(nj, seed=wseed).items()}
    frames, _ = synthetic.make_video(2, H, W, nj, seed=vseed)
    for i in range(2):
        with torch.no_grad():
            out = pose_net.test(torch.from_numpy(frames[i][None].astype(np.float32)), Wt)
        assert np.abs(out["part_prob"].numpy() - g["prob%d" % i]).max() < 1e-5
        assert np.abs(out["locref"].numpy() - g["locref%d" % i]).max() < 1e-4
        scm, loc = pose_net.extract_cnn_output(out["part_prob"].numpy(), out["locref"].numpy())
        pose, _ = pose_net.argmax_pose_predict(scm, loc, 8.0)
        assert np.abs(pose - g["pose_np%d" % i]).max() < 1e-4
        # PoseNet.inference (TF graph version) emits (row-derived, col-derived, likelihood): same numbers, (y, x) order
        assert np.abs(g["pose_tf%d" % i][:, [1, 0, 2]] - g["pose_np%d" % i]).max() < 1e-3


def test_estimate_pose_golden():
    g = load("estimate_pose.npz")
    nj, wseed, vseed, H, W, T = [int(v) for v in g["meta"]]
    Wt = {k: torch.from_numpy(v) for k, v in synthetic.make_weights(nj, seed=wseed, location_refinement=False).items()}
    frames, _ = synthetic.make_video(T, H, W, nj, seed=vseed)
    markers = np.zeros((T, nj, 2))
    lik = np.zeros((T, nj))
    for t in range(T):
        with torch.no_grad():
            net = pose_net.extract_features(torch.from_numpy(frames[t][None].astype(np.float32)), Wt)
            pred = pose_net.prediction_layer(net, Wt, "part_pred")
            mu, _ = dgp_ops.argmax_2d_from_cm(pred, nj, 1, 1)
        markers[t], _, lik[t] = dgp_ops.estimate_pose_readout(mu.numpy(), pred.numpy())
    x, y = dgp_ops.estimate_pose_xy(markers)
    assert np.abs(x - g["x"]).max() < 1e-3 and np.abs(y - g["y"]).max() < 1e-3
    assert np.abs(lik - g["likelihoods"]).max() < 1e-5


def test_dgp_loss_golden():
    g = load("dgp_loss.npz")
    nj, wseed, vseed, H, W, nt = [int(v) for v in g["meta"]]
    Wt = {k: torch.from_numpy(v) for k, v in synthetic.make_weights(nj, seed=wseed).items()}
    frames, _ = synthetic.make_video(nt, H, W, nj, seed=vseed)
    with torch.no_grad():
        net = pose_net.extract_features(torch.from_numpy(frames.astype(np.float32)), Wt)
        pred = pose_net.prediction_layer(net, Wt, "part_pred")
        loc = pose_net.prediction_layer(net, Wt, "locref_pred")
    nx, ny = pred.shape[1], pred.shape[2]
    xg, yg = np.meshgrid(np.linspace(0, nx - 1, nx), np.linspace(0, ny - 1, ny))
    for tag, wt in (("wt0", 0.0), ("wt1", 1.0)):
        cfg = dgp_loss.default_dgp_cfg(wt=wt)
        ws, ws_max = dgp_loss.spatial_clique_params(g["labels"], g["S0"], cfg)
        batch = {"targets": g["labels"], "visible_marker_pl": g["visible_marker"], "hidden_marker_pl": g["hidden_marker"],
                 "visible_marker_in_targets_pl": g["vis_in_targets"], "nt_batch_pl": nt, "locref_map": g["locref_map"],
                 "locref_mask": g["locref_mask"], "alpha_tf": np.array([xg, yg]).swapaxes(1, 2),
                 "vector_field_tf": g["vector_field"], "wt_batch_pl": np.ones(nt - 1) * wt,
                 "wt_batch_mask_pl": g["wt_batch_mask"]}
        with torch.no_grad():
            loss, total, total_vis = dgp_loss.dgp_loss_from_heads(pred, loc, batch, cfg, g["S0"], ws, ws_max, 120, 10)
        for k, v in loss.items():
            ref = float(g["%s_%s" % (tag, k)])
            assert abs(float(v) - ref) <= 1e-5 * max(1.0, abs(ref)), (tag, k, float(v), ref)
        assert abs(float(total_vis) - float(g[tag + "_total_loss_visible"])) < 1e-5
