"""The CUDA path vs the committed golden vectors that the reference's own code produced (tests/golden/*.npz)."""
import os

import numpy as np
import pytest
import torch

from deepgraphpose_b200 import synthetic

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load(name):
    with np.load(os.path.join(G, name)) as z:
        return {k: z[k] for k in z.files}


def test_softargmax_vs_reference_golden():
    """fp32 in, fp32 out: argmax_2d_from_cm of the reference (fitdgp_util.py:342-402) within 1e-3 scoremap px."""
    from deepgraphpose_b200 import fitdgp_util
    g = load("softargmax.npz")
    for tag in "abc":
        nj, gamma, gl = g[tag + "_par"]
        mu, sm = fitdgp_util.argmax_2d_from_cm(torch.from_numpy(g[tag + "_x"]).cuda(), int(nj), float(gamma), float(gl))
        assert np.abs(mu.cpu().numpy() - g[tag + "_mu"]).max() < 1e-3
        assert np.abs(sm.cpu().numpy() - g[tag + "_sm"]).max() < 1e-6 + 1e-4 * g[tag + "_sm"].max()


def test_estimate_pose_vs_reference_golden():
    """The whole path (ResNet-50 + head + soft-argmax + read-out) vs the reference's own estimate_pose output, at the
    north_star tolerances: coordinates <= 0.5 image px, likelihoods (sigmoid at the peak) <= 1e-2."""
    from deepgraphpose_b200 import eval as dgp_eval
    g = load("estimate_pose.npz")
    nj, wseed, vseed, H, W, T = [int(v) for v in g["meta"]]
    frames, _ = synthetic.make_video(T, H, W, nj, seed=vseed)
    cfg = {"num_joints": nj, "net_type": "resnet_50", "stride": 8.0}
    labels = dgp_eval.estimate_pose(cfg, "synthetic:%d" % wseed, frames, "/tmp", save_pose=False, batch=3)
    ex = np.abs(labels["x"] - g["x"]).max()
    ey = np.abs(labels["y"] - g["y"]).max()
    assert ex < 0.5 and ey < 0.5, (ex, ey)
    # the likelihood is the sigmoid at the arg-max pixel of the <=2x2 window around mu; wherever both paths pick the same
    # pixel it must agree to the scoremap tolerance, and they must pick the same pixel almost everywhere
    dl = np.abs(labels["likelihoods"] - g["likelihoods"])
    assert (dl < 1e-2).mean() >= 0.9 and np.median(dl) < 2e-3, (dl.max(), np.median(dl))


def test_posenet_vs_reference_golden():
    from deepgraphpose_b200.pose_net import PoseNet
    g = load("posenet.npz")
    nj, wseed, vseed, H, W = [int(v) for v in g["meta"]]
    frames, _ = synthetic.make_video(2, H, W, nj, seed=vseed)
    pn = PoseNet({"num_joints": nj, "location_refinement": True}, variables=synthetic.make_weights(nj, seed=wseed))
    out = pn.test(frames)
    prob = out["part_prob"].cpu().numpy()
    loc = out["locref"].cpu().numpy()
    for i in range(2):
        assert np.abs(prob[i:i + 1] - g["prob%d" % i]).max() < 1e-2
        assert np.abs(loc[i:i + 1] - g["locref%d" % i]).max() < 4e-3 * np.abs(g["locref%d" % i]).max()
