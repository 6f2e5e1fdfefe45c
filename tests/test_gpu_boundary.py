"""Every reference-named entry point of the boundary imported from deepgraphpose_b200 and diffed against
golden vectors the reference's own functions produced (tests/golden/make_golden.py -> posenet.npz, boundary.npz)."""
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


@pytest.fixture(scope="module")
def pn():
    from deepgraphpose_b200.pose_net import PoseNet
    g = load("boundary.npz")
    nj, wseed, vseed, H, W = [int(v) for v in g["meta"]]
    net = PoseNet({"num_joints": nj, "location_refinement": True, "net_type": "resnet_50", "stride": 8.0,
                   "locref_stdev": 7.2801, "deconvolutionstride": 2, "output_stride": 16},
                  variables=synthetic.make_weights(nj, seed=wseed))
    frames, _ = synthetic.make_video(2, H, W, nj, seed=vseed)
    yield net, g, frames, nj
    net.engine.close()


def rel(a, b):
    return float(np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(np.asarray(b)).max())


def test_posenet_inference_returns_y_x_likelihood_like_the_reference():
    """PoseNet.inference (pose_net.py:92-128, batch_size 1) emits (row-derived, col-derived, likelihood) = (y, x, lik):
    against the reference's own TF-graph output pose_tf* of posenet.npz."""
    from deepgraphpose_b200.pose_net import PoseNet
    g = load("posenet.npz")
    nj, wseed, vseed, H, W = [int(v) for v in g["meta"]]
    frames, _ = synthetic.make_video(2, H, W, nj, seed=vseed)
    net = PoseNet({"num_joints": nj, "location_refinement": True}, variables=synthetic.make_weights(nj, seed=wseed))
    for i in range(2):
        pose = net.inference(frames[i:i + 1])["pose"].cpu().numpy()
        ref = g["pose_tf%d" % i]
        assert pose.shape == ref.shape == (nj, 3)
        assert np.abs(pose[:, :2] - ref[:, :2]).max() < 0.5, (pose, ref)       # image px
        assert np.abs(pose[:, 2] - ref[:, 2]).max() < 1e-2
        # and it is NOT the (x, y) order of argmax_pose_predict: the numpy version holds the swapped columns
        assert np.abs(pose[:, [1, 0, 2]] - g["pose_np%d" % i]).max() < 0.5
    net.engine.close()


def test_extract_features_vs_reference_golden(pn):
    net_, g, frames, nj = pn
    net, end_points = net_.extract_features(frames[:1])
    assert tuple(net.shape) == g["net"].shape and net.dtype == torch.float32
    assert rel(net.cpu().numpy(), g["net"]) < 4e-3
    assert end_points["resnet_v1_50/block4"] is net


def test_extract_features_then_prediction_layers_equals_get_net(pn):
    net_, g, frames, nj = pn
    net, ep = net_.extract_features(frames)
    heads = net_.prediction_layers(net, ep)
    fused = net_.get_net(frames)
    assert torch.equal(heads["part_pred"], fused["part_pred"]) and torch.equal(heads["locref"], fused["locref"])


def test_prediction_layer_and_dgp_prediction_layer_vs_reference_golden(pn):
    """The module-level prediction_layer (pose_net.py:18-26) and dgp_prediction_layer (fitdgp_util.py:18-74), fed the
    reference's own `net`, against the reference's own outputs; graph variables and init_flag constants."""
    from deepgraphpose_b200 import fitdgp_util, pose_net
    net_, g, frames, nj = pn
    cfg = net_.cfg
    ref_net = torch.from_numpy(g["net"]).cuda()
    part = pose_net.prediction_layer(cfg, ref_net, "part_pred", nj, engine=net_)
    loc = pose_net.prediction_layer(cfg, ref_net, "locref_pred", 2 * nj, engine=net_)
    assert rel(part.cpu().numpy(), g["part_pred"]) < 2e-3 and rel(loc.cpu().numpy(), g["locref_pred"]) < 2e-3
    # through the Features tensor of extract_features the engine is implied, as TF variable scopes imply the variables
    net, _ = net_.extract_features(frames[:1])
    part2 = pose_net.prediction_layer(cfg, net, "part_pred", nj)
    assert rel(part2.cpu().numpy(), g["part_pred"]) < 4e-3
    dgp = fitdgp_util.dgp_prediction_layer(None, None, cfg, ref_net, name="part_pred", num_outputs=nj, init_flag=False,
                                           nc=None, train_flag=True, stride=2, engine=net_.engine)
    assert torch.equal(dgp, part)
    const = fitdgp_util.dgp_prediction_layer(g["w_const"], g["b_const"], cfg, ref_net, "confidencemap", nj, True, 2048, True)
    assert rel(const.cpu().numpy(), g["part_pred_const"]) < 2e-3
    with pytest.raises(ValueError):
        pose_net.prediction_layer(cfg, ref_net, "part_pred", nj + 1, engine=net_)
    # Dataset._compute_pred_dims' throw-away layer: only its output shape matters
    fresh = fitdgp_util.dgp_prediction_layer(None, None, cfg, ref_net, "confidencemap", nj, 0, 3, 1)
    assert tuple(fresh.shape) == tuple(g["part_pred"].shape)


def test_posenet_inference_batched_branch_vs_reference_golden(pn):
    """batch_size > 1 (pose_net.py:129-163): peaks / likelihoods per (frame, joint); with reference_batched_locref=True the
    offsets follow the reference's un-transposed reshape literally (golden pose_tf_batched), by default each frame's own."""
    net_, g, frames, nj = pn
    ref = g["pose_tf_batched"]
    quirk = net_.inference(frames, reference_batched_locref=True)["pose"].cpu().numpy()
    assert quirk.shape == ref.shape == (2 * nj, 3)
    assert np.abs(quirk[:, :2] - ref[:, :2]).max() < 0.5 and np.abs(quirk[:, 2] - ref[:, 2]).max() < 1e-2
    fixed = net_.inference(frames)["pose"].cpu().numpy()
    single = np.concatenate([net_.inference(frames[i:i + 1])["pose"].cpu().numpy() for i in range(2)])
    assert np.array_equal(fixed, single)
    assert np.abs(fixed[:, 2] - ref[:, 2]).max() < 1e-2        # likelihoods are unaffected by the quirk


def test_argmax_2d_from_cm_threshold_vs_reference_golden():
    from deepgraphpose_b200 import fitdgp_util
    g = load("boundary.npz")
    nj = int(g["meta"][0])
    mu, sm = fitdgp_util.argmax_2d_from_cm(torch.from_numpy(g["th_x"]).cuda(), nj, 1.0, 1, th=float(g["th"]))
    assert np.abs(mu.cpu().numpy() - g["th_mu"]).max() < 1e-3
    assert np.abs(sm.cpu().numpy() - g["th_sm"]).max() < 1e-6 + 1e-4 * g["th_sm"].max()
    assert (sm.cpu().numpy() == 0).sum() == (g["th_sm"] == 0).sum() > 0


def test_deconv2d_matches_oracle_transposed_conv():
    """dgp_deconv2d for head widths that are not the network's own (odd Cout, Cin = 256): out[2i+k] += x[i] w[k] + bias."""
    from deepgraphpose_b200.engine import Engine
    from oracle import tf_ops
    rng = np.random.default_rng(2)
    eng = Engine(4)
    x = torch.from_numpy(rng.standard_normal((2, 5, 7, 256)).astype(np.float32))
    w = (rng.standard_normal((3, 3, 7, 256)) * 0.05).astype(np.float32)
    b = rng.standard_normal(7).astype(np.float32)
    ref = tf_ops.conv2d_transpose_same_s2(x.half().float(), torch.from_numpy(w).half().float(), torch.from_numpy(b))
    got = eng.deconv2d(x.cuda(), w, b)
    assert tuple(got.shape) == (2, 10, 14, 7)
    assert (got.cpu() - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()
    eng.close()
