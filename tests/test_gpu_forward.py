"""Whole hot path on the GPU vs the oracle: layer-wise activations, logits, scoremaps, coordinates, shims."""
import numpy as np
import pytest
import torch

from deepgraphpose_b200 import synthetic
from oracle import dgp_ops, pose_net

pytestmark = pytest.mark.gpu

# BASELINE.json north_star tolerances, asserted as such for the storage mode the shims and bench.py run (fp16 operands on
# tcgen05 kind::f16, fp32 accumulate): sigmoid scoremaps <= 1e-2 max-abs, soft-argmax coordinates <= 0.5 px.
# argmax_2d_from_cm returns SCOREMAP pixels (fitdgp_util.py:342-402; estimate_pose scales by the stride afterwards,
# eval.py:331-357), so 0.5 px of the soft-argmax output is the north_star bound (MU_TOL_PX).  The 8x stricter reading
# (0.5 IMAGE px = 0.0625 scoremap px, MU_TOL_SCOREMAP_PX) is asserted wherever it holds: the layer-wise shapes, the
# golden estimate_pose video and the boundary tests.  On the full-size random-init maps of configs a/b/c/e the soft-argmax
# at gamma = 1 is ill-conditioned (several far-apart pixels carry equal softmax weight: a logit error d on a pixel of weight
# p, 40 rows from the mean, moves the mean by 40*p*d): measured 0.05 - 0.54 image px over seeds and shapes
# (tools/diag_precision.py, profiles/r02_precision.md), i.e. around that stricter line, so it is reported there and not
# asserted.  Measured worst cases: logits 1.8e-3 of their max, sigmoid 5.4e-3 (flat set) / 9.2e-3 (trained-like set).
ACT_REL_TOL = 4e-3
LOGIT_REL_TOL = 4e-3
SIGMOID_TOL = 1e-2
MU_TOL_PX = 0.5
MU_TOL_SCOREMAP_PX = 0.0625
# The bf16 storage mode (precision="bf16", not what bench.py reports) carries 8 mantissa bits through 53 conv layers and
# cannot meet those tolerances on this net (bf16 weights alone give 1.9e-2, DESIGN.md "Numerics"); its own, wider bounds:
BF16_ACT_REL_TOL = 2.5e-2
BF16_SIGMOID_TOL = 5e-2


@pytest.fixture(scope="module")
def setup():
    from deepgraphpose_b200.engine import Engine
    nj = 4
    W = synthetic.make_weights(nj, seed=0)
    eng = Engine(nj, location_refinement=True)
    eng.load_weights(W)
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    yield eng, W, Wt, nj
    eng.close()


@pytest.mark.parametrize("shape", [(2, 235, 301), (1, 470, 640), (3, 64, 96), (2, 101, 75), (1, 747, 832)])
def test_forward_layerwise(setup, shape):
    eng, W, Wt, nj = setup
    T, H, Wd = shape
    frames, _ = synthetic.make_video(T, H, Wd, nj, seed=H)
    ep = {}
    with torch.no_grad():
        net = pose_net.extract_features(torch.from_numpy(frames.astype(np.float32)), Wt, ep)
        pred = pose_net.prediction_layer(net, Wt, "part_pred")
        loc = pose_net.prediction_layer(net, Wt, "locref_pred")
    eng.keep_activations(True)
    logits, locref = eng.forward(torch.from_numpy(frames).cuda())
    torch.cuda.synchronize()
    eng.keep_activations(False)
    # the debug plan runs conv1 and pool1 as two launches; the regular plan fuses the pool into conv1's epilogue (2-D patch
    # tiles, conv1's output never stored): same 16-bit values, so everything downstream is bit-identical
    logits_fused, locref_fused = eng.forward(torch.from_numpy(frames).cuda())
    assert torch.equal(logits_fused, logits) and torch.equal(locref_fused, locref)
    for name, ref in ep.items():
        got = torch.from_numpy(eng.get_activation(name))
        assert got.shape == ref.shape, name
        rel = (got - ref).abs().max().item() / ref.abs().max().item()
        assert rel < ACT_REL_TOL, (name, rel)
    assert logits.shape == pred.shape and locref.shape == loc.shape
    assert (logits.cpu() - pred).abs().max().item() / pred.abs().max().item() < LOGIT_REL_TOL
    assert (locref.cpu() - loc).abs().max().item() / loc.abs().max().item() < LOGIT_REL_TOL
    assert (torch.sigmoid(logits.cpu()) - torch.sigmoid(pred)).abs().max().item() < SIGMOID_TOL
    mu_ref, _ = dgp_ops.argmax_2d_from_cm(pred, nj, 1.0, 1.0)
    out = eng.softargmax(logits, locref)
    assert (out["mu"].cpu() - mu_ref).abs().max().item() < MU_TOL_SCOREMAP_PX


def test_forward_bf16_storage_mode(setup):
    """precision='bf16' (same kernels, bf16 instead of fp16 storage) stays within ITS documented bounds: an optional mode for
    weights whose activations would overflow fp16; it does not meet BASELINE's 1e-2 and is not the benchmarked mode."""
    from deepgraphpose_b200.engine import Engine
    _, W, Wt, nj = setup
    eng = Engine(nj, location_refinement=True, precision="bf16")
    eng.load_weights(W)
    for (T, H, Wd) in [(2, 235, 301), (1, 470, 640)]:
        frames, _ = synthetic.make_video(T, H, Wd, nj, seed=H)
        with torch.no_grad():
            net = pose_net.extract_features(torch.from_numpy(frames.astype(np.float32)), Wt)
            pred = pose_net.prediction_layer(net, Wt, "part_pred")
            loc = pose_net.prediction_layer(net, Wt, "locref_pred")
        logits, locref = eng.forward(torch.from_numpy(frames).cuda())
        assert (logits.cpu() - pred).abs().max().item() / pred.abs().max().item() < BF16_ACT_REL_TOL
        assert (locref.cpu() - loc).abs().max().item() / loc.abs().max().item() < BF16_ACT_REL_TOL
        assert (torch.sigmoid(logits.cpu()) - torch.sigmoid(pred)).abs().max().item() < BF16_SIGMOID_TOL
    eng.close()


BASELINE_SHAPES = [dict(nj=5, H=747, W=832, locref=True, skel="demo"),      # configs[0]: the bundled demo project
                   dict(nj=4, H=747, W=832, locref=False, skel="chain"),    # configs[1]: the benchmarked workload
                   dict(nj=16, H=1024, W=1280, locref=False, skel="chain"),  # configs[2]
                   dict(nj=20, H=480, W=640, locref=True, skel="dense")]    # configs[4]


@pytest.mark.parametrize("cfg", BASELINE_SHAPES, ids=["a-demo", "b-reaching", "c-1280x1024", "e-640x480"])
def test_baseline_config_shapes_meet_baseline_tolerances(cfg):
    """One frame of every BASELINE.json inference configuration through the whole path vs the fp32 oracle, at the north_star
    tolerances: sigmoid scoremaps <= 1e-2 max-abs, soft-argmax coordinates <= 0.5 px (units of argmax_2d_from_cm's output,
    see the header); integer peaks bit-exact and coordinates <= 1e-3 px when computed from the same fp32 maps; potentials
    vs the oracle."""
    from deepgraphpose_b200.engine import Engine
    nj = cfg["nj"]
    W = synthetic.make_weights(nj, seed=2, location_refinement=cfg["locref"])
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    eng = Engine(nj, location_refinement=cfg["locref"])
    assert eng.precision == "fp16"
    eng.load_weights(W)
    frames, _ = synthetic.make_video(1, cfg["H"], cfg["W"], nj, seed=11)
    with torch.no_grad():
        net = pose_net.extract_features(torch.from_numpy(frames.astype(np.float32)), Wt)
        pred = pose_net.prediction_layer(net, Wt, "part_pred")
        loc = pose_net.prediction_layer(net, Wt, "locref_pred") if cfg["locref"] else None
    logits, locref = eng.forward(torch.from_numpy(frames).cuda())
    assert logits.shape == pred.shape
    assert (logits.cpu() - pred).abs().max().item() / pred.abs().max().item() < LOGIT_REL_TOL
    assert (torch.sigmoid(logits.cpu()) - torch.sigmoid(pred)).abs().max().item() < SIGMOID_TOL
    if cfg["locref"]:
        assert (locref.cpu() - loc).abs().max().item() / loc.abs().max().item() < LOGIT_REL_TOL
    out = eng.softargmax(logits, locref)
    mu_oracle, _ = dgp_ops.argmax_2d_from_cm(pred, nj, 1.0, 1.0)             # oracle network -> oracle soft-argmax
    mu_err = (out["mu"].cpu() - mu_oracle).abs().max().item()
    print("soft-argmax vs oracle net: %.4f scoremap px = %.3f image px" % (mu_err, 8.0 * mu_err))
    assert mu_err < MU_TOL_PX
    mu_ref, _ = dgp_ops.argmax_2d_from_cm(logits.cpu(), nj, 1.0, 1.0)        # same fp32 maps -> tight
    assert (out["mu"].cpu() - mu_ref).abs().max().item() < 1e-3
    _, pk, _ = dgp_ops.estimate_pose_readout(out["mu"].cpu().numpy(), logits.cpu().numpy())
    assert (pk == out["peak"][0].cpu().numpy()).all()
    edges = {"chain": synthetic.chain_skeleton(nj), "dense": synthetic.dense_skeleton(nj), "demo": [(0, 1), (3, 4)]}[cfg["skel"]]
    pot = eng.potentials(out["mu"], edges)
    d_ref = dgp_ops.skeleton_distances(out["mu"].cpu(), dgp_ops.skeleton_matrix(edges, nj))
    assert (pot["skel"].cpu() - d_ref).abs().max().item() < 1e-3
    eng.close()


def test_trained_like_weights_meet_the_scoremap_tolerance():
    """The second synthetic weight set (synthetic.make_weights(trained_like=True): near-zero / zero BatchNorm gammas as in an
    ImageNet checkpoint, part_pred gain x2 with bias -4 -> sparse, confident scoremaps): sigmoid <= 1e-2 at the benchmarked
    shape.  Its random multi-peaked maps make the soft-argmax ill-conditioned (two far-apart pixels of equal weight), so
    the coordinate check is made where the reference's tolerance is meaningful: from the same fp32 maps."""
    from deepgraphpose_b200.engine import Engine
    nj = 4
    W = synthetic.make_weights(nj, seed=0, location_refinement=False, trained_like=True)
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    eng = Engine(nj, location_refinement=False)
    eng.load_weights(W)
    frames, _ = synthetic.make_video(1, 747, 832, nj, seed=7)
    with torch.no_grad():
        pred = pose_net.prediction_layer(pose_net.extract_features(torch.from_numpy(frames.astype(np.float32)), Wt), Wt, "part_pred")
    logits, _ = eng.forward(torch.from_numpy(frames).cuda())
    prob, ref = torch.sigmoid(logits.cpu()), torch.sigmoid(pred)
    assert (prob - ref).abs().max().item() < SIGMOID_TOL
    assert float((ref > 0.5).float().mean()) < 0.35          # sparse map, not the flat 0.5 field of the random-init set
    out = eng.softargmax(logits)
    mu_ref, _ = dgp_ops.argmax_2d_from_cm(logits.cpu(), nj, 1.0, 1.0)
    assert (out["mu"].cpu() - mu_ref).abs().max().item() < 1e-3
    _, pk, _ = dgp_ops.estimate_pose_readout(out["mu"].cpu().numpy(), logits.cpu().numpy())
    assert (pk == out["peak"][0].cpu().numpy()).all()
    eng.close()


def test_real_demo_frame():
    """BASELINE configs[0]: a labelled frame of the bundled Reaching-Mackenzie demo project (5 bodyparts, skeleton
    Hand-Finger1 / Joystick1-Joystick2, data/.../config.yaml:6-31) through the whole path vs the oracle."""
    import os
    from PIL import Image
    from deepgraphpose_b200.engine import Engine
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "demo_frame_img119.png")
    frame = np.ascontiguousarray(np.asarray(Image.open(path).convert("RGB")))[None]
    nj = 5
    W = synthetic.make_weights(nj, seed=0)
    Wt = {k: torch.from_numpy(v) for k, v in W.items()}
    with torch.no_grad():
        net = pose_net.extract_features(torch.from_numpy(frame.astype(np.float32)), Wt)
        pred = pose_net.prediction_layer(net, Wt, "part_pred")
    eng = Engine(nj)
    eng.load_weights(W)
    logits, locref = eng.forward(torch.from_numpy(frame).cuda())
    assert logits.shape == pred.shape
    assert (logits.cpu() - pred).abs().max().item() / pred.abs().max().item() < LOGIT_REL_TOL
    out = eng.softargmax(logits, locref)
    mu_ref, _ = dgp_ops.argmax_2d_from_cm(logits.cpu(), nj, 1.0, 1.0)
    assert (out["mu"].cpu() - mu_ref).abs().max().item() < 1e-3
    _, pk, lk = dgp_ops.estimate_pose_readout(out["mu"].cpu().numpy(), logits.cpu().numpy())
    assert (pk == out["peak"][0].cpu().numpy()).all()
    pot = eng.potentials(out["mu"], [(0, 1), (3, 4)])
    d_ref = dgp_ops.skeleton_distances(out["mu"].cpu(), dgp_ops.skeleton_matrix([(0, 1), (3, 4)], nj))
    assert (pot["skel"].cpu() - d_ref).abs().max().item() < 1e-3
    eng.close()


def test_batch_invariance(setup):
    """A frame's outputs do not depend on the batch it is processed in (needed for bit-exact frame sharding)."""
    eng, W, Wt, nj = setup
    frames, _ = synthetic.make_video(5, 96, 128, nj, seed=3)
    f = torch.from_numpy(frames).cuda()
    la, _ = eng.forward(f)
    lb = torch.cat([eng.forward(f[:2])[0], eng.forward(f[2:])[0]])
    assert torch.equal(la, lb)
    oa = eng.softargmax(la)
    ob = eng.softargmax(lb[3:4])
    assert torch.equal(oa["mu"][3:4], ob["mu"]) and torch.equal(oa["peak"][3:4], ob["peak"])


def test_session_shim_and_estimate_pose(setup):
    """setup_dgp_eval_graph / sess.run / estimate_pose keep the reference call surface (eval.py:147-372)."""
    from deepgraphpose_b200 import eval as dgp_eval
    eng, W, Wt, nj = setup
    frames, _ = synthetic.make_video(5, 96, 128, nj, seed=4)
    cfg = {"num_joints": nj, "net_type": "resnet_50", "stride": 8.0}
    Wp = {k: v for k, v in W.items() if "locref" not in k}
    sess, mu_n, softmax_tensor, scmap, locref, inputs = dgp_eval.setup_dgp_eval_graph(cfg, Wp)
    assert locref is None
    mu_b, sc_b = sess.run([mu_n, scmap], feed_dict={inputs: frames[0][None, :, :, :]})
    assert mu_b.shape == (1, nj, 2) and sc_b.shape == (1, 12, 16, nj) and mu_b.dtype == np.float32
    # the reference's own post-processing applied to these fetches == our fused read-out
    markers, peaks, lik = dgp_ops.estimate_pose_readout(mu_b, sc_b)
    res = dgp_eval.estimate_pose_frames(sess.engine, frames, batch=2)
    assert res["x"].shape == (5, nj)
    assert np.array_equal(res["mu_likelihoods"][0], peaks)
    assert np.allclose(res["likelihoods"][0], lik, atol=1e-6)
    xr, yr = dgp_ops.estimate_pose_xy(markers[None])
    assert np.allclose(res["x"][0], xr[0], atol=1e-4) and np.allclose(res["y"][0], yr[0], atol=1e-4)
    sm = sess.run(softmax_tensor, feed_dict={inputs: frames[:2].astype(np.float32)})
    assert sm.shape == (2, 12, 16, nj) and abs(sm[0, :, :, 0].sum() - 1) < 1e-4
    sess.close()
    with pytest.raises(RuntimeError):
        sess.run(mu_n, feed_dict={inputs: frames[:1]})
    labels = dgp_eval.estimate_pose(cfg, Wp, frames, "/tmp", save_pose=False, batch=4)
    assert np.allclose(labels["x"], res["x"]) and set(labels) == {"x", "y", "likelihoods"}


def test_posenet_shim(setup):
    from deepgraphpose_b200.pose_net import PoseNet
    eng, W, Wt, nj = setup
    frames, _ = synthetic.make_video(2, 64, 96, nj, seed=5)
    pn = PoseNet({"num_joints": nj, "location_refinement": True, "net_type": "resnet_50"}, variables=W)
    out = pn.test(frames)
    with torch.no_grad():
        ref = pose_net.test(torch.from_numpy(frames.astype(np.float32)), Wt)
    assert (out["part_prob"].cpu() - ref["part_prob"]).abs().max().item() < SIGMOID_TOL
    pose = pn.inference(frames)["pose"].cpu().numpy()
    prob = out["part_prob"].cpu().numpy()
    loc = out["locref"].cpu().numpy()
    for b in range(2):
        scm, off = pose_net.extract_cnn_output(prob[b:b + 1], loc[b:b + 1])
        ref_pose, _ = pose_net.argmax_pose_predict(scm, off, 8.0)           # (x, y, lik); PoseNet.inference emits (y, x, lik)
        assert np.abs(ref_pose - pose[b * nj:(b + 1) * nj][:, [1, 0, 2]]).max() < 1e-3


def test_forward_errors(setup):
    from deepgraphpose_b200._lib import DgpError
    from deepgraphpose_b200.engine import Engine
    eng, W, Wt, nj = setup
    with pytest.raises(ValueError):
        eng.forward(torch.zeros(1, 64, 64, 3, device="cuda"))           # float frames
    e2 = Engine(nj)
    with pytest.raises(DgpError):
        e2.forward(torch.zeros(1, 64, 64, 3, dtype=torch.uint8, device="cuda"))   # weights not loaded
    bad = dict(W)
    bad["resnet_v1_50/conv1/weights"] = np.zeros((3, 3, 3, 64), np.float32)
    with pytest.raises(DgpError):
        e2.load_weights(bad)
    e2.close()


def test_evaluate_dgp_frames_branches():
    """The three pose read-outs of evaluate_dgp (eval.py:744-790) through the shim: shapes, likelihood column, consistency of
    the no-locref branch with the soft-argmax and of the 'dlc' branch with argmax_pose_predict on the same maps."""
    from deepgraphpose_b200 import synthetic
    from deepgraphpose_b200.engine import Engine
    from deepgraphpose_b200.eval import evaluate_dgp_frames
    from oracle import dgp_ops, pose_net
    nj = 4
    eng = Engine(nj)
    eng.load_weights(synthetic.make_weights(nj, seed=2))
    frames, _ = synthetic.make_video(3, 96, 128, nj, seed=4)
    dlc = evaluate_dgp_frames(eng, frames, True, "dlc", batch=2)
    dgp = evaluate_dgp_frames(eng, frames, True, "dgp", batch=2)
    nol = evaluate_dgp_frames(eng, frames, False, "dlc", batch=2)
    assert dlc.shape == dgp.shape == nol.shape == (3, nj * 3)
    assert np.isfinite(dlc).all() and np.isfinite(dgp).all() and (dgp.reshape(3, nj, 3)[:, :, 2] == 1).all()
    logits, locref = eng.forward(torch.from_numpy(frames).cuda())
    lg, lr = logits.cpu().numpy(), locref.cpu().numpy()
    for t in range(3):
        mu, st = dgp_ops.argmax_2d_from_cm(torch.from_numpy(lg[t:t + 1]), nj, 1.0, 1.0)
        assert np.abs(nol[t].reshape(nj, 3) - dgp_ops.evaluate_dgp_pose_noloc(mu.numpy())).max() < 2e-3
        assert np.abs(dgp[t].reshape(nj, 3) - dgp_ops.evaluate_dgp_pose_dgp_branch(st.numpy(), lr[t:t + 1])).max() < 2e-3
        sc, off = pose_net.extract_cnn_output(1.0 / (1.0 + np.exp(-lg[t:t + 1])), lr[t:t + 1])
        ref, _ = pose_net.argmax_pose_predict(sc, off)
        assert np.abs(dlc[t].reshape(nj, 3) - ref).max() < 2e-3
    eng.close()


def test_estimate_pose_from_video_file_writes_dlc_csv(tmp_path):
    """estimate_pose on a video FILE (OpenCV decode) with save_pose=True: DLC-format csv next to the expected name,
    read back by load_pose_from_dlc_to_dict; a second call skips the already-labelled video like the reference."""
    cv2 = pytest.importorskip("cv2")
    from deepgraphpose_b200 import eval as dgp_eval
    from deepgraphpose_b200 import synthetic
    nj, T, H, W = 4, 6, 96, 128
    frames, _ = synthetic.make_video(T, H, W, nj, seed=6)
    path = str(tmp_path / "reach.avi")
    vw = cv2.VideoWriter(path, cv2.VideoWriter_fourcc(*"MJPG"), 10.0, (W, H))
    if not vw.isOpened():
        pytest.skip("no MJPG encoder in this OpenCV build")
    for f in frames:
        vw.write(f[:, :, ::-1])
    vw.release()
    cfg = {"num_joints": nj, "all_joints_names": ["a", "b", "c", "d"], "stride": 8.0}
    labels = dgp_eval.estimate_pose(cfg, "synthetic:3", path, str(tmp_path), save_pose=True, batch=4)
    assert labels["x"].shape == (T, nj) and np.isfinite(labels["x"]).all()
    csv = str(tmp_path / "reach_labeled.csv")
    back = dgp_eval.load_pose_from_dlc_to_dict(csv)
    assert np.allclose(back["x"], labels["x"]) and np.allclose(back["likelihoods"], labels["likelihoods"])
    head = open(csv).read().splitlines()[:3]
    assert head[0].startswith("scorer,synthetic:3") and head[1].startswith("bodyparts,a,a,a,b") and head[2].startswith("coords,x,y,likelihood")
    assert dgp_eval.estimate_pose(cfg, "synthetic:3", path, str(tmp_path), save_pose=True) == csv
