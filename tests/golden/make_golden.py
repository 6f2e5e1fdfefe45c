"""Generate tests/golden/*.npz by executing the REFERENCE'S OWN code (imported unmodified from /root/reference) on top
of oracle/tf1_shim.py.  Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Inputs are either stored in the fixture (small tensors) or regenerated from seeds by
deepgraphpose_b200.synthetic (weights, frames), so the fixtures stay a few hundred KB.
"""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
OUT = os.path.dirname(os.path.abspath(__file__))

from oracle import tf1_shim  # noqa: E402

tf1_shim.install()
import tensorflow as tf  # noqa: E402  (the shim)

TF = tf.compat.v1
from deeplabcut.pose_estimation_tensorflow.nnet import pose_net as ref_pose_net  # noqa: E402
from deeplabcut.pose_estimation_tensorflow.nnet import predict as ref_predict  # noqa: E402
from deepgraphpose.models import eval as ref_eval  # noqa: E402
from deepgraphpose.models import fitdgp as ref_fitdgp  # noqa: E402
from deepgraphpose.models import fitdgp_util as ref_util  # noqa: E402

from deepgraphpose_b200 import synthetic  # noqa: E402


class Cfg(dict):
    """EasyDict-like config (the reference reads attributes and calls .keys())."""
    __getattr__ = dict.__getitem__
    __setattr__ = dict.__setitem__


def base_cfg(nj, **kw):
    c = Cfg(net_type="resnet_50", num_joints=nj, location_refinement=True, intermediate_supervision=False,
            mean_pixel=[123.68, 116.779, 103.939], stride=8.0, locref_stdev=7.2801, weight_decay=0.0001, batch_size=1,
            deconvolutionstride=2, output_stride=16, all_joints_names=["j%d" % i for i in range(nj)],
            locref_huber_loss=True, locref_loss_weight=0.05)
    c.update(kw)
    return c


def golden_softargmax():
    rng = np.random.default_rng(11)
    out = {}
    for tag, (N, H, W, nj, gamma, gl, scale) in {"a": (2, 12, 16, 3, 1.0, 1, 3.0), "b": (1, 30, 38, 5, 1.0, 1, 5.0),
                                                  "c": (2, 10, 14, 4, 2.0, 2, 2.0)}.items():
        x = (rng.standard_normal((N, H, W, nj)) * scale).astype(np.float32)
        x[0, 0, 0, 0] += 6 * scale           # a peak in the corner (zero-padded blur bias)
        x[-1, H - 1, W // 2, nj - 1] += 6 * scale
        ph = TF.placeholder(TF.float32, shape=[None, None, None, nj])
        mu, sm = ref_util.argmax_2d_from_cm(ph, nj, gamma, gl)
        mu_v, sm_v = TF.Session().run([mu, sm], {ph: x})
        out.update({tag + "_x": x, tag + "_mu": mu_v, tag + "_sm": sm_v, tag + "_par": np.array([nj, gamma, gl], np.float32)})
    np.savez_compressed(os.path.join(OUT, "softargmax.npz"), **out)


def golden_posenet():
    """PoseNet.test / inference (pose_net.py:84-163) + argmax_pose_predict (predict.py:62-77), synthetic weights seed 3."""
    nj = 3
    tf1_shim.set_variables(synthetic.make_weights(nj, seed=3))
    frames, _ = synthetic.make_video(2, 64, 96, nj, seed=7)
    cfg = base_cfg(nj)
    inputs = TF.placeholder(tf.float32, shape=[1, None, None, 3])
    pn = ref_pose_net.PoseNet(cfg)
    heads = pn.test(inputs)
    pose_t = ref_pose_net.PoseNet(cfg).inference(inputs)
    sess = TF.Session()
    out = {}
    for i in range(2):
        prob, locref = sess.run([heads["part_prob"], heads["locref"]], {inputs: frames[i][None]})
        pose_tf = sess.run(pose_t["pose"], {inputs: frames[i][None]})
        scmap, loc = ref_predict.extract_cnn_output([prob, locref.copy()], cfg)
        pose_np = ref_predict.argmax_pose_predict(scmap, loc, cfg.stride)
        out.update({"prob%d" % i: prob, "locref%d" % i: locref, "pose_tf%d" % i: pose_tf, "pose_np%d" % i: pose_np})
    out["meta"] = np.array([nj, 3, 7, 64, 96])  # nj, weight seed, video seed, H, W
    np.savez_compressed(os.path.join(OUT, "posenet.npz"), **out)


def golden_estimate_pose():
    """The literal estimate_pose loop (eval.py:217-372) on a fake 5-frame clip."""
    nj = 4
    W = synthetic.make_weights(nj, seed=5, location_refinement=False)
    tf1_shim.set_variables(W)
    frames, _ = synthetic.make_video(5, 64, 96, nj, seed=9)
    cfg = base_cfg(nj, location_refinement=False)

    class FakeClip:
        fps, duration, size = 5.0, 1.0, (96, 64)
        def __init__(self, *a): pass
        def iter_frames(self): return iter(frames)
        def close(self): pass

    ref_eval.VideoFileClip = FakeClip
    ref_eval.img_as_ubyte = lambda a: a
    import importlib
    importlib.import_module("deepgraphpose.utils_model").get_train_config = lambda proj, shuffle=1: cfg
    with tempfile.TemporaryDirectory() as d:
        yml = os.path.join(d, "config.yaml")
        open(yml, "w").write("bodyparts: [a, b, c, d]\n")
        labels = ref_eval.estimate_pose(yml, "synthetic.ckpt", os.path.join(d, "clip.avi"), d, save_pose=False)
    np.savez_compressed(os.path.join(OUT, "estimate_pose.npz"), x=labels["x"], y=labels["y"],
                        likelihoods=labels["likelihoods"], meta=np.array([nj, 5, 9, 64, 96, 5]))


def golden_dgp_loss():
    """dgp_loss (fitdgp.py:848-1144) on a 4-frame batch: 2 visible frames (one NaN label), 2 hidden; wt = 0 and wt > 0."""
    nj = 3
    Wts = synthetic.make_weights(nj, seed=4)
    tf1_shim.set_variables(Wts)
    nt, H, Wd = 4, 64, 96
    frames, _ = synthetic.make_video(nt, H, Wd, nj, seed=13)
    nx_out, ny_out = 8, 12
    rng = np.random.default_rng(21)
    labels = np.stack([rng.uniform(1, nx_out - 2, (2, nj)), rng.uniform(1, ny_out - 2, (2, nj))], axis=2)
    labels[1, 2, :] = np.nan
    S0 = np.zeros((2, nj))
    S0[0, 0], S0[0, 1], S0[1, 1], S0[1, 2] = 1, -1, 1, -1
    out = {"labels": labels, "S0": S0, "meta": np.array([nj, 4, 13, H, Wd, nt])}
    # batch bookkeeping exactly as gen_idx_chunk (dataset.py:187-239): frames 0,2 visible; 1,3 hidden
    visible_frames, hidden_frames = np.array([0, 2]), np.array([1, 3])
    from itertools import chain  # noqa: F401
    nan_ind = sorted(int(nj * visible_frames[i] + j) for j in range(nj) for i in range(2) if np.isnan(labels[i, j, 0]))
    hidden_marker = np.sort(list(np.sort(np.array([hidden_frames * nj + i for i in range(nj)]).flatten())) + nan_ind)
    vm0 = np.sort(np.array([visible_frames * nj + i for i in range(nj)]).flatten())
    visible_marker = np.sort(np.setdiff1d(vm0, nan_ind))
    vis_in_targets = np.nonzero(np.in1d(vm0, visible_marker))[0]
    locref_map = rng.normal(0, 0.5, (nt, nx_out, ny_out, 2 * nj))
    locref_mask = (rng.uniform(size=(nt, nx_out, ny_out, 2 * nj)) < 0.15).astype(np.float64)
    locref_map[[1, 3]] = 0
    locref_mask[[1, 3]] = 0
    xg, yg = np.meshgrid(np.linspace(0, nx_out - 1, nx_out), np.linspace(0, ny_out - 1, ny_out))
    alpha = np.array([xg, yg]).swapaxes(1, 2)
    vector_field = np.abs(rng.normal(0, 1.0, (nt - 1, H, Wd)))
    out.update(visible_marker=visible_marker, hidden_marker=hidden_marker, vis_in_targets=vis_in_targets,
               locref_map=locref_map, locref_mask=locref_mask, vector_field=vector_field,
               wt_batch_mask=np.array([1.0, 1.0, 0.0]))

    class DS:
        pass
    ds = DS()
    ds.labels = labels
    db = DS()
    db.S0, db.nj, db.n_frames_total, db.n_visible_frames_total, db.datasets = S0, nj, 120, 10, [ds]
    for tag, wt in (("wt0", 0.0), ("wt1", 1.0)):
        cfg = base_cfg(nj, ws=1000.0, ws_max=1.2, wt=wt, wt_max=0.0, wn_visible=5.0, wn_hidden=3.0, gamma=1, gm2=1, gm3=3,
                       lengthscale=1, gauss_len=1, lr=0.005)
        loss, total_loss, total_loss_visible, ph = ref_fitdgp.dgp_loss(db, cfg)
        feed = {ph["inputs"]: frames.astype(np.float64), ph["targets"]: labels, ph["locref_map"]: locref_map,
                ph["locref_mask"]: locref_mask, ph["visible_marker_pl"]: visible_marker,
                ph["hidden_marker_pl"]: hidden_marker, ph["visible_marker_in_targets_pl"]: vis_in_targets,
                ph["wt_batch_mask_pl"]: out["wt_batch_mask"], ph["vector_field_tf"]: vector_field,
                ph["nt_batch_pl"]: nt, ph["wt_batch_pl"]: np.ones(nt - 1) * wt, ph["alpha_tf"]: alpha}
        vals, tl, tlv = TF.Session().run([loss, total_loss, total_loss_visible], feed)
        for k, v in vals.items():
            out["%s_%s" % (tag, k)] = np.asarray(v)
        out[tag + "_total_loss_visible"] = np.asarray(tlv)
        print(tag, {k: float(v) for k, v in vals.items()})
    np.savez_compressed(os.path.join(OUT, "dgp_loss.npz"), **out)


def golden_boundary():
    """The reference-named pieces of the boundary: PoseNet.extract_features, prediction_layer,
    dgp_prediction_layer (graph variables and init_flag constants), the batched branch of PoseNet.inference
    (pose_net.py:129-163) and argmax_2d_from_cm(th=...).  Synthetic weights seed 3, video seed 7, 64x96 frames."""
    nj = 3
    W = synthetic.make_weights(nj, seed=3)
    tf1_shim.set_variables(W)
    frames, _ = synthetic.make_video(2, 64, 96, nj, seed=7)
    cfg = base_cfg(nj)
    out = {"meta": np.array([nj, 3, 7, 64, 96])}
    inputs = TF.placeholder(tf.float32, shape=[None, None, None, 3])
    pn = ref_pose_net.PoseNet(cfg)
    net, end_points = pn.extract_features(inputs)
    with tf.variable_scope("pose", reuse=None):
        part = ref_pose_net.prediction_layer(cfg, net, "part_pred", nj)
        loc = ref_pose_net.prediction_layer(cfg, net, "locref_pred", 2 * nj)
        part_dgp = ref_util.dgp_prediction_layer(None, None, cfg, net, name="part_pred", num_outputs=nj, init_flag=False,
                                                 nc=None, train_flag=True, stride=cfg.deconvolutionstride)
    rng = np.random.default_rng(5)
    w_const = (rng.standard_normal((3, 3, nj, 2048)) * 0.01).astype(np.float32)
    b_const = rng.standard_normal((1, nj)).astype(np.float32)
    part_const = ref_util.dgp_prediction_layer(w_const, b_const, cfg, net, "confidencemap", nj, True, 2048, True)
    sess = TF.Session()
    net_v, part_v, loc_v, part_dgp_v, part_const_v = sess.run([net, part, loc, part_dgp, part_const], {inputs: frames[:1]})
    out.update(net=net_v, part_pred=part_v, locref_pred=loc_v, part_pred_dgp=part_dgp_v, w_const=w_const, b_const=b_const,
               part_pred_const=part_const_v)
    # batched PoseNet.inference: batch_size = 2 takes the else-branch (pose_net.py:129-163)
    cfg2 = base_cfg(nj, batch_size=2)
    inputs2 = TF.placeholder(tf.float32, shape=[2, None, None, 3])
    pose_b = ref_pose_net.PoseNet(cfg2).inference(inputs2)
    out["pose_tf_batched"] = TF.Session().run(pose_b["pose"], {inputs2: frames})
    # argmax_2d_from_cm with the threshold branch (fitdgp_util.py:379-389)
    x = (rng.standard_normal((2, 12, 16, nj)) * 3.0).astype(np.float32)
    ph = TF.placeholder(TF.float32, shape=[None, None, None, nj])
    mu, sm = ref_util.argmax_2d_from_cm(ph, nj, 1.0, 1, th=0.3)
    mu_v, sm_v = TF.Session().run([mu, sm], {ph: x})
    out.update(th_x=x, th_mu=mu_v, th_sm=sm_v, th=np.float32(0.3))
    np.savez_compressed(os.path.join(OUT, "boundary.npz"), **out)


if __name__ == "__main__":
    if "--boundary-only" in sys.argv:
        golden_boundary()
        sys.exit(0)
    golden_softargmax()
    golden_posenet()
    golden_estimate_pose()
    golden_dgp_loss()
    golden_boundary()
    for f in sorted(os.listdir(OUT)):
        if f.endswith(".npz"):
            print(f, os.path.getsize(os.path.join(OUT, f)))
