"""Drop-in for the hot-path part of deepgraphpose.models.fitdgp (reference: src/deepgraphpose/models/fitdgp.py).

``dgp_loss(data_batcher, dgp_cfg)`` keeps the reference's signature and 4-tuple return
``(loss, total_loss, total_loss_visible, placeholders)`` (:848-1144); the handles are evaluated by ``TrainSession.run``
with the reference's feed_dict keys.  ``momentum_train_op`` stands in for fitdgp.py:706-713 (MomentumOptimizer + global-norm
clipping); fetching it from ``TrainSession.run`` runs the whole training step on the GPU (network forward, fused loss
kernels, dgrad/wgrad of every conv on tcgen05, optimizer) and, when ``torch.distributed`` is initialised, averages the
gradients over the ranks with one NCCL all-reduce of the flat gradient buffer (``dp.allreduce_gradients``).
"""
import ctypes as C
from types import SimpleNamespace

import numpy as np
import torch

from . import _lib
from .engine import Engine, _ptr, _stream
from .session import Handle

LOSS_KEYS = ("visible_loss_pred", "hidden_loss_pred", "visible_loss_locref", "ws_loss", "wt_loss", "total_loss")
PLACEHOLDER_KEYS = ("inputs", "targets", "locref_map", "locref_mask", "visible_marker_pl", "hidden_marker_pl",
                    "visible_marker_in_targets_pl", "wt_batch_mask_pl", "vector_field_tf", "nt_batch_pl", "wt_batch_pl",
                    "alpha_tf")
# One placeholder beyond the reference's twelve: feeding the batch positions of the visible frames (fit_dgp's own
# `visible_frame_within_batch`, fitdgp.py:763) INSTEAD of locref_map / locref_mask makes the device build the two maps
# (coord2map feeder kernel) rather than copying 2 x (nt,H,W,2nj) float64 arrays from the host every step.
EXTRA_PLACEHOLDER_KEYS = ("visible_frame_within_batch",)


def _get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def skeleton_edges(S0):
    """Rows of the reference's S0 incidence matrix (fitdgp.py:607-617) -> (+1 joint, -1 joint) pairs."""
    S0 = np.asarray(S0)
    edges = []
    for row in S0:
        a, b = np.where(row > 0)[0], np.where(row < 0)[0]
        if len(a) != 1 or len(b) != 1:
            raise ValueError("S0 rows must hold exactly one +1 and one -1")
        edges.append((int(a[0]), int(b[0])))
    return edges


def spatial_clique_params(joint_locs, S0, stride, ws, ws_max):
    """Host precompute of dgp_loss (fitdgp.py:874-892): per limb l = (a, b) of the skeleton and over every labelled frame t,
    ``L[t, l] = |label[t, a] - label[t, b]| * stride + stride / 2`` in image pixels, then the clique's upper bound
    ``ws_max_l = ws_max * max_t L`` and weight ``ws_l = ws / mean_t L``.  A coordinate difference with an unlabelled (NaN) end
    counts as 0, so a limb missing from a frame enters the mean with length stride/2 -- the reference's behaviour (its
    ``!= 0`` mask never excludes anything because stride/2 was already added), kept for parity.
    joint_locs: one (n_vis_i, nj, 2) array of scoremap (row, col) labels per dataset."""
    edges = skeleton_edges(S0)
    labelled = [np.asarray(j, dtype=np.float64) for j in joint_locs if len(j) > 0]
    if not labelled:
        raise ValueError("spatial_clique_params needs at least one labelled frame (the reference takes max over frames)")
    lab = np.concatenate(labelled, axis=0)
    head = np.array([e[0] for e in edges], dtype=np.int64)
    tail = np.array([e[1] for e in edges], dtype=np.int64)
    diff = lab[:, head, :] - lab[:, tail, :]                 # (T, nl, 2); NaN wherever an end point is unlabelled
    diff = np.where(np.isnan(diff), 0.0, diff)
    length_px = np.hypot(diff[..., 0], diff[..., 1]) * stride + 0.5 * stride
    upper = ws_max * length_px.max(axis=0) if len(edges) else np.zeros(0)
    weight = ws / (length_px.mean(axis=0) + 1e-20) if len(edges) else np.zeros(0)
    return weight.astype(np.float32), upper.astype(np.float32)


def _loss_args(dev, feed, cfg, edges, ws, ws_max, n_frames_total, n_visible_frames_total, nt, H, W, nj, pred=None,
               locref=None, with_locref=None, engine=None):
    """(dgp_loss_cfg, dgp_loss_batch, keep-alive list) from the reference's feed_dict values (fitdgp.py:797-815).
    When the feed carries ``visible_frame_within_batch`` instead of ``locref_map`` / ``locref_mask``, the two maps are
    generated on the device (Engine.locref_targets, the coord2map feeder) instead of being copied from the host."""
    def f32(a):   # host arrays are uploaded; CUDA tensors (feeds produced on the device) are used in place
        if isinstance(a, torch.Tensor):
            return a.to(device=dev, dtype=torch.float32).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device=dev)

    def i32(a):
        if isinstance(a, torch.Tensor):
            return a.to(device=dev, dtype=torch.int32).contiguous()
        return torch.as_tensor(np.ascontiguousarray(a, dtype=np.int32), device=dev)
    targets = f32(np.asarray(feed["targets"]).reshape(-1, nj, 2))
    vis, hid, vit = i32(feed["visible_marker_pl"]), i32(feed["hidden_marker_pl"]), i32(feed["visible_marker_in_targets_pl"])
    keep = [targets, vis, hid, vit]
    b = _lib.DgpLossBatch()
    b.nt, b.H, b.W = nt, H, W
    if pred is not None:
        b.pred_dev = pred.data_ptr()
    b.targets_dev, b.nv = targets.data_ptr(), targets.shape[0]
    if with_locref is None:
        with_locref = locref is not None
    if with_locref:
        if "locref_map" in feed:
            lm, lk = f32(feed["locref_map"]), f32(feed["locref_mask"])
        else:
            lm, lk = engine.locref_targets(feed["targets"], feed["visible_frame_within_batch"], nt, H, W,
                                           float(_get(cfg, "pos_dist_thresh", 17.0)), float(_get(cfg, "locref_stdev", 7.2801)))
        keep += [lm, lk]
        b.locref_map_dev, b.locref_mask_dev = lm.data_ptr(), lk.data_ptr()
        if locref is not None:
            b.locref_dev = locref.data_ptr()
    b.visible_marker_dev, b.nbv = vis.data_ptr(), vis.numel()
    b.hidden_marker_dev, b.nbh = hid.data_ptr(), hid.numel()
    b.visible_marker_in_targets_dev = vit.data_ptr()
    if len(edges) > 0:
        e, w1, w2 = i32(np.asarray(edges).reshape(-1, 2)), f32(ws), f32(ws_max)
        keep += [e, w1, w2]
        b.edges_dev, b.nl, b.ws_dev, b.ws_max_dev = e.data_ptr(), len(edges), w1.data_ptr(), w2.data_ptr()
    wt = float(_get(cfg, "wt", 0.0))
    if wt > 0:
        vf = f32(feed["vector_field_tf"])
        if vf.dim() != 3 or vf.shape[0] != nt - 1:
            raise ValueError("vector_field_tf must be (nt-1, Hin, Win) = (%d, ., .), got %s" % (nt - 1, tuple(vf.shape)))
        wb = f32(np.asarray(feed["wt_batch_pl"], dtype=np.float32) * np.asarray(feed["wt_batch_mask_pl"], dtype=np.float32))
        keep += [vf, wb]
        b.vector_field_dev, b.Hin, b.Win, b.wt_batch_dev = vf.data_ptr(), vf.shape[1], vf.shape[2], wb.data_ptr()
    c = _lib.DgpLossCfg(float(_get(cfg, "gamma", 1)), float(_get(cfg, "gauss_len", 1)), float(_get(cfg, "lengthscale", 1)),
                        wt, float(_get(cfg, "wt_max", 0)), float(_get(cfg, "wn_visible", 5)), float(_get(cfg, "wn_hidden", 3)),
                        float(_get(cfg, "locref_loss_weight", 0.05)), float(n_frames_total), float(n_visible_frames_total),
                        int(_get(cfg, "gm2", 1)), int(_get(cfg, "gm3", 3)),
                        0 if bool(_get(cfg, "locref_huber_loss", True)) else 1)
    return c, b, keep


def loss_forward(engine, pred, locref, feed, cfg, edges, ws, ws_max, n_frames_total, n_visible_frames_total,
                 backward=False, visible_only=False):
    """One call into dgp_loss_forward (or dgp_loss_backward).  pred/locref: CUDA tensors from Engine.forward; feed:
    reference feed_dict values.  With backward=True returns (losses, (grad_pred, grad_locref))."""
    dev = pred.device
    nt, H, W, nj = pred.shape
    c, b, keep = _loss_args(dev, feed, cfg, edges, ws, ws_max, n_frames_total, n_visible_frames_total, nt, H, W, nj, pred,
                            locref)
    out = torch.empty(6, dtype=torch.float32, device=dev)
    if backward:
        g_pred = torch.empty_like(pred)
        g_loc = torch.empty_like(locref) if locref is not None else None
        engine._check(engine.lib.dgp_loss_backward(engine.h, C.byref(c), C.byref(b), _ptr(out), _ptr(g_pred), _ptr(g_loc),
                                                   int(visible_only), _stream(dev)))
        vals = out.cpu().numpy()
        del keep
        return dict(zip(LOSS_KEYS, [np.float32(v) for v in vals])), (g_pred, g_loc)
    all_markers = torch.empty((nt * nj, 2), dtype=torch.float32, device=dev)
    engine._check(engine.lib.dgp_loss_forward(engine.h, C.byref(c), C.byref(b), _ptr(out), _ptr(all_markers), _stream(dev)))
    vals = out.cpu().numpy()
    del keep
    return dict(zip(LOSS_KEYS, [np.float32(v) for v in vals])), all_markers


def train_forward_backward(engine, frames, feed, cfg, edges, ws, ws_max, n_frames_total, n_visible_frames_total,
                           visible_only=False, sync=True):
    """Forward + loss + full backward of one fit_dgp batch (fitdgp.py:817-818 without the apply step): fills the
    engine's gradient buffer.  frames: uint8 CUDA tensor (nt,H,W,3).  Returns the loss dict (or the device tensor of the 6
    loss values when sync=False)."""
    from .engine import output_dims
    dev = frames.device
    nt, Hin, Win, _ = frames.shape
    _, (H, W) = output_dims(Hin, Win)
    engine.train_enable()
    c, b, keep = _loss_args(dev, feed, cfg, edges, ws, ws_max, n_frames_total, n_visible_frames_total, nt, H, W, engine.nj,
                            with_locref=engine.location_refinement, engine=engine)
    out = torch.empty(6, dtype=torch.float32, device=dev)
    engine._check(engine.lib.dgp_train_forward_backward(engine.h, _ptr(frames.contiguous()), nt, Hin, Win, C.byref(c), C.byref(b),
                                                        int(visible_only), _ptr(out), _stream(dev)))
    if not sync:
        engine._keep = keep  # device buffers must outlive the asynchronous launches
        return out
    vals = out.cpu().numpy()
    del keep
    return dict(zip(LOSS_KEYS, [np.float32(v) for v in vals]))


def dgp_loss(data_batcher, dgp_cfg, variables=None, device=None):
    """fitdgp.py:848-1144.  Returns (loss, total_loss, total_loss_visible, placeholders) of handles; evaluate them with
    ``TrainSession(...)``.run(fetches, feed_dict) using the reference's placeholder keys.

    The reference builds the graph here and restores ``init_weights`` into it afterwards (``restorer.restore``,
    fitdgp.py:689-720).  The engine needs its variables at construction, so they come from ``variables`` (a
    ``{tf_var_name: ndarray}`` dict, an ``.npz``, a TensorFlow checkpoint prefix, or ``'synthetic[:seed]'`` spelled out) or,
    when that is None, from ``dgp_cfg.init_weights`` -- the snapshot the reference would restore.  There is no default: a
    drop-in call never trains from random weights silently."""
    from .eval import load_variables
    S0 = np.asarray(data_batcher.S0)
    nj = int(data_batcher.nj)
    gm2, gm3 = int(_get(dgp_cfg, "gm2", 1)), int(_get(dgp_cfg, "gm3", 3))
    if gm2 not in (0, 1, 2) or gm3 not in (0, 3):
        raise Exception("Not implemented")  # fitdgp.py:1021, 1037
    if gm3 == 3 and gm2 == 0:
        # the reference dies here with NameError: pred_h_scaled1 is only defined for gm2 in {1, 2} (fitdgp.py:994-1033)
        raise NameError("name 'pred_h_scaled1' is not defined (gm3=3 needs gm2 in {1, 2}, fitdgp.py:1027)")
    if variables is None:
        variables = _get(dgp_cfg, "init_weights", None)
        if variables is None:
            raise ValueError("dgp_loss needs the network variables: pass variables=... or set dgp_cfg.init_weights to the "
                             "snapshot fit_dgp would restore (fitdgp.py:592, 720)")
    stride = float(_get(dgp_cfg, "stride", 8.0))
    ws, ws_max = spatial_clique_params([d.labels for d in data_batcher.datasets], S0, stride,
                                       float(_get(dgp_cfg, "ws", 1000.0)), float(_get(dgp_cfg, "ws_max", 1.2)))
    eng = Engine(nj, location_refinement=True, device=device, stride=stride,
                 locref_stdev=float(_get(dgp_cfg, "locref_stdev", 7.2801)),
                 mean_pixel=tuple(_get(dgp_cfg, "mean_pixel", (123.68, 116.779, 103.939))),
                 precision=_get(dgp_cfg, "precision", "fp16"))
    eng.load_weights(load_variables(variables, nj, True))
    graph = SimpleNamespace(engine=eng, cfg=dgp_cfg, edges=skeleton_edges(S0) if S0.shape[0] else [], ws=ws, ws_max=ws_max,
                            n_frames_total=float(data_batcher.n_frames_total),
                            n_visible_frames_total=float(data_batcher.n_visible_frames_total))
    keys = [k for k in LOSS_KEYS if (k != "wt_loss" or float(_get(dgp_cfg, "wt", 0)) > 0) and (k != "ws_loss" or S0.shape[0] > 0)]
    loss = {k: Handle(k, k) for k in keys}
    for hd in loss.values():
        hd.graph = graph
    total_loss_visible = Handle("total_loss_visible", "total_loss_visible")
    total_loss_visible.graph = graph
    placeholders = {k: Handle(k, k) for k in PLACEHOLDER_KEYS + EXTRA_PLACEHOLDER_KEYS}
    return loss, loss["total_loss"], total_loss_visible, placeholders


def momentum_train_op(total_loss, learning_rate=None, momentum=0.9, clip_norm=10.0, group=None):
    """fitdgp.py:706-713: ``optimizer = MomentumOptimizer(learning_rate, 0.9); grads = compute_gradients(total_loss,
    trainable_variables()); grads, _ = clip_by_global_norm(grads, 10.0); train_op = apply_gradients(...)``.
    ``total_loss`` is the handle to differentiate (``total_loss`` or ``total_loss_visible``); ``learning_rate`` a float or
    the placeholder handle returned by ``learning_rate_placeholder()``.  Returns the train_op handle."""
    op = Handle("train_op", "train_op")
    op.graph = total_loss.graph
    op.visible_only = total_loss.kind == "total_loss_visible"
    op.learning_rate, op.momentum, op.clip_norm, op.group = learning_rate, float(momentum), float(clip_norm), group
    return op


def learning_rate_placeholder():
    """``learning_rate = TF.placeholder(tf.float32, shape=[])`` (fitdgp.py:686)."""
    return Handle("learning_rate", "learning_rate")


class TrainSession:
    """``sess.run([loss, train_op], feed_dict)`` for the handles of ``dgp_loss`` / ``momentum_train_op``."""

    def __init__(self, placeholders):
        self.ph = placeholders

    def variables(self, graph, names):
        """Current values of trainable variables by TF name (what ``saver.save`` would write, fitdgp.py:830-839)."""
        return {n: graph.engine.get_variable(n) for n in names}

    def run(self, fetches, feed_dict):
        from . import dp
        feed = {}
        for k, v in feed_dict.items():
            if isinstance(k, str) and k in self.ph:     # plain placeholder names are accepted as keys too
                feed[k] = v
                continue
            for name, hd in self.ph.items():
                if k is hd:
                    feed[name] = v
        flat = []
        def collect(f):
            if isinstance(f, dict):
                for v in f.values():
                    collect(v)
            elif isinstance(f, (list, tuple)):
                for v in f:
                    collect(v)
            else:
                flat.append(f)
        collect(fetches)
        g = flat[0].graph
        frames = np.asarray(feed["inputs"])
        if frames.dtype != np.uint8:
            frames = np.clip(np.round(frames), 0, 255).astype(np.uint8)
        frames = torch.from_numpy(np.ascontiguousarray(frames)).to(g.engine.device)
        train_ops = [f for f in flat if f.kind == "train_op"]
        if train_ops:
            op = train_ops[0]
            lr = op.learning_rate
            if isinstance(lr, Handle):
                lr = [v for k, v in feed_dict.items() if k is lr]
                if not lr:
                    raise ValueError("feed_dict must provide the learning_rate placeholder")
                lr = lr[0]
            # nothing blocks the host between the backward, the gradient all-reduce (whose first bucket overlaps the rest of
            # the backward on a side stream) and the optimizer: the six loss values are read only after all three are enqueued
            out = train_forward_backward(g.engine, frames, feed, g.cfg, g.edges, g.ws, g.ws_max, g.n_frames_total,
                                         g.n_visible_frames_total, visible_only=op.visible_only, sync=False)
            dp.ensure_comm(g.engine, op.group)      # NCCL process group: the C handle owns the all-reduce from here on
            scale = dp.allreduce_gradients(g.engine, op.group)
            g.engine.optimizer_step(float(lr), op.momentum, op.clip_norm, scale)
            vals = dict(zip(LOSS_KEYS, [np.float32(v) for v in out.cpu().numpy()]))
        else:
            pred, locref = g.engine.forward(frames)
            vals, _ = loss_forward(g.engine, pred, locref, feed, g.cfg, g.edges, g.ws, g.ws_max, g.n_frames_total,
                                   g.n_visible_frames_total)
        vals["total_loss_visible"] = np.float32(vals["visible_loss_pred"] + vals["visible_loss_locref"])
        vals["train_op"] = None
        def build(f):
            if isinstance(f, dict):
                return {k: build(v) for k, v in f.items()}
            if isinstance(f, (list, tuple)):
                return type(f)(build(v) for v in f)
            return vals[f.kind]
        return build(fetches)
