"""The DGP loss graph restated on torch CPU (oracle only; autograd gives the reference gradients).

Follows /root/reference/src/deepgraphpose/models/fitdgp.py ``dgp_loss`` :848-1144
(host precompute :874-892, placeholders :896-933, targets :947-976, CE losses :978-1039,
locref Huber :1041-1055, spatial clique :1062-1076, temporal clique :1078-1124), the
Huber loss at /root/reference/src/DeepLabCut/deeplabcut/pose_estimation_tensorflow/nnet/losses.py:16-45
and the optimizer at fitdgp.py:706-713.
"""
from types import SimpleNamespace

import numpy as np
import torch

from . import dgp_ops, tf_ops
from .pose_net import STRIDE


def default_dgp_cfg(**kw):
    """Hyper-parameters written onto dlc_cfg by fit_dgp (fitdgp.py:637-654) + demo flags."""
    cfg = dict(
        stride=STRIDE, ws=1000.0, ws_max=1.2, wt=0.0, wt_max=0.0, wn_visible=5.0, wn_hidden=3.0,
        gamma=1.0, gauss_len=1.0, lengthscale=1.0, gm2=1, gm3=3, locref_huber_loss=True,
        locref_loss_weight=0.05, lr=0.005,
    )
    cfg.update(kw)
    return SimpleNamespace(**cfg)


def spatial_clique_params(joint_loc_full, S0, cfg):
    """fitdgp.py:874-892 (float64 numpy, literal). joint_loc_full (Nvis,nj,2) with NaNs."""
    nj = S0.shape[1]
    joint_loc_full = np.asarray(joint_loc_full, dtype=np.float64).reshape(-1, nj, 2)
    joint_loc_full1 = np.copy(joint_loc_full).swapaxes(1, 2).reshape(-1, nj)
    joint_loc_full1[np.isnan(joint_loc_full1)] = 1e10
    limb_full = np.matmul(joint_loc_full1, S0.T)
    limb_full[np.abs(limb_full) > 1e5] = 0
    limb_full = np.reshape(limb_full, [joint_loc_full.shape[0], 2, -1])
    limb_full = np.sqrt(np.sum(np.square(limb_full), 1))
    limb_full = limb_full.T * cfg.stride + cfg.stride / 2
    ws_max = np.max(np.nan_to_num(limb_full), 1) * cfg.ws_max
    with np.errstate(invalid="ignore", divide="ignore"):
        limb_mean = np.true_divide(limb_full.sum(1), (limb_full != 0).sum(1))
    ws = 1 / (np.nan_to_num(limb_mean) + 1e-20) * cfg.ws
    return ws, ws_max


def huber_loss(labels, predictions, weight, k=1.0):
    """losses.py:16-45."""
    diff = predictions - labels
    abs_diff = torch.abs(diff)
    losses = torch.where(abs_diff < k, 0.5 * torch.square(diff), k * abs_diff - 0.5 * k ** 2)
    return tf_ops.compute_weighted_loss(losses, weight)


def crop_and_resize_mean(image, boxes, box_ind, crop_h, crop_w):
    """mean over tf.image.crop_and_resize(image[...,None], boxes, box_ind, [crop_h,crop_w]) (bilinear, extrap 0).

    image (B,H,W); boxes (K,4) = (y1,x1,y2,x2) normalised; returns (K,) means.
    """
    B, H, W = image.shape
    out = []
    for k in range(boxes.shape[0]):
        y1, x1, y2, x2 = [boxes[k, i] for i in range(4)]
        img = image[int(box_ind[k])]
        if crop_h > 1:
            ys = y1 * (H - 1) + torch.arange(crop_h, dtype=image.dtype) * ((y2 - y1) * (H - 1) / (crop_h - 1))
        else:
            ys = (0.5 * (y1 + y2) * (H - 1)).reshape(1)
        if crop_w > 1:
            xs = x1 * (W - 1) + torch.arange(crop_w, dtype=image.dtype) * ((x2 - x1) * (W - 1) / (crop_w - 1))
        else:
            xs = (0.5 * (x1 + x2) * (W - 1)).reshape(1)
        vy = (ys >= 0) & (ys <= H - 1)
        vx = (xs >= 0) & (xs <= W - 1)
        y0 = torch.floor(ys).clamp(0, H - 1).long()
        y1i = torch.ceil(ys).clamp(0, H - 1).long()
        x0 = torch.floor(xs).clamp(0, W - 1).long()
        x1i = torch.ceil(xs).clamp(0, W - 1).long()
        ly = (ys - torch.floor(ys))[:, None]
        lx = (xs - torch.floor(xs))[None, :]
        tl = img[y0][:, x0]
        tr = img[y0][:, x1i]
        bl = img[y1i][:, x0]
        br = img[y1i][:, x1i]
        top = tl + (tr - tl) * lx
        bot = bl + (br - bl) * lx
        val = top + (bot - top) * ly
        val = val * (vy[:, None] & vx[None, :]).to(image.dtype)
        out.append(val.mean())
    return torch.stack(out)


def dgp_loss_from_heads(pred, locref_pred, batch, cfg, S0, ws, ws_max, n_frames_total, n_visible_frames_total):
    """The graph from fitdgp.py:947-1128 given the head outputs.

    pred (nt,H,W,nj) logits, locref_pred (nt,H,W,2nj); ``batch`` holds the placeholder feeds
    (keys as in fitdgp.py:1130-1142, numpy arrays).  Returns (loss dict, total_loss, total_loss_visible).
    """
    f32 = torch.float32
    nj = S0.shape[1]
    nl = S0.shape[0]
    nt, nx_out, ny_out, _ = pred.shape
    n_hidden_frames_total = n_frames_total - n_visible_frames_total

    targets = torch.as_tensor(np.asarray(batch["targets"], dtype=np.float32)).reshape(-1, nj, 2)
    targets_nonan = torch.where(torch.isnan(targets), torch.zeros_like(targets), targets)
    visible_marker = torch.as_tensor(np.asarray(batch["visible_marker_pl"], dtype=np.int64))
    hidden_marker = torch.as_tensor(np.asarray(batch["hidden_marker_pl"], dtype=np.int64))
    visible_marker_in_targets = torch.as_tensor(np.asarray(batch["visible_marker_in_targets_pl"], dtype=np.int64))
    nt_batch = int(batch["nt_batch_pl"])
    locref_map = torch.as_tensor(np.asarray(batch["locref_map"], dtype=np.float32))
    locref_mask = torch.as_tensor(np.asarray(batch["locref_mask"], dtype=np.float32))
    alpha = torch.as_tensor(np.asarray(batch["alpha_tf"], dtype=np.float32))

    targets_pred, _ = dgp_ops.argmax_2d_from_cm(pred, nj, cfg.gamma, cfg.gauss_len)
    targets_pred_marker = targets_pred.reshape(-1, 2)
    targets_pred_hidden_marker = targets_pred_marker[hidden_marker]
    targets_visible_marker = targets_nonan.reshape(-1, 2)[visible_marker_in_targets]
    targets_all_marker = dgp_ops.combine_all_marker(
        targets_pred_hidden_marker, targets_visible_marker, hidden_marker, visible_marker, nj, nt_batch)

    target_expand = targets_all_marker[:, :, None, None]
    alpha_expand = alpha[None]
    targets_gauss = torch.exp(-torch.sum(torch.square(alpha_expand - target_expand), dim=1) / (2 * (cfg.lengthscale ** 2)))
    gauss_max = targets_gauss.amax(dim=(1, 2)) + 1e-5
    targets_gauss = targets_gauss / gauss_max[:, None, None]  # (nt*nj, H, W)

    nbh = torch.tensor(float(hidden_marker.shape[0]), dtype=f32)
    nbv = torch.tensor(float(visible_marker.shape[0]), dtype=f32)
    nbv = torch.sign(nbv) * nbv + (1 - torch.sign(nbv)) * nbh

    targets_gauss_v = targets_gauss[visible_marker]
    targets_gauss_h = targets_gauss[hidden_marker]
    pred_t = pred.permute(0, 3, 1, 2).reshape(-1, nx_out, ny_out)
    pred_v = pred_t[visible_marker]
    pred_h = pred_t[hidden_marker]

    if cfg.gm2 in (1, 2):
        pred_h_sigmoid = torch.sigmoid(pred_h)
        if pred_h_sigmoid.shape[0] > 0:
            pgm_h1 = pred_h_sigmoid.amax(dim=(1, 2))
        else:
            pgm_h1 = pred_h_sigmoid.new_zeros((0,))
        pgm_h2 = pgm_h1[:, None, None]
        if cfg.gm2 == 1:
            targets_gauss_h = targets_gauss_h * pgm_h2
        pred_h_scaled = pred_h_sigmoid * pgm_h2
        pred_h_scaled1 = -torch.log(1 - pred_h_scaled + 1e-20) + torch.log(pred_h_scaled + 1e-20)
    elif cfg.gm2 != 0:
        raise Exception("Not implemented")

    loss = {}
    loss["visible_loss_pred"] = tf_ops.compute_weighted_loss(
        tf_ops.sigmoid_cross_entropy_with_logits(targets_gauss_v, pred_v), 1.0)
    ratio = (n_visible_frames_total / n_hidden_frames_total) * nbh / nbv * cfg.wn_hidden / cfg.wn_visible
    if cfg.gm3 == 3:
        if cfg.gm2 == 0:
            raise NameError("pred_h_scaled1 undefined for gm2=0 (reference quirk, fitdgp.py:1026-1027)")
        loss["hidden_loss_pred"] = tf_ops.compute_weighted_loss(
            tf_ops.sigmoid_cross_entropy_with_logits(targets_gauss_h, pred_h_scaled1), (1 - pgm_h2)) * ratio
    elif cfg.gm3 == 0:
        loss["hidden_loss_pred"] = tf_ops.compute_weighted_loss(
            tf_ops.sigmoid_cross_entropy_with_logits(targets_gauss_h, pred_h), 1.0) * ratio
    else:
        raise Exception("Not implemented")
    total_loss = loss["visible_loss_pred"] + loss["hidden_loss_pred"]

    lp = locref_pred.permute(0, 3, 1, 2).reshape(-1, 2, nx_out, ny_out)[visible_marker]
    lm = locref_map.permute(0, 3, 1, 2).reshape(-1, 2, nx_out, ny_out)[visible_marker]
    lk = locref_mask.permute(0, 3, 1, 2).reshape(-1, 2, nx_out, ny_out)[visible_marker]
    if cfg.locref_huber_loss:
        ll = huber_loss(lm, lp, lk)
    else:
        ll = tf_ops.compute_weighted_loss(torch.square(lp - lm), lk)
    loss["visible_loss_locref"] = cfg.locref_loss_weight * ll
    total_loss = total_loss + loss["visible_loss_locref"]

    targets_all_marker_3c = targets_all_marker.reshape(nt_batch, nj, -1)
    if nl > 0:
        d = dgp_ops.skeleton_distances(targets_all_marker_3c, S0, cfg.stride)
        ws_max_t = torch.as_tensor(ws_max, dtype=f32)[:, None]
        ws_t = torch.as_tensor(ws, dtype=f32).reshape(-1, 1)
        d_th = torch.relu(d - ws_max_t) + ws_max_t
        l_ws = torch.sum(d_th * ws_t) / float(nx_out) / float(ny_out)
        l_ws = l_ws * n_visible_frames_total / nbv / (n_visible_frames_total + n_hidden_frames_total) / cfg.wn_visible
        loss["ws_loss"] = l_ws
        total_loss = total_loss + l_ws

    if cfg.wt > 0:
        vf = torch.as_tensor(np.asarray(batch["vector_field_tf"], dtype=np.float32))
        wt_batch = torch.as_tensor(np.asarray(batch["wt_batch_pl"], dtype=np.float32)) * \
            torch.as_tensor(np.asarray(batch["wt_batch_mask_pl"], dtype=np.float32))
        tm = targets_all_marker_3c * cfg.stride + 0.5 * cfg.stride
        t0, t1 = tm[:-1], tm[1:]
        time_dif0 = torch.sqrt(torch.sum(torch.square(t0 - t1), 2))
        nx_in, ny_in = float(vf.shape[1]), float(vf.shape[2])
        r0, c0 = t0[:, :, 0].reshape(-1), t0[:, :, 1].reshape(-1)
        r1, c1 = t1[:, :, 0].reshape(-1), t1[:, :, 1].reshape(-1)
        window = 10
        rmin = torch.clamp(torch.minimum(r0, r1) - window, min=0.0)
        rmax = torch.clamp(torch.maximum(r0, r1) + window, max=nx_in)
        cmin = torch.clamp(torch.minimum(c0, c1) - window, min=0.0)
        cmax = torch.clamp(torch.maximum(c0, c1) + window, max=ny_in)
        boxes = torch.stack((rmin / nx_in, cmin / ny_in, rmax / nx_in, cmax / ny_in), dim=1)
        box_ind = torch.arange(nt_batch - 1).repeat_interleave(nj)
        meanflow = crop_and_resize_mean(vf, boxes, box_ind, int(nx_in), int(ny_in)).reshape(-1, nj)
        inv = 1 / (meanflow + 1e-10)
        inv = torch.clamp(inv, max=1.0)
        inv = torch.exp(torch.log(inv) * 3)
        inv = torch.clamp(inv, max=1.0)
        inv = inv * wt_batch.reshape(-1, 1) / float(nx_out) / float(ny_out)
        d_wt = (torch.relu(time_dif0 - cfg.wt_max) + cfg.wt_max) * inv
        l_wt = torch.sqrt(torch.sum(torch.square(d_wt)))
        l_wt = l_wt * n_visible_frames_total / nbv / (n_visible_frames_total + n_hidden_frames_total) / cfg.wn_visible
        loss["wt_loss"] = l_wt
        total_loss = total_loss + l_wt

    loss["total_loss"] = total_loss
    total_loss_visible = loss["visible_loss_pred"] + loss["visible_loss_locref"]
    return loss, total_loss, total_loss_visible


def momentum_step(params, grads, accums, lr=0.005, momentum=0.9, clip_norm=10.0):
    """fitdgp.py:706-713: clip_by_global_norm(10) then tf.train.MomentumOptimizer (accum = m*accum + g; w -= lr*accum)."""
    gnorm = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    scale = clip_norm / torch.clamp(gnorm, min=clip_norm)
    new_p, new_a = [], []
    for p, g, a in zip(params, grads, accums):
        g = g * scale
        a = momentum * a + g
        new_a.append(a)
        new_p.append(p - lr * a)
    return new_p, new_a, gnorm
