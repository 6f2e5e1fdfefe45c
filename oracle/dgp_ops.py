"""DGP graph helpers restated on torch CPU / numpy (oracle only).

Follows /root/reference/src/deepgraphpose/models/fitdgp_util.py
(``make_gaussian_2d_kernel`` :281-286, ``apply_gaussian_2d_kernel`` :289-315,
``make_2Dgrids`` :318-339, ``argmax_2d_from_cm`` :342-402, ``combine_all_marker`` :232-272)
and /root/reference/src/deepgraphpose/models/eval.py (``estimate_pose`` read-out :328-357,
``evaluate_dgp`` :744-790).
"""
import numpy as np
import torch
import torch.nn.functional as F

from .pose_net import LOCREF_STDEV, STRIDE


def make_gaussian_2d_kernel(sigma, truncate=1.0, dtype=torch.float32):
    """fitdgp_util.py:281-286."""
    radius = int(sigma * truncate)
    x = torch.arange(-radius, radius + 1, dtype=dtype)
    k = torch.exp(-0.5 * torch.square(x / sigma))
    k = k / k.sum()
    return k[:, None] * k


def apply_gaussian_2d_kernel(image, gauss_len, nj):
    """fitdgp_util.py:289-315: zero-pad gauss_len per side, depthwise VALID conv."""
    kernel = make_gaussian_2d_kernel(gauss_len, dtype=image.dtype)
    pad = int(gauss_len)
    x = image.permute(0, 3, 1, 2)
    x = F.pad(x, (pad, pad, pad, pad))
    w = kernel[None, None].repeat(nj, 1, 1, 1)
    y = F.conv2d(x, w, groups=nj)
    return y.permute(0, 2, 3, 1)


def argmax_2d_from_cm(tensor, nj, gamma=1, gauss_len=2, th=None):
    """fitdgp_util.py:342-402. tensor (N,H,W,C) -> ((N,C,2) (row,col), (N,H,W,C))."""
    N, H, W, C = tensor.shape
    features = tensor.permute(0, 3, 1, 2)
    flat = features.reshape(N * C, -1)
    sm = torch.softmax(flat * gamma, dim=1)
    sm = sm.reshape(N, C, H, W).permute(0, 2, 3, 1)
    sm = apply_gaussian_2d_kernel(sm, gauss_len, nj)
    s = sm.sum(dim=(1, 2), keepdim=True)
    sm = sm / (s + sm.new_tensor(1e-100))  # 1e-100 == 0 in fp32, as in TF
    if th is not None:
        st = sm.permute(0, 3, 1, 2).reshape(-1, H, W)
        mst = st.amax(dim=(1, 2), keepdim=True)
        st = torch.where(st < mst * th, torch.zeros_like(st), st)
        sm = st.reshape(-1, nj, H, W).permute(0, 2, 3, 1)
        s = sm.sum(dim=(1, 2), keepdim=True)
        sm = sm / (s + sm.new_tensor(1e-100))
    sm0 = sm
    rows = torch.arange(H, dtype=tensor.dtype).view(1, H, 1, 1)
    cols = torch.arange(W, dtype=tensor.dtype).view(1, 1, W, 1)
    mu_r = (sm * rows).sum(dim=(1, 2))
    mu_c = (sm * cols).sum(dim=(1, 2))
    return torch.stack([mu_r, mu_c], dim=-1), sm0


def estimate_pose_readout(mu_n_batch, scmap_np):
    """eval.py:329-343 for ONE frame (numpy, literal).

    mu_n_batch (1,nj,2) float32, scmap_np (1,H,W,nj) float32 logits.
    Returns markers (nj,2) float64, mu_likelihoods (nj,2) int, likelihoods (nj,) float64.
    """
    nj = mu_n_batch.shape[1]
    markers = np.zeros((nj, 2))
    mu_likelihoods = np.zeros((nj, 2)).astype("int")
    likelihoods = np.zeros((nj,))
    offset_mu_jj = 0
    markers[:] = mu_n_batch[0]
    softmaxtensor = scmap_np[0]
    with np.errstate(over="ignore", invalid="ignore"):
        for jj_idx in range(nj):
            mu_jj = markers[jj_idx]
            ends_floor = np.floor(mu_jj).astype("int") - offset_mu_jj
            ends_ceil = np.ceil(mu_jj).astype("int") + 1 + offset_mu_jj
            sigmoid_pred_np_jj = np.exp(softmaxtensor[:, :, jj_idx]) / (np.exp(softmaxtensor[:, :, jj_idx]) + 1)
            spred_centered = sigmoid_pred_np_jj[ends_floor[0]:ends_ceil[0], ends_floor[1]:ends_ceil[1]]
            mu_likelihoods[jj_idx] = np.unravel_index(np.argmax(spred_centered), spred_centered.shape)
            mu_likelihoods[jj_idx] += [ends_floor[0], ends_floor[1]]
            likelihoods[jj_idx] = sigmoid_pred_np_jj[int(mu_likelihoods[jj_idx][0]), int(mu_likelihoods[jj_idx][1])]
    return markers, mu_likelihoods, likelihoods


def estimate_pose_xy(markers, stride=STRIDE, scale_x=1.0, scale_y=1.0):
    """eval.py:352-357: markers (T,nj,2)(row,col) -> x (T,nj), y (T,nj)."""
    xr = markers[:, :, 1] * stride + 0.5 * stride
    yr = markers[:, :, 0] * stride + 0.5 * stride
    return xr * scale_x, yr * scale_y


def evaluate_dgp_pose_dgp_branch(st, lr, stride=STRIDE, locref_stdev=LOCREF_STDEV):
    """eval.py:751-785 ('dgp' loc_ref_calc) for one frame (numpy, literal incl. the row/col quirk).

    st (1,H,W,nj) blurred softmax, lr (1,H,W,2nj) raw locref. Returns pose (nj,3) = (x,y,1).
    """
    locref = np.squeeze(lr).copy()
    shape = locref.shape
    locref = np.reshape(locref, (shape[0], shape[1], -1, 2))
    locref *= locref_stdev
    nx, ny = shape[0], shape[1]
    nj = locref.shape[2]
    xg, yg = np.meshgrid(np.linspace(0, nx - 1, nx), np.linspace(0, ny - 1, ny))
    alpha = np.array([xg, yg]).swapaxes(1, 2)
    pose_hard_st1 = []
    for joint_idx in range(nj):
        st_j = np.expand_dims(st[0, :, :, joint_idx], 0)
        lr_j = np.transpose(locref[:, :, joint_idx, :], [2, 0, 1])
        spatial_soft_argmax = np.sum(np.sum(st_j * alpha, 1), 1) * stride + 0.5 * stride
        offset = np.sum(np.sum(st_j * lr_j, 1), 1)
        pose_hard_st1.append(np.hstack(((spatial_soft_argmax + offset)[::-1])))
    return np.hstack((np.array(pose_hard_st1), np.ones((nj, 1))))


def evaluate_dgp_pose_noloc(mu_n_batch, stride=STRIDE):
    """eval.py:787-790."""
    nj = mu_n_batch.shape[1]
    pose = mu_n_batch * stride + 0.5 * stride
    return np.hstack([pose[0, :, ::-1], np.ones((nj, 1))])


def combine_all_marker(targets_pred_hidden_marker, targets_visible_marker, hidden_marker, visible_marker, nj, nt):
    """fitdgp_util.py:232-272 (tf.scatter_nd accumulates duplicates; index_add does the same)."""
    n = nt * nj
    mu = torch.zeros(n, 2, dtype=targets_pred_hidden_marker.dtype)
    mu = mu.index_add(0, hidden_marker.long(), targets_pred_hidden_marker)
    yv = torch.zeros(n, 2, dtype=targets_pred_hidden_marker.dtype)
    yv = yv.index_add(0, visible_marker.long(), targets_visible_marker)
    return mu + yv


def skeleton_matrix(edges, nj):
    """fitdgp.py:607-617: S0 (nl, nj) with +1 / -1 per limb."""
    S0 = np.zeros((len(edges), nj))
    for s, (a, b) in enumerate(edges):
        S0[s, a] = 1
        S0[s, b] = -1
    return S0


def skeleton_distances(mu, S0, stride=STRIDE):
    """fitdgp.py:1063-1069: d[l,t] = || S (mu_t*stride + stride/2) ||_2 ; mu (T,nj,2) torch."""
    T, nj, _ = mu.shape
    S = torch.as_tensor(S0, dtype=mu.dtype)
    nl = S.shape[0]
    tm = mu.permute(1, 2, 0).reshape(nj, -1) * stride + 0.5 * stride
    d = torch.sqrt(torch.sum(torch.square((S @ tm).reshape(nl, 2, -1)), dim=1))
    return d  # (nl, T)


def temporal_distances(mu, stride=STRIDE):
    """fitdgp.py:1080-1083: delta[t,j] = || mu_t - mu_{t+1} ||_2 in image pixels; mu (T,nj,2)."""
    tm = mu * stride + 0.5 * stride
    return torch.sqrt(torch.sum(torch.square(tm[:-1] - tm[1:]), dim=2))
