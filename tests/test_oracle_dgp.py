"""Known-answer tests for the DGP part of the oracle (soft-argmax, read-outs, potentials, loss pieces)."""
import numpy as np
import torch

from oracle import dgp_loss, dgp_ops, pose_net


def test_gaussian_kernel_values():
    k = dgp_ops.make_gaussian_2d_kernel(1.0)
    assert k.shape == (3, 3)
    assert abs(k[1, 1].item() - 0.2041800) < 1e-6 and abs(k[0, 0].item() - 0.0751136) < 1e-6
    assert abs(k[0, 1].item() - 0.1238414) < 1e-6 and abs(k.sum().item() - 1) < 1e-6


def test_softargmax_delta_interior_and_border():
    H, W, nj = 12, 16, 2
    x = torch.full((1, H, W, nj), -1e4)
    x[0, 5, 7, 0] = 50.0     # interior delta -> exactly the peak (blur is symmetric)
    x[0, 0, 0, 1] = 50.0     # corner delta -> biased inwards by the zero-padded blur + renormalisation
    mu, sm = dgp_ops.argmax_2d_from_cm(x, nj, 1, 1)
    assert torch.allclose(mu[0, 0], torch.tensor([5.0, 7.0]), atol=1e-5)
    k = np.array([0.27406862, 0.45186276, 0.27406862])
    bias = k[2] / (k[1] + k[2])  # mass at index 0 and 1 only
    assert torch.allclose(mu[0, 1], torch.tensor([bias, bias], dtype=torch.float32), atol=1e-5)
    assert torch.allclose(sm.sum(dim=(1, 2)), torch.ones(1, nj), atol=1e-5)


def test_softargmax_uniform_is_centre():
    mu, _ = dgp_ops.argmax_2d_from_cm(torch.zeros(2, 10, 14, 3), 3, 1, 1)
    assert torch.allclose(mu[..., 0], torch.full((2, 3), 4.5), atol=1e-4)
    assert torch.allclose(mu[..., 1], torch.full((2, 3), 6.5), atol=1e-4)


def test_estimate_pose_readout_window_and_ties():
    H, W = 6, 8
    sc = np.full((1, H, W, 2), -3.0, np.float32)
    sc[0, 2, 3, 0] = 1.0
    sc[0, 3, 4, 0] = 2.0
    mu = np.array([[[2.4, 3.6], [4.0, 5.0]]], np.float32)
    markers, peaks, lik = dgp_ops.estimate_pose_readout(mu, sc)
    # window rows 2..3, cols 3..4 -> max at (3,4)
    assert peaks[0].tolist() == [3, 4]
    assert abs(lik[0] - 1 / (1 + np.exp(-2.0))) < 1e-6
    # integral mu -> 1x1 window; all-equal -> first index
    assert peaks[1].tolist() == [4, 5]
    # saturation ties resolve to the first (row-major) element
    sc[0, :, :, 1] = 40.0
    mu[0, 1] = [1.5, 1.5]
    _, peaks, lik = dgp_ops.estimate_pose_readout(mu, sc)
    assert peaks[1].tolist() == [1, 1] and lik[1] == 1.0


def test_argmax_pose_predict_locref_order():
    scmap = np.zeros((5, 7, 2), np.float32)
    scmap[3, 2, 0] = 0.9
    scmap[1, 6, 1] = 0.8
    locref = np.zeros((1, 5, 7, 4), np.float32)
    locref[0, 3, 2, 0:2] = [1.0, -2.0]   # (dx, dy) of joint 0
    sc, off = pose_net.extract_cnn_output(scmap[None], locref)
    pose, peaks = pose_net.argmax_pose_predict(sc, off, 8.0)
    assert peaks.tolist() == [[3, 2], [1, 6]]
    # x = col*8+4+dx*7.2801 ; y = row*8+4+dy*7.2801
    assert np.allclose(pose[0], [2 * 8 + 4 + 7.2801, 3 * 8 + 4 - 2 * 7.2801, 0.9], atol=1e-5)
    assert np.allclose(pose[1], [6 * 8 + 4, 1 * 8 + 4, 0.8], atol=1e-6)


def test_skeleton_and_temporal_algebra():
    mu = torch.tensor([[[0.0, 0.0], [3.0, 4.0]], [[1.0, 1.0], [1.0, 2.0]]])
    S0 = dgp_ops.skeleton_matrix([(0, 1)], 2)
    d = dgp_ops.skeleton_distances(mu, S0, 8.0)
    assert torch.allclose(d, torch.tensor([[40.0, 8.0]]))
    t = dgp_ops.temporal_distances(mu, 8.0)
    assert torch.allclose(t, torch.tensor([[8 * 2 ** 0.5, 8 * (4 + 4) ** 0.5]]))


def test_combine_all_marker_scatter():
    hid = torch.tensor([[1.0, 2.0], [3.0, 4.0]])
    vis = torch.tensor([[9.0, 8.0]])
    out = dgp_ops.combine_all_marker(hid, vis, torch.tensor([0, 3]), torch.tensor([2]), 2, 2)
    assert out.tolist() == [[1.0, 2.0], [0.0, 0.0], [9.0, 8.0], [3.0, 4.0]]


def test_spatial_clique_params_quirks():
    cfg = dgp_loss.default_dgp_cfg()
    S0 = dgp_ops.skeleton_matrix([(0, 1)], 2)
    labels = np.array([[[0.0, 0.0], [3.0, 4.0]], [[np.nan, np.nan], [1.0, 1.0]]])
    ws, ws_max = dgp_loss.spatial_clique_params(labels, S0, cfg)
    # limb lengths: 5*8+4 = 44 and (missing -> 0)*8+4 = 4 (the quirk: missing limbs count as stride/2)
    assert np.allclose(ws_max, [1.2 * 44])
    assert np.allclose(ws, [1000 / 24.0])


def test_huber_and_loss_runs():
    l = dgp_loss.huber_loss(torch.zeros(4), torch.tensor([0.5, -0.5, 2.0, -3.0]), torch.tensor([1.0, 1.0, 1.0, 0.0]))
    assert abs(l.item() - (0.125 + 0.125 + 1.5) / 3) < 1e-6
    # a full loss evaluation on a tiny batch with autograd
    g = torch.Generator().manual_seed(0)
    nt, H, W, nj = 3, 8, 10, 2
    pred = torch.randn(nt, H, W, nj, generator=g, requires_grad=True)
    loc = torch.randn(nt, H, W, 2 * nj, generator=g, requires_grad=True)
    S0 = dgp_ops.skeleton_matrix([(0, 1)], nj)
    labels = np.array([[[2.0, 3.0], [np.nan, np.nan]]])
    cfg = dgp_loss.default_dgp_cfg(wt=1.0)
    ws, ws_max = dgp_loss.spatial_clique_params(labels, S0, cfg)
    xg, yg = np.meshgrid(np.linspace(0, H - 1, H), np.linspace(0, W - 1, W))
    batch = {
        "targets": labels, "visible_marker_pl": np.array([2]), "hidden_marker_pl": np.array([0, 1, 3, 4, 5]),
        "visible_marker_in_targets_pl": np.array([0]), "nt_batch_pl": nt,
        "locref_map": np.zeros((nt, H, W, 2 * nj)), "locref_mask": np.ones((nt, H, W, 2 * nj)),
        "alpha_tf": np.array([xg, yg]).swapaxes(1, 2), "vector_field_tf": np.abs(np.random.default_rng(0).normal(size=(nt - 1, 20, 24))),
        "wt_batch_pl": np.ones(nt - 1), "wt_batch_mask_pl": np.array([1.0, 0.0]),
    }
    loss, total, total_vis = dgp_loss.dgp_loss_from_heads(pred, loc, batch, cfg, S0, ws, ws_max, 100, 10)
    assert set(loss) == {"visible_loss_pred", "hidden_loss_pred", "visible_loss_locref", "ws_loss", "wt_loss", "total_loss"}
    total.backward()
    assert torch.isfinite(pred.grad).all() and pred.grad.abs().sum() > 0
    assert torch.isfinite(total) and torch.isfinite(total_vis)
