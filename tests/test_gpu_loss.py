"""dgp_loss forward on the GPU (dgp_loss_forward) vs the oracle and vs the reference-generated golden losses."""
import os

import numpy as np
import pytest
import torch

from deepgraphpose_b200 import synthetic
from oracle import dgp_loss as oracle_loss
from oracle import dgp_ops

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
LOSS_REL_TOL = 1e-3   # BASELINE.json: DGP loss <= 1e-3 relative


def make_batch(rng, nt, H, W, nj, vis_frames, nan_joints=()):
    hid_frames = np.array([t for t in range(nt) if t not in vis_frames])
    vis_frames = np.array(vis_frames)
    labels = np.stack([rng.uniform(1, H - 2, (len(vis_frames), nj)), rng.uniform(1, W - 2, (len(vis_frames), nj))], axis=2)
    for (i, j) in nan_joints:
        labels[i, j, :] = np.nan
    nan_ind = sorted(int(nj * vis_frames[i] + j) for (i, j) in nan_joints)
    hidden = np.sort(list(np.array([hid_frames * nj + i for i in range(nj)], dtype=np.int64).flatten()) + nan_ind).astype(np.int64)
    vm0 = np.sort(np.array([vis_frames * nj + i for i in range(nj)], dtype=np.int64).flatten())
    visible = np.sort(np.setdiff1d(vm0, nan_ind))
    vit = np.nonzero(np.isin(vm0, visible))[0]
    lm = rng.normal(0, 0.7, (nt, H, W, 2 * nj))
    lk = (rng.uniform(size=(nt, H, W, 2 * nj)) < 0.2).astype(np.float64)
    xg, yg = np.meshgrid(np.linspace(0, H - 1, H), np.linspace(0, W - 1, W))
    return labels, {"targets": labels, "visible_marker_pl": visible, "hidden_marker_pl": hidden,
                    "visible_marker_in_targets_pl": vit, "nt_batch_pl": nt, "locref_map": lm, "locref_mask": lk,
                    "alpha_tf": np.array([xg, yg]).swapaxes(1, 2)}


@pytest.mark.parametrize("case", [dict(gm2=1, gm3=3, wt=0.0), dict(gm2=1, gm3=3, wt=1.0), dict(gm2=2, gm3=3, wt=0.0),
                                  dict(gm2=0, gm3=0, wt=0.0), dict(gm2=1, gm3=0, wt=2.0), dict(gm2=1, gm3=3, wt=0.0, novis=True)])
def test_loss_forward_matches_oracle(case):
    from deepgraphpose_b200 import fitdgp
    from deepgraphpose_b200.engine import Engine
    rng = np.random.default_rng(17)
    nt, H, W, nj = 5, 20, 24, 4
    novis = case.get("novis", False)
    labels, batch = make_batch(rng, nt, H, W, nj, [] if novis else [1, 3], () if novis else ((1, 2),))
    pred = (rng.standard_normal((nt, H, W, nj)) * 3).astype(np.float32)
    loc = rng.standard_normal((nt, H, W, 2 * nj)).astype(np.float32)
    edges = synthetic.chain_skeleton(nj)
    S0 = dgp_ops.skeleton_matrix(edges, nj)
    cfg = oracle_loss.default_dgp_cfg(gm2=case["gm2"], gm3=case["gm3"], wt=case["wt"], wt_max=0.5)
    lab_for_ws = labels if len(labels) else np.array([[[2.0, 3.0], [5.0, 9.0], [7.0, 4.0], [1.0, 1.0]]])
    ws, ws_max = oracle_loss.spatial_clique_params(lab_for_ws, S0, cfg)
    ws2, ws_max2 = fitdgp.spatial_clique_params([lab_for_ws], S0, 8.0, 1000.0, 1.2)
    assert np.allclose(ws, ws2, rtol=1e-6) and np.allclose(ws_max, ws_max2, rtol=1e-6)
    batch["vector_field_tf"] = np.abs(rng.normal(0, 1.0, (nt - 1, 8 * H, 8 * W)))
    batch["wt_batch_pl"] = np.ones(nt - 1) * case["wt"]
    batch["wt_batch_mask_pl"] = np.array([1.0, 0.0, 1.0, 1.0])
    with torch.no_grad():
        ref, ref_total, _ = oracle_loss.dgp_loss_from_heads(torch.from_numpy(pred), torch.from_numpy(loc), batch, cfg, S0,
                                                            ws, ws_max, 200, 20)
    eng = Engine(nj)
    got, all_markers = fitdgp.loss_forward(eng, torch.from_numpy(pred).cuda(), torch.from_numpy(loc).cuda(), batch, cfg,
                                           edges, ws, ws_max, 200, 20)
    for k, v in ref.items():
        assert abs(float(got[k]) - float(v)) <= LOSS_REL_TOL * max(abs(float(v)), 1e-6), (k, float(got[k]), float(v))
    eng.close()


@pytest.mark.parametrize("case", [dict(gm2=1, gm3=3), dict(gm2=2, gm3=3), dict(gm2=0, gm3=0), dict(gm2=1, gm3=0),
                                  dict(gm2=1, gm3=3, visible_only=True), dict(gm2=1, gm3=3, wt=20.0), dict(gm2=1, gm3=0, wt=60.0),
                                  dict(gm2=1, gm3=3, wt=40.0, wt_max=30.0)])
def test_loss_backward_matches_oracle_autograd(case):
    """d total_loss / d pred and d locref from dgp_loss_backward vs torch autograd through the oracle graph (gradients
    flow through the Gaussian targets into the soft-argmax, through the confidence max and the (1-c) weights)."""
    from deepgraphpose_b200 import fitdgp
    from deepgraphpose_b200.engine import Engine
    rng = np.random.default_rng(23)
    nt, H, W, nj = 4, 16, 20, 4
    labels, batch = make_batch(rng, nt, H, W, nj, [0, 2], ((0, 1),))
    pred = torch.from_numpy((rng.standard_normal((nt, H, W, nj)) * 2).astype(np.float32)).requires_grad_(True)
    loc = torch.from_numpy(rng.standard_normal((nt, H, W, 2 * nj)).astype(np.float32)).requires_grad_(True)
    edges = synthetic.chain_skeleton(nj)
    S0 = dgp_ops.skeleton_matrix(edges, nj)
    wt = case.get("wt", 0.0)
    cfg = oracle_loss.default_dgp_cfg(gm2=case["gm2"], gm3=case["gm3"], wt=wt, wt_max=case.get("wt_max", 0.0))
    if wt > 0:
        # temporal clique: smooth flow magnitude around 1 so that both branches of min(1/m, 1) occur, and box edges that
        # are clipped (markers near the border) as well as free ones
        yy, xx = np.meshgrid(np.arange(8 * H), np.arange(8 * W), indexing="ij")
        batch["vector_field_tf"] = np.stack([0.9 + 0.8 * np.sin(yy / (7.0 + t)) * np.cos(xx / (9.0 + 2 * t)) + 0.3 * rng.uniform(size=yy.shape)
                                             for t in range(nt - 1)])
        batch["wt_batch_pl"] = np.ones(nt - 1) * wt
        batch["wt_batch_mask_pl"] = np.array([1.0, 0.0, 1.0])
    ws, ws_max = oracle_loss.spatial_clique_params(labels, S0, cfg)
    ws_max = ws_max * 0.3   # make the skeleton hinge active for some limbs so its gradient is exercised
    ref, ref_total, ref_vis = oracle_loss.dgp_loss_from_heads(pred, loc, batch, cfg, S0, ws, ws_max, 200, 20)
    vis_only = case.get("visible_only", False)
    (ref_vis if vis_only else ref_total).backward()
    eng = Engine(nj)
    got, (g_pred, g_loc) = fitdgp.loss_forward(eng, pred.detach().cuda(), loc.detach().cuda(), batch, cfg, edges, ws, ws_max,
                                               200, 20, backward=True, visible_only=vis_only)
    assert abs(float(got["total_loss"]) - float(ref_total)) <= LOSS_REL_TOL * abs(float(ref_total))
    gp, gl = pred.grad, loc.grad
    assert (g_pred.cpu() - gp).abs().max().item() <= 2e-3 * gp.abs().max().item() + 1e-9, \
        ((g_pred.cpu() - gp).abs().max().item(), gp.abs().max().item())
    assert (g_loc.cpu() - gl).abs().max().item() <= 1e-4 * gl.abs().max().item() + 1e-12
    eng.close()


def test_loss_rejects_bad_flags():
    from deepgraphpose_b200 import fitdgp
    from deepgraphpose_b200._lib import DgpError
    from deepgraphpose_b200.engine import Engine
    rng = np.random.default_rng(3)
    labels, batch = make_batch(rng, 2, 8, 8, 4, [0])
    eng = Engine(4)
    pred = torch.zeros(2, 8, 8, 4, device="cuda")
    with pytest.raises(DgpError):   # gm3=3 with gm2=0: NameError in the reference (fitdgp.py:1026)
        fitdgp.loss_forward(eng, pred, None, batch, dict(gm2=0, gm3=3), [], [], [], 10, 2)
    with pytest.raises(DgpError):
        fitdgp.loss_forward(eng, pred, None, batch, dict(gm2=5, gm3=0), [], [], [], 10, 2)
    eng.close()


def test_dgp_loss_shim_vs_reference_golden():
    """dgp_loss(data_batcher, cfg) + TrainSession.run with the reference's placeholder keys vs the golden losses the
    reference's own dgp_loss produced (tests/golden/dgp_loss.npz), at the north_star tolerance (loss <= 1e-3 relative)."""
    from deepgraphpose_b200 import fitdgp
    with np.load(os.path.join(G, "dgp_loss.npz")) as z:
        g = {k: z[k] for k in z.files}
    nj, wseed, vseed, H, W, nt = [int(v) for v in g["meta"]]
    frames, _ = synthetic.make_video(nt, H, W, nj, seed=vseed)

    class DS:
        pass
    ds, db = DS(), DS()
    ds.labels = g["labels"]
    db.S0, db.nj, db.n_frames_total, db.n_visible_frames_total, db.datasets = g["S0"], nj, 120, 10, [ds]
    for tag, wt in (("wt0", 0.0), ("wt1", 1.0)):
        cfg = dict(stride=8.0, ws=1000.0, ws_max=1.2, wt=wt, wt_max=0.0, wn_visible=5.0, wn_hidden=3.0, gamma=1, gm2=1,
                   gm3=3, lengthscale=1, gauss_len=1, locref_loss_weight=0.05)
        loss, total_loss, total_loss_visible, ph = fitdgp.dgp_loss(db, cfg, variables="synthetic:%d" % wseed)
        sess = fitdgp.TrainSession(ph)
        feed = {ph["inputs"]: frames, ph["targets"]: g["labels"], ph["locref_map"]: g["locref_map"],
                ph["locref_mask"]: g["locref_mask"], ph["visible_marker_pl"]: g["visible_marker"],
                ph["hidden_marker_pl"]: g["hidden_marker"], ph["visible_marker_in_targets_pl"]: g["vis_in_targets"],
                ph["wt_batch_mask_pl"]: g["wt_batch_mask"], ph["vector_field_tf"]: g["vector_field"],
                ph["nt_batch_pl"]: nt, ph["wt_batch_pl"]: np.ones(nt - 1) * wt}
        vals, tl = sess.run([loss, total_loss], feed)
        assert set(vals) == {k[len(tag) + 1:] for k in g if k.startswith(tag + "_") and k != tag + "_total_loss_visible"}
        ref_total = float(g[tag + "_total_loss"])
        assert abs(float(tl) - ref_total) < LOSS_REL_TOL * ref_total, (float(tl), ref_total)
        assert abs(float(vals["ws_loss"]) - float(g[tag + "_ws_loss"])) < LOSS_REL_TOL * float(g[tag + "_ws_loss"])
