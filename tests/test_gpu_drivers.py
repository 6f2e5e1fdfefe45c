"""The reference's driver functions on top of the hot path: fit_dgp / fit_dgp_labeledonly step loops (fitdgp.py:257-845) with a
synthetic MultiDataset-like batcher, and evaluate_dgp (eval.py:656-813) on a synthetic DeepLabCut project."""
import os
import pickle

import numpy as np
import pytest
import torch

from deepgraphpose_b200 import synthetic

pytestmark = pytest.mark.gpu
NJ, NT, HIN, WIN = 4, 12, 64, 96


class _Dataset:
    def __init__(self, labels, nx_out, ny_out):
        self.labels, self.nx_out, self.ny_out = labels, nx_out, ny_out


class _Batcher:
    """The slice of the reference's MultiDataset interface the step loops touch (dataset.py:1113-1180)."""

    def __init__(self, seed=0):
        self.frames, _ = synthetic.make_video(NT, HIN, WIN, NJ, seed=seed)
        self.H, self.W = 2 * -(-HIN // 16), 2 * -(-WIN // 16)
        rng = np.random.default_rng(seed)
        self.visible = np.array([1, 5, 9])
        self.labels = np.stack([rng.uniform(1, self.H - 2, (3, NJ)), rng.uniform(1, self.W - 2, (3, NJ))], axis=2)
        self.datasets = [_Dataset(self.labels, self.H, self.W)]
        self.S0 = np.array([[1, -1, 0, 0], [0, 1, -1, 0], [0, 0, 1, -1]], dtype=np.float64)
        self.nj = NJ
        self.n_frames_total, self.n_visible_frames_total = float(NT), 3.0
        self.resets = 0

    def reset(self):
        self.resets += 1

    def next_batch(self, _, dataset_i, visible_frame_batch, hidden_frame_batch):
        vis = np.sort(np.asarray(visible_frame_batch, dtype=np.int64))
        hid = np.sort(np.asarray(hidden_frame_batch, dtype=np.int64))
        allf = np.sort(np.concatenate([vis, hid]))
        joint_loc = self.labels[[int(np.where(self.visible == v)[0][0]) for v in vis]]
        pos = {f: k for k, f in enumerate(allf)}
        visible_marker = np.concatenate([pos[v] * NJ + np.arange(NJ) for v in vis]) if len(vis) else np.array([], dtype=np.int64)
        hidden_marker = np.concatenate([pos[h] * NJ + np.arange(NJ) for h in hid]) if len(hid) else np.array([], dtype=np.int64)
        vit = np.arange(len(vis) * NJ)
        wt_mask = np.ones(max(len(allf) - 1, 0))
        return (vis, hid, None, self.frames[allf].astype(np.float64), joint_loc, wt_mask, None,
                (visible_marker, hidden_marker, vit)), dataset_i


def _cfg(**kw):
    base = dict(batch_size=4, n_times_all_frames=1, lr=0.005, wt=0, wt_max=0, ws=1000.0, ws_max=1.2, wn_visible=5.0, wn_hidden=3.0,
                gamma=1.0, gm2=1, gm3=3, lengthscale=1.0, stride=8.0, locref_loss_weight=0.05, snapshot_prefix="snap")
    base.update(kw)
    return base


def test_fit_dgp_loop_trains_and_snapshots():
    from deepgraphpose_b200 import fitdgp
    b = _Batcher(seed=3)
    frame_lists = ([b.visible], [np.array([], dtype=np.int64)], [np.arange(NT)])
    saved = []
    windows = [np.array([0, 1, 2, 3, 0], dtype=np.int32), np.array([4, 5, 6, 7, 0], dtype=np.int32)] * 3
    hist = fitdgp.fit_dgp("synthetic:0", "/nonexistent", batch_size=4, gm2=1, gm3=3, wt=0, saveiters=8, displayiters=100,
                          data_batcher=b, dgp_cfg=_cfg(), frame_lists=frame_lists, batch_ind_all=windows,
                          snapshot_fn=lambda eng, name, it: saved.append((name, it)), verbose=False)
    assert b.resets == 1 and len(hist) == 6
    tot = [float(h["total_loss"]) for h in hist]
    assert all(np.isfinite(tot)) and tot[4] < tot[0] and tot[5] < tot[1]        # the same two batches, three passes each
    assert ("snap-step2-", 0) in saved and ("snap-step2-final-", 5) in saved and ("snap-step2-", 4) in saved
    # without a batch list the reference's gen_batch schedule is drawn
    hist2 = fitdgp.fit_dgp("synthetic:0", "/nonexistent", batch_size=4, gm2=1, gm3=3, maxiters=3, data_batcher=_Batcher(seed=3),
                           dgp_cfg=_cfg(), frame_lists=frame_lists, snapshot_fn=lambda *a: None, verbose=False)
    assert 1 <= len(hist2) <= 3
    with pytest.raises(NotImplementedError):
        fitdgp.fit_dgp("snapshot-step0-final--0", "/some/dlc/project")


def test_fit_dgp_labeledonly_follows_the_visible_loss():
    from deepgraphpose_b200 import fitdgp
    b = _Batcher(seed=4)
    frame_lists = ([b.visible], [np.array([], dtype=np.int64)], [np.arange(NT)])
    sched = [np.array([5, 0])] * 4
    hist = fitdgp.fit_dgp_labeledonly("synthetic:0", "/nonexistent", saveiters=100, data_batcher=b, dgp_cfg=_cfg(),
                                      frame_lists=frame_lists, batch_ind_all=sched, snapshot_fn=lambda *a: None, verbose=False)
    vis = [float(h["visible_loss_pred"]) + float(h["visible_loss_locref"]) for h in hist]
    assert len(hist) == 4 and all(float(h["hidden_loss_pred"]) == 0.0 for h in hist) and vis[-1] < vis[0]


def test_evaluate_dgp_on_a_synthetic_dlc_project(tmp_path):
    """evaluate_dgp(config.yaml, weights): RMSE DataFrame of a DeepLabCut-layout project (labelled PNGs, CollectedData csv,
    Documentation pickle with the train/test split) == distances computed by hand from evaluate_dgp_frames."""
    import pandas as pd
    import yaml
    from PIL import Image
    from deepgraphpose_b200 import eval as dgp_eval
    from deepgraphpose_b200.engine import Engine
    T, H, W = 5, 96, 128
    frames, tracks = synthetic.make_video(T, H, W, NJ, seed=9)
    names = ["a", "b", "c", "d"]
    proj = tmp_path / "proj"
    (proj / "labeled-data" / "vid").mkdir(parents=True)
    tsf = proj / "training-datasets" / "iteration-0" / "UnaugmentedDataSet_ReachJan1"
    tsf.mkdir(parents=True)
    index = []
    for t in range(T):
        rel = os.path.join("labeled-data", "vid", "img%03d.png" % t)
        Image.fromarray(frames[t]).save(proj / rel)
        index.append(rel)
    cols = pd.MultiIndex.from_product([["Me"], names, ["x", "y"]], names=["scorer", "bodyparts", "coords"])
    lab = np.stack([tracks[:, :, 1], tracks[:, :, 0]], axis=2).reshape(T, -1).astype(np.float64)
    lab[2, 0:2] = np.nan
    pd.DataFrame(lab, columns=cols, index=index).to_csv(tsf / "CollectedData_Me.csv")
    with open(tsf / "Documentation_data-Reach_80shuffle1.pickle", "wb") as f:
        pickle.dump([None, np.array([0, 1, 3]), np.array([2, 4]), 0.8], f)
    cfg = {"Task": "Reach", "date": "Jan1", "scorer": "Me", "iteration": 0, "TrainingFraction": [0.8], "bodyparts": names,
           "pcutoff": 0.1, "project_path": str(proj)}
    with open(proj / "config.yaml", "w") as f:
        yaml.safe_dump(cfg, f)
    dlc_cfg = {"num_joints": NJ, "all_joints_names": names, "stride": 8.0, "location_refinement": True, "net_type": "resnet_50"}
    W_ = synthetic.make_weights(NJ, seed=1)
    for mode in ("dlc", "dgp"):
        rmse = dgp_eval.evaluate_dgp(str(proj / "config.yaml"), W_, shuffle=1, loc_ref=True, loc_ref_calc=mode, dlc_cfg=dlc_cfg)
        assert rmse.shape == (T, NJ) and list(rmse.columns) == names
        eng = Engine(NJ)
        eng.load_weights(W_)
        pose = dgp_eval.evaluate_dgp_frames(eng, frames, True, mode).reshape(T, NJ, 3)
        want = np.hypot(lab.reshape(T, NJ, 2)[:, :, 0] - pose[:, :, 0], lab.reshape(T, NJ, 2)[:, :, 1] - pose[:, :, 1])
        got = rmse.values.astype(np.float64)
        assert np.isnan(got[2, 0]) and np.allclose(got[~np.isnan(want)], want[~np.isnan(want)], atol=1e-9)
        eng.close()
    r0 = dgp_eval.evaluate_dgp(str(proj / "config.yaml"), {k: v for k, v in W_.items() if "locref" not in k}, loc_ref=False, dlc_cfg=dlc_cfg)
    assert r0.shape == (T, NJ) and np.isfinite(r0.values[0]).all()
