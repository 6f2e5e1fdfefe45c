"""tests/golden/feeders.npz: outputs of the REFERENCE'S OWN ``coord2map`` (src/deepgraphpose/dataset.py:246-271) and
``PoseDataset.compute_target_part_scoremap`` (DeepLabCut .../dataset/pose_defaultdataset.py:220-266).  The two modules cannot
be imported here (moviepy / skimage / tensorflow are absent), so the two function definitions are cut out of the reference
files with ``ast`` and executed unmodified.  Also: gen_idx_chunk, export_pose_like_dlc, learn_wt, calculate_motion_energy /
select_hidden_frames / get_neighboring_window and gen_batch, the same way.
Run in the build container only:  python tests/golden/make_golden_feeders.py"""
import ast
import os
import types

import numpy as np

REF = "/root/reference/src"
OUT = os.path.dirname(os.path.abspath(__file__))


def cut(path, name):
    src = open(path).read()
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.FunctionDef) and node.name == name:
            return ast.get_source_segment(src, node)
    raise KeyError(name)


if not hasattr(np, "asscalar"):
    np.asscalar = lambda a: a.item()  # removed in numpy >= 1.23; the reference pins an older numpy
ns = {"np": np, "arr": np.array, "cat": np.concatenate}
exec(cut(REF + "/DeepLabCut/deeplabcut/pose_estimation_tensorflow/dataset/pose_defaultdataset.py", "compute_target_part_scoremap"), ns)
exec(cut(REF + "/deepgraphpose/dataset.py", "coord2map"), ns)


class PData:
    """The attributes PoseDataset.__init__ sets (pose_defaultdataset.py:25-31) for the demo's pose_cfg.yaml."""

    def __init__(self, nj):
        self.cfg = types.SimpleNamespace(pos_dist_thresh=17, num_joints=nj)
        self.locref_scale = 1.0 / 7.2801
        self.stride = 8.0
        self.half_stride = 4.0

    def compute_scmap_weights(self, *a):
        return None


PData.compute_target_part_scoremap = ns["compute_target_part_scoremap"]

rng = np.random.default_rng(5)
out = {}
for tag, (n_vis, nj, nx, ny) in {"a": (3, 4, 12, 16), "b": (1, 5, 30, 38), "c": (2, 3, 94, 104)}.items():
    jl = np.stack([rng.uniform(-0.4, nx - 0.6, (n_vis, nj)), rng.uniform(-0.4, ny - 0.6, (n_vis, nj))], axis=2)
    jl[0, 1] = np.nan                       # missing label
    jl[-1, 0] = [0.0, 0.0]                  # corner
    jl[-1, nj - 1] = [nx - 1.0, ny - 1.0]   # opposite corner
    jl[0, 2] = [3.0, 5.0]                   # exactly on a cell centre: the 17 px circle passes through cell centres (8, 15)
    t, m = ns["coord2map"](PData(nj), jl, nx, ny, nj)
    out.update({tag + "_joint_loc": jl, tag + "_targets": t.astype(np.float32), tag + "_mask": m.astype(np.uint8),
                tag + "_dims": np.array([nx, ny, nj])})
# ---- marker index vectors: the reference's gen_idx_chunk (dataset.py:187-239) on the batches synthetic.make_training_batch draws
import sys
from itertools import chain

sys.path.insert(0, os.path.dirname(os.path.dirname(OUT)))
from deepgraphpose_b200 import synthetic  # noqa: E402

ns["chain"] = chain
exec(cut(REF + "/deepgraphpose/dataset.py", "gen_idx_chunk"), ns)
for tag, (nt, H, W, nj, vis, nan_joints, seed) in {"m0": (10, 94, 104, 4, [0, 3, 6], ((0, 1),), 100),
                                                  "m1": (5, 20, 24, 5, [1, 3], ((0, 0), (1, 4), (1, 2)), 3),
                                                  "m2": (4, 12, 16, 3, [], (), 9), "m3": (3, 12, 16, 3, [0, 1, 2], (), 2)}.items():
    labels, feed = synthetic.make_training_batch(nt, H, W, nj, vis, nan_joints, seed=seed)
    hid = np.array([t for t in range(nt) if t not in vis], dtype=np.int64)
    v, h, vit = ns["gen_idx_chunk"](np.array(vis, dtype=np.int64), hid, labels)
    out.update({tag + "_args": np.array([nt, H, W, nj, seed]), tag + "_vis": np.array(vis, dtype=np.int64),
                tag + "_nan": np.array(nan_joints, dtype=np.int64).reshape(-1, 2), tag + "_labels": labels,
                tag + "_visible_marker": np.asarray(v, dtype=np.int64), tag + "_hidden_marker": np.asarray(h, dtype=np.int64),
                tag + "_vit": np.asarray(vit, dtype=np.int64)})
# ---- DLC csv export: the reference's export_pose_like_dlc (models/eval.py:621-645); to_hdf needs pytables (absent) -> no-op
import tempfile

import pandas as pd

ns2 = {"np": np}
exec(cut(REF + "/deepgraphpose/models/eval.py", "export_pose_like_dlc"), ns2)
_to_hdf = pd.DataFrame.to_hdf
pd.DataFrame.to_hdf = lambda self, *a, **k: None
lab = {"x": rng.uniform(0, 800, (4, 3)), "y": rng.uniform(0, 700, (4, 3)), "likelihoods": rng.uniform(0, 1, (4, 3))}
lab["x"][1, 2] = np.nan
with tempfile.TemporaryDirectory() as d:
    ns2["export_pose_like_dlc"](lab, "snapshot-step2-final--0", ["hand", "finger", "elbow"], os.path.join(d, "vid_labeled"))
    csv_text = open(os.path.join(d, "vid_labeled.csv")).read()
pd.DataFrame.to_hdf = _to_hdf
out.update({"csv_x": lab["x"], "csv_y": lab["y"], "csv_l": lab["likelihoods"], "csv_text": np.array(csv_text)})
# ---- learn_wt (models/fitdgp_util.py:454-467): Farneback flow magnitude, the reference's own function on 3 synthetic frames
import cv2

ns3 = {"np": np, "cv2": cv2}
exec(cut(REF + "/deepgraphpose/models/fitdgp_util.py", "learn_wt"), ns3)
vid, _ = synthetic.make_video(3, 64, 96, 3, seed=77)
vf = ns3["learn_wt"](vid.astype(np.float64))
out.update({"flow_seed": np.array(77), "flow_field": vf.astype(np.float32)})
# ---- hidden-frame selection: the reference's calculate_motion_energy / select_hidden_frames / get_neighboring_window
# (dataset.py:29-119) on a synthetic clip (a stand-in for moviepy's VideoFileClip yields the frames) -- note the uint8 wrap
class _Clip:
    def __init__(self, frames):
        self._f = frames
        self.fps, self.duration = 10.0, len(frames) / 10.0

    def iter_frames(self):
        return iter(self._f)

    def close(self):
        pass


me_vid, _ = synthetic.make_video(24, 48, 64, 3, seed=21)
ns4 = {"np": np, "VideoFileClip": lambda path: _Clip(me_vid)}
for fn in ("calculate_motion_energy", "make_neighboring_window", "get_neighboring_window", "select_hidden_frames"):
    exec(cut(REF + "/deepgraphpose/dataset.py", fn), ns4)
me = ns4["calculate_motion_energy"]("clip.mp4")
order = np.argsort(me)[::-1]
pv = np.array([3, 15])
import contextlib, io
with contextlib.redirect_stdout(io.StringIO()):
    sel = [ns4["select_hidden_frames"](2, pv, order, len(me), nmax, jump) for nmax, jump in ((14, None), (20, 0), (8, None), (24, 1))]
out.update({"me_seed": np.array(21), "me_values": me, "me_pv": pv, "me_order": order, "me_windowed": ns4["get_neighboring_window"](pv, 2, len(me))})
for k, v in enumerate(sel):
    out["me_sel%d" % k] = np.asarray(v, dtype=np.int64)
# ---- gen_batch (models/fitdgp_util.py:146-202) with both RNGs seeded
import random

np.int = int    # removed from numpy; the reference pins an older one
ns5 = {"np": np, "random": random}
exec(cut(REF + "/deepgraphpose/models/fitdgp_util.py", "gen_batch"), ns5)
gb_cfg = types.SimpleNamespace(batch_size=4, n_times_all_frames=3)
gb_vis = [np.array([2, 9, 17]), np.array([1])]
gb_hid = [np.array([5, 6, 30, 31]), np.array([], dtype=np.int64)]
gb_all = [np.array([0, 1, 2, 3, 4, 7, 8, 9, 10, 11, 15, 16, 17, 18, 19]), np.array([0, 1, 2])]
np.random.seed(11)
random.seed(12)
with contextlib.redirect_stdout(io.StringIO()):
    gb = ns5["gen_batch"](gb_vis, gb_hid, gb_all, gb_cfg, 500)
out["gen_batch_lens"] = np.array([len(b) for b in gb])
out["gen_batch_flat"] = np.concatenate(gb)
np.savez_compressed(os.path.join(OUT, "feeders.npz"), **out)
print({k: v.shape for k, v in out.items()})
