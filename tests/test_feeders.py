"""coord2map / compute_target_part_scoremap (host feeders of the training step): oracle vs the golden vectors produced by
the reference's own functions (CPU), and the CUDA feeder kernel vs both (GPU, bit-exact)."""
import os

import numpy as np
import pytest

from oracle import feeders

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "feeders.npz")


def _cases():
    with np.load(G) as z:
        d = {k: z[k] for k in z.files}
    return [(t, d[t + "_joint_loc"], d[t + "_targets"], d[t + "_mask"], [int(v) for v in d[t + "_dims"]]) for t in "abc"]


@pytest.mark.parametrize("case", _cases(), ids=lambda c: c[0])
def test_oracle_coord2map_matches_reference_golden(case):
    _, jl, targets, mask, (nx, ny, nj) = case
    t, m = feeders.coord2map(jl, nx, ny, nj)
    assert np.array_equal(m.astype(np.uint8), mask)
    assert np.array_equal(t.astype(np.float32), targets)


def test_oracle_batch_scatter_and_empty():
    t, m = feeders.coord2map(np.zeros((0, 3, 2)), 6, 7, 3)
    assert t.shape[0] == 0 and m.shape[0] == 0
    jl = np.array([[[2.0, 3.0], [np.nan, np.nan]]])
    lt, lm = feeders.batch_locref_maps(jl, [2], 4, 6, 7, 2)
    assert lt.shape == (4, 6, 7, 4) and lm[[0, 1, 3]].sum() == 0 and lm[2, :, :, :2].sum() > 0 and lm[2, :, :, 2:].sum() == 0
    # centre cell: dx = dy = 0 -> target 0 with mask 1; its right neighbour: dx = -8 px
    assert lm[2, 2, 3, 0] == 1 and lt[2, 2, 3, 0] == 0 and np.isclose(lt[2, 2, 4, 0], -8 / 7.2801)


@pytest.mark.gpu
@pytest.mark.parametrize("case", _cases(), ids=lambda c: c[0])
def test_gpu_locref_targets_bit_exact(case):
    from deepgraphpose_b200 import dataset
    from deepgraphpose_b200.engine import Engine
    _, jl, targets, mask, (nx, ny, nj) = case
    eng = Engine(nj)
    n_vis = jl.shape[0]
    nt = n_vis + 2
    pos = np.arange(n_vis) * 2 % nt if n_vis > 1 else np.array([1])
    lmap, lmask = eng.locref_targets(jl, pos, nt, nx, ny)
    lmap, lmask = lmap.cpu().numpy(), lmask.cpu().numpy()
    ref_t, ref_m = feeders.batch_locref_maps(jl, pos, nt, nx, ny, nj)
    assert np.array_equal(lmask, ref_m.astype(np.float32))
    assert np.array_equal(lmap, ref_t.astype(np.float32))
    for v, t in enumerate(pos):   # and against the reference's own output
        assert np.array_equal(lmap[t], targets[v]) and np.array_equal(lmask[t].astype(np.uint8), mask[v])

    class P:
        cfg = dict(pos_dist_thresh=17, locref_stdev=7.2801)
    t2, m2 = dataset.coord2map(P(), jl, nx, ny, nj, engine=eng)
    assert t2.dtype == np.float64 and np.array_equal(t2.astype(np.float32), targets) and np.array_equal(m2.astype(np.uint8), mask)
    eng.close()


@pytest.mark.parametrize("tag", ["m0", "m1", "m2", "m3"])
def test_synthetic_training_batch_matches_reference_gen_idx_chunk(tag):
    """The marker index vectors of synthetic.make_training_batch (what bench / tests feed) == the reference's own
    gen_idx_chunk (dataset.py:187-239) on the same labels: NaN joints of visible frames move to the hidden list,
    visible_marker_in_targets indexes targets.reshape(-1, 2)."""
    from deepgraphpose_b200 import synthetic
    with np.load(G) as z:
        g = {k[len(tag) + 1:]: z[k] for k in z.files if k.startswith(tag + "_")}
    nt, H, W, nj, seed = [int(v) for v in g["args"]]
    labels, feed = synthetic.make_training_batch(nt, H, W, nj, g["vis"].tolist(), [tuple(r) for r in g["nan"].tolist()], seed=seed)
    assert np.array_equal(np.isnan(labels), np.isnan(g["labels"])) and np.allclose(np.nan_to_num(labels), np.nan_to_num(g["labels"]))
    assert np.array_equal(feed["visible_marker_pl"], g["visible_marker"])
    assert np.array_equal(feed["hidden_marker_pl"], g["hidden_marker"])
    assert np.array_equal(feed["visible_marker_in_targets_pl"], g["vit"])
    assert feed["visible_frame_within_batch"] == g["vis"].tolist() and feed["nt_batch_pl"] == nt


def test_dlc_csv_export_matches_reference_bytes(tmp_path):
    """export_pose_like_dlc (eval.py:621-645): the csv our shim writes == the bytes the reference's own function wrote for
    the same labels (golden), and load_pose_from_dlc_to_dict reads it back."""
    from deepgraphpose_b200.eval import export_pose_like_dlc, load_pose_from_dlc_to_dict
    with np.load(G) as z:
        lab = {"x": z["csv_x"], "y": z["csv_y"], "likelihoods": z["csv_l"]}
        ref_text = str(z["csv_text"])
    save = str(tmp_path / "vid_labeled")
    export_pose_like_dlc(lab, "snapshot-step2-final--0", ["hand", "finger", "elbow"], save)
    assert open(save + ".csv").read() == ref_text
    back = load_pose_from_dlc_to_dict(save + ".csv")
    for k in lab:
        assert np.allclose(back[k], lab[k], equal_nan=True)


def test_learn_wt_matches_reference_flow():
    """learn_wt (fitdgp_util.py:454-467) == the reference's own function (same OpenCV build) on the same frames."""
    pytest.importorskip("cv2")
    from deepgraphpose_b200 import synthetic
    from deepgraphpose_b200.fitdgp_util import learn_wt
    with np.load(G) as z:
        ref = z["flow_field"]
    vid, _ = synthetic.make_video(3, 64, 96, 3, seed=77)
    got = learn_wt(vid.astype(np.float64))
    assert got.shape == ref.shape == (2, 64, 96) and float(ref.max()) > 0
    assert np.allclose(got, ref, rtol=1e-5, atol=1e-5)


def _golden():
    with np.load(G) as z:
        return {k: z[k] for k in z.files}


def test_gen_batch_matches_reference_golden():
    """fitdgp.gen_batch draws from np.random / random in the reference's order: same seeds -> the same batch list as the
    reference's own gen_batch (fitdgp_util.py:146-202, run by tests/golden/make_golden_feeders.py)."""
    import random
    from types import SimpleNamespace
    from deepgraphpose_b200 import fitdgp
    g = _golden()
    cfg = SimpleNamespace(batch_size=4, n_times_all_frames=3)
    vis = [np.array([2, 9, 17]), np.array([1])]
    hid = [np.array([5, 6, 30, 31]), np.array([], dtype=np.int64)]
    allf = [np.array([0, 1, 2, 3, 4, 7, 8, 9, 10, 11, 15, 16, 17, 18, 19]), np.array([0, 1, 2])]
    np.random.seed(11)
    random.seed(12)
    got = fitdgp.gen_batch(vis, hid, allf, cfg, 500)
    assert [len(b) for b in got] == g["gen_batch_lens"].tolist()
    assert np.array_equal(np.concatenate(got), g["gen_batch_flat"])
    assert all(b.dtype == np.int32 for b in got)


def test_hidden_frame_selection_matches_reference_golden():
    """dataset.get_neighboring_window / select_hidden_frames vs the reference's own functions (dataset.py:46-119)."""
    from deepgraphpose_b200 import dataset
    g = _golden()
    me, pv, order = g["me_values"], g["me_pv"], g["me_order"]
    assert np.array_equal(dataset.get_neighboring_window(pv, 2, len(me)), g["me_windowed"])
    for k, (nmax, jump) in enumerate(((14, None), (20, 0), (8, None), (24, 1))):
        assert np.array_equal(dataset.select_hidden_frames(2, pv, order, len(me), nmax, jump), g["me_sel%d" % k]), k


@pytest.mark.gpu
def test_gpu_motion_energy_bit_exact_vs_reference_golden():
    """dataset.calculate_motion_energy (dgp_motion_energy kernel: exact byte sums of the uint8-wrapped frame difference) ==
    the reference's calculate_motion_energy on the same clip, bit for bit, whatever the chunking; odd byte counts and
    unaligned frames take the scalar path."""
    import torch
    from deepgraphpose_b200 import dataset, synthetic
    from deepgraphpose_b200.engine import Engine
    g = _golden()
    vid, _ = synthetic.make_video(24, 48, 64, 3, seed=int(g["me_seed"]))
    eng = Engine(3, location_refinement=False)
    for chunk in (256, 5, 1):
        me = dataset.calculate_motion_energy(vid, engine=eng, chunk=chunk)
        assert me.dtype == np.float64 and np.array_equal(me, g["me_values"]), chunk
    rng = np.random.default_rng(0)
    odd = rng.integers(0, 256, (7, 37, 29, 3), dtype=np.uint8)          # 3219 bytes per frame: not a multiple of 16
    ref = np.array([0.0] + [np.mean(np.abs(odd[t] - odd[t - 1])) for t in range(1, 7)])
    assert np.array_equal(dataset.calculate_motion_energy(odd, engine=eng), ref)
    big = rng.integers(0, 256, (3, 747, 832, 3), dtype=np.uint8)
    sums = eng.motion_energy_sums(torch.from_numpy(big).cuda()).cpu().numpy()
    assert sums[0] == 0 and sums[1] == int((big[1] - big[0]).astype(np.uint64).sum()) and sums[2] == int((big[2] - big[1]).astype(np.uint64).sum())
    eng.close()


@pytest.mark.gpu
def test_gpu_marker_indices_bit_exact_vs_reference_golden():
    """dataset.gen_idx_chunk (dgp_marker_indices) == the reference's own gen_idx_chunk on the batches of the golden file
    (NaN labels in visible frames, no visible frame at all, no hidden frame at all)."""
    from deepgraphpose_b200 import dataset
    from deepgraphpose_b200.engine import Engine
    g = _golden()
    for tag in ("m0", "m1", "m2", "m3"):
        nt, H, W, nj, seed = [int(v) for v in g[tag + "_args"]]
        vis = g[tag + "_vis"]
        hid = np.array([t for t in range(nt) if t not in set(vis.tolist())], dtype=np.int64)
        eng = Engine(nj, location_refinement=False)
        v, h, vit = dataset.gen_idx_chunk(vis, hid, g[tag + "_labels"], engine=eng)
        assert np.array_equal(v, g[tag + "_visible_marker"]), tag
        assert np.array_equal(h, g[tag + "_hidden_marker"]), tag
        assert np.array_equal(vit, g[tag + "_vit"]), tag
        eng.close()
