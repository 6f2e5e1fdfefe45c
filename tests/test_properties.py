"""Property tests (hypothesis) of the host-side contracts of the path: shape algebra, marker index vectors, the locref
feeder restatement and the skeleton-clique host precompute.  CPU only."""
import ctypes as C

import numpy as np
from hypothesis import assume, given, settings
from hypothesis import strategies as st

from deepgraphpose_b200 import _lib, fitdgp, sharding, synthetic
from oracle import dgp_loss as oracle_loss
from oracle import dgp_ops, feeders, resnet_v1, tf_ops


@settings(max_examples=200, deadline=None)
@given(st.integers(32, 2000), st.integers(32, 2000))
def test_output_dims_closed_form_equals_the_layer_chain(lib_built, H, W):
    """dgp_output_dims (C ABI, replaces Dataset._compute_pred_dims) == oracle closed form == the SAME-padding chain of
    conv1 (s2) -> pool (s2) -> block1 (s2) -> block2 (s2) -> deconv (x2)."""
    lib = _lib.load()
    a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    assert lib.dgp_output_dims(H, W, C.byref(a), C.byref(b), C.byref(c), C.byref(d)) == 0
    (fh, fw), (oh, ow) = resnet_v1.output_dims(H, W)
    assert (a.value, b.value, c.value, d.value) == (fh, fw, oh, ow)
    h, w = H, W
    for k, s in ((7, 2), (3, 2), (3, 2), (3, 2)):
        h, w = tf_ops.same_pad(h, k, s)[2], tf_ops.same_pad(w, k, s)[2]
    assert (h, w) == (fh, fw) and (oh, ow) == (2 * fh, 2 * fw)


@settings(max_examples=100, deadline=None)
@given(st.integers(0, 5000), st.integers(1, 16))
def test_shard_ranges_partition_the_video(T, world):
    r = [sharding.shard_range(T, k, world) for k in range(world)]
    assert r[0][0] == 0 and r[-1][1] == T and all(a[1] == b[0] for a, b in zip(r[:-1], r[1:]))
    sizes = [b - a for a, b in r]
    assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 12), st.integers(1, 8), st.data())
def test_training_batch_marker_vectors(nt, nj, data):
    vis = sorted(data.draw(st.sets(st.integers(0, nt - 1), max_size=nt)))
    nan = sorted(data.draw(st.sets(st.tuples(st.integers(0, max(len(vis) - 1, 0)), st.integers(0, nj - 1)), max_size=4))) if vis else []
    labels, feed = synthetic.make_training_batch(nt, 20, 24, nj, vis, nan, seed=1)
    v, h, vit = feed["visible_marker_pl"], feed["hidden_marker_pl"], feed["visible_marker_in_targets_pl"]
    # every marker of the batch is either visible or hidden, never both
    assert sorted(np.concatenate([v, h]).tolist()) == list(range(nt * nj))
    # visible markers belong to visible frames and carry a finite label; vit indexes targets.reshape(-1, 2)
    flat = labels.reshape(-1, 2)
    assert len(vit) == len(v) and np.isfinite(flat[vit]).all()
    for m, row in zip(v, vit):
        t, j = divmod(int(m), nj)
        assert t in vis and row == vis.index(t) * nj + j
    # NaN-labelled joints of visible frames are hidden markers (dataset.py:206-220)
    for (i, j) in nan:
        assert vis[i] * nj + j in h.tolist()


@settings(max_examples=40, deadline=None)
@given(st.integers(6, 40), st.integers(6, 40), st.floats(-3.0, 45.0), st.floats(-3.0, 45.0))
def test_locref_feeder_geometry(nx, ny, r, c):
    """coord2map: exactly the cells whose centre lies within 17 px get mask 1 and the target (dx, dy) / locref_stdev.
    A joint whose image coordinates sum to exactly 0 is dropped by the reference's `nan_to_num(...).sum != 0` filter
    (dataset.py:255) -- pinned separately in test_locref_feeder_drops_joint_whose_coordinates_sum_to_zero."""
    assume((c * 8 + 4) + (r * 8 + 4) != 0)
    t, m = feeders.coord2map(np.array([[[r, c]]]), nx, ny, 1)
    jj, ii = np.meshgrid(np.arange(nx), np.arange(ny), indexing="ij")
    dx, dy = (c * 8 + 4) - (ii * 8.0 + 4), (r * 8 + 4) - (jj * 8.0 + 4)
    inside = dx ** 2 + dy ** 2 <= 17.0 ** 2
    assert np.array_equal(m[0, :, :, 0] == 1, inside) and np.array_equal(m[0, :, :, 1] == 1, inside)
    assert np.allclose(t[0, :, :, 0][inside], dx[inside] / 7.2801) and np.allclose(t[0, :, :, 1][inside], dy[inside] / 7.2801)
    assert (t[0][~inside] == 0).all() and np.abs(t).max() <= 17.0 / 7.2801 + 1e-12


def test_locref_feeder_drops_joint_whose_coordinates_sum_to_zero():
    """Reference quirk (dataset.py:255): joints are kept iff nan_to_num(x, y).sum() != 0, so a finite label with
    x + y == 0 (e.g. scoremap (-0.5, -0.5) -> image (0, 0)) is treated like a missing one: no mask, no target."""
    for r, c in ((-0.5, -0.5), (-1.0, 0.0), (0.25, -1.25)):
        t, m = feeders.coord2map(np.array([[[r, c], [1.0, 1.0]]]), 8, 8, 2)
        assert not m[0, :, :, :2].any() and not t[0, :, :, :2].any()
        assert m[0, :, :, 2:].any()


@settings(max_examples=40, deadline=None)
@given(st.integers(1, 6), st.integers(2, 6), st.integers(0, 2 ** 31 - 1))
def test_spatial_clique_params_shim_equals_oracle(n_vis, nj, seed):
    """Host precompute of dgp_loss (fitdgp.py:874-892): the shim's numpy restatement == the oracle's, incl. NaN labels and
    the missing-limb quirk (a missing limb contributes stride/2 to the mean)."""
    rng = np.random.default_rng(seed)
    labels = np.stack([rng.uniform(0, 90, (n_vis, nj)), rng.uniform(0, 100, (n_vis, nj))], axis=2)
    labels[rng.uniform(size=(n_vis, nj)) < 0.2] = np.nan
    edges = synthetic.chain_skeleton(nj)
    S0 = dgp_ops.skeleton_matrix(edges, nj)
    cfg = oracle_loss.default_dgp_cfg()
    ws, ws_max = oracle_loss.spatial_clique_params(labels, S0, cfg)
    ws2, ws_max2 = fitdgp.spatial_clique_params([labels], S0, 8.0, 1000.0, 1.2)
    assert np.allclose(ws, ws2, rtol=1e-6, equal_nan=True) and np.allclose(ws_max, ws_max2, rtol=1e-6, equal_nan=True)
    assert fitdgp.skeleton_edges(S0) == [tuple(e) for e in edges]
