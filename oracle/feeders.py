"""Host feeders of the fit_dgp training step restated in numpy (oracle only).

``coord2map`` follows /root/reference/src/deepgraphpose/dataset.py:246-271, which calls
``PoseDataset.compute_target_part_scoremap``
(/root/reference/src/DeepLabCut/deeplabcut/pose_estimation_tensorflow/dataset/pose_defaultdataset.py:220-266) once per
visible frame: labels in scoremap units (row, col) -> image pixels ``*8+4`` -> (x, y); every scoremap cell whose centre
``(i*stride + stride/2, j*stride + stride/2)`` lies within ``pos_dist_thresh`` pixels of the joint gets
``locref_map[j, i, 2*joint + (0, 1)] = (dx, dy) / locref_stdev`` and ``locref_mask = 1``.  The reference's search window
(``round(max(j_x_sm - dist_thresh - 1, 0))`` ...) is +-18 cells = +-144 px and never binds for a 17 px radius, so the
restatement tests every cell.  Joints whose label is NaN are skipped (dataset.py:255-257).
"""
import numpy as np


def compute_target_part_scoremap(joint_xy, joint_id, size, num_joints, pos_dist_thresh=17.0, stride=8.0,
                                 locref_stdev=7.2801, scale=1.0):
    """pose_defaultdataset.py:220-266 for one 'person': joint_xy (k,2) image (x,y); returns (locref_map, locref_mask)."""
    height, width = int(size[0]), int(size[1])
    half = stride / 2.0
    thr2 = (pos_dist_thresh * scale) ** 2
    locref_scale = 1.0 / locref_stdev
    locref_map = np.zeros((height, width, 2 * num_joints))
    locref_mask = np.zeros((height, width, 2 * num_joints))
    pt_x = np.arange(width, dtype=np.float64) * stride + half
    pt_y = np.arange(height, dtype=np.float64) * stride + half
    for k, j_id in enumerate(joint_id):
        dx = float(joint_xy[k, 0]) - pt_x[None, :]
        dy = float(joint_xy[k, 1]) - pt_y[:, None]
        inside = dx ** 2 + dy ** 2 <= thr2
        dxb, dyb = np.broadcast_to(dx, inside.shape), np.broadcast_to(dy, inside.shape)
        locref_mask[:, :, 2 * j_id][inside] = 1
        locref_mask[:, :, 2 * j_id + 1][inside] = 1
        locref_map[:, :, 2 * j_id][inside] = dxb[inside] * locref_scale
        locref_map[:, :, 2 * j_id + 1][inside] = dyb[inside] * locref_scale
    return locref_map, locref_mask


def coord2map(joint_loc, nx_out, ny_out, nj, pos_dist_thresh=17.0, stride=8.0, locref_stdev=7.2801):
    """dataset.py:246-271.  joint_loc (n_vis, nj, 2) scoremap (row, col), NaN = missing.  Returns
    (locref_targets (n_vis,nx_out,ny_out,2nj), locref_mask) -- or two empty arrays when n_vis == 0."""
    joint_loc = np.asarray(joint_loc, dtype=np.float64)
    targets, masks = [], []
    for ii in range(joint_loc.shape[0]):
        joint_ii = np.flip(joint_loc[ii] * 8 + 4, 1)           # hard-coded *8+4 (dataset.py:252), (row,col) -> (x,y)
        keep = np.where(np.nan_to_num(joint_ii).sum(1) != 0)[0]
        keep = np.array([j for j in keep if not np.isnan(joint_ii[j]).any()], dtype=np.int64)
        t, m = compute_target_part_scoremap(joint_ii[keep], keep, (nx_out, ny_out), nj, pos_dist_thresh, stride, locref_stdev)
        targets.append(t)
        masks.append(m)
    return np.array(targets), np.array(masks)


def batch_locref_maps(joint_loc, visible_frame_within_batch, nt, nx_out, ny_out, nj, **kw):
    """fitdgp.py:781-795: scatter the visible frames' maps into zero tensors over the whole batch."""
    lt = np.zeros((nt, nx_out, ny_out, 2 * nj))
    lm = np.zeros((nt, nx_out, ny_out, 2 * nj))
    t, m = coord2map(joint_loc, nx_out, ny_out, nj, **kw)
    if len(visible_frame_within_batch) > 0 and t.shape[0] != 0:
        lt[np.asarray(visible_frame_within_batch)] = t
        lm[np.asarray(visible_frame_within_batch)] = m
    return lt, lm
