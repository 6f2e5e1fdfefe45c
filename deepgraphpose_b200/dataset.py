"""Drop-ins for the host feeders of deepgraphpose.dataset that sit on the training hot loop or scan whole videos (reference:
src/deepgraphpose/dataset.py): coord2map (per step), calculate_motion_energy + select_hidden_frames (hidden-frame selection).
Video decoding, batching and the label files stay with the reference's Dataset class."""
import numpy as np
import torch


def coord2map(pdata, joint_loc, nx_out, ny_out, nj, engine=None):
    """dataset.py:246-271 with the same arguments and return values (float64 ndarrays (n_vis,nx_out,ny_out,2nj), squeezed
    the way the reference squeezes them).  ``pdata`` supplies ``cfg.pos_dist_thresh`` / ``cfg.locref_stdev`` like the
    reference's PoseDataset; the maps are computed by the CUDA feeder kernel of ``engine`` (required: there is no CPU path)."""
    if engine is None:
        raise ValueError("coord2map needs the Engine whose GPU computes the maps (deepgraphpose_b200 has no CPU fallback)")
    cfg = getattr(pdata, "cfg", pdata)
    get = (lambda k, d: cfg.get(k, d)) if isinstance(cfg, dict) else (lambda k, d: getattr(cfg, k, d))
    joint_loc = np.asarray(joint_loc, dtype=np.float64)
    n_vis = joint_loc.shape[0]
    if n_vis == 0:
        return np.array([]), np.array([])
    lmap, lmask = engine.locref_targets(joint_loc, np.arange(n_vis), n_vis, nx_out, ny_out,
                                        float(get("pos_dist_thresh", 17)), float(get("locref_stdev", 7.2801)))
    return lmap.cpu().numpy().astype(np.float64), lmask.cpu().numpy().astype(np.float64)


def gen_idx_chunk(visible_frame_indices, hidden_frame_indices, joint_loc, engine=None):
    """dataset.py:187-239 with the reference's arguments and return values (three sorted int arrays: visible_marker,
    hidden_marker, visible_marker_in_targets); computed by the marker-index kernel of ``engine`` (required).  The frame indices are
    positions within the batch (as ``next_batch`` passes them), ``joint_loc`` the (n_vis, nj, 2) labels of the visible frames in
    ascending frame order, NaN = unlabelled (such markers count as hidden)."""
    if engine is None:
        raise ValueError("gen_idx_chunk needs the Engine whose GPU computes the index vectors (no CPU fallback)")
    vis = np.asarray(visible_frame_indices, dtype=np.int64).reshape(-1)
    hid = np.asarray(hidden_frame_indices, dtype=np.int64).reshape(-1)
    nt = int(max(vis.max(initial=-1), hid.max(initial=-1)) + 1)
    vm, hm, vit = engine.marker_indices(vis, hid, joint_loc, max(nt, 1))
    return vm.cpu().numpy().astype("int"), hm.cpu().numpy().astype("int"), vit.cpu().numpy().astype("int")


def calculate_motion_energy(video, engine=None, chunk=256):
    """dataset.py:29-43: ``motion_energy[t] = np.mean(np.abs(frame[t] - frame[t-1]))`` with the frames as the decoder delivers
    them (uint8: the difference wraps modulo 256 and ``abs`` is the identity -- reproduced, it is what ranks the hidden
    frames), ``motion_energy[0] = 0``.  ``video`` is a path (decoded frame by frame with OpenCV, RGB) or an iterable / array of
    uint8 (H,W,3) frames; the byte sums are computed on the GPU of ``engine`` (dgp_motion_energy, exact integers) chunk by
    chunk, so host memory stays bounded and the result equals the reference's float64 means bit for bit."""
    if engine is None:
        raise ValueError("calculate_motion_energy needs the Engine whose GPU computes the sums (no CPU fallback)")
    if isinstance(video, (str, bytes)) or hasattr(video, "__fspath__"):
        from .eval import _iter_video
        frames = _iter_video(video)
    else:
        frames = iter(video)
    out = []
    carry = None            # last frame of the previous chunk, on the device
    buf = []
    nbytes = None

    def flush():
        nonlocal carry, buf
        if not buf:
            return
        x = torch.from_numpy(np.ascontiguousarray(np.stack(buf))).to(engine.device)
        if carry is not None:
            x = torch.cat([carry[None], x])
        sums = engine.motion_energy_sums(x).cpu().numpy()
        out.append(sums if carry is None else sums[1:])
        carry = x[-1].clone()
        buf = []

    for fr in frames:
        fr = np.asarray(fr)
        if fr.dtype != np.uint8:
            raise ValueError("calculate_motion_energy expects uint8 frames (as clip.iter_frames() yields them)")
        nbytes = fr.size
        buf.append(fr)
        if len(buf) == chunk:
            flush()
    flush()
    if not out:
        return np.zeros((0,))
    return np.concatenate(out).astype(np.float64) / float(nbytes)


def make_neighboring_window(window_size=5):
    """dataset.py:104-110: the offsets -n..n."""
    return np.arange(-int(window_size), int(window_size) + 1)


def get_neighboring_window(pv_all, ns, nt_max, nt_min=0):
    """dataset.py:113-119: sorted union of the +-ns windows around the frames ``pv_all``, clipped to [nt_min, nt_max)."""
    pv_all = np.asarray(pv_all)
    idx = np.unique(pv_all[:, None] + make_neighboring_window(ns)[None, :])
    return idx[(idx >= nt_min) & (idx < nt_max)]


def select_hidden_frames(ns, pv_all, pvh_sorted, n_frames, n_max_frames, ns_jump=None, verbose=False):
    """dataset.py:46-101: walk the frames in decreasing motion energy (``pvh_sorted``), skip those inside the +-ns window of a
    visible frame or closer than ``max(ns - ns_jump, 1)`` to an already chosen frame, and stop when the windows of all chosen
    frames would exceed ``n_max_frames``.  Returns the selected hidden-frame indices in selection order."""
    if ns_jump is None:
        ns_jump = ns
    min_gap = max(ns - ns_jump, 1)
    pv_all = np.asarray(pv_all)
    chosen = np.empty(0, dtype="int")
    visible_window = get_neighboring_window(pv_all, ns, n_frames)
    if len(visible_window) >= n_max_frames:
        return chosen
    anchors = pv_all.copy()
    pvh_sorted = np.asarray(pvh_sorted)
    skipped = 0
    for cand in pvh_sorted[~np.isin(pvh_sorted, visible_window)]:
        if len(anchors) > 0 and np.min(np.abs(cand - anchors)) < min_gap:
            skipped += 1
            continue
        if len(get_neighboring_window(np.append(anchors, cand), ns, n_frames)) > n_max_frames:
            break
        chosen = np.append(chosen, cand)
        anchors = np.append(anchors, cand)
    if verbose:
        print("Selected additional {} hidden frames".format(len(chosen)))
        print("Skipped {} high motion energy (me) frames".format(skipped))
    return chosen
