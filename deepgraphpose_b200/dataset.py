"""Drop-ins for the host feeders of deepgraphpose.dataset that sit on the training hot loop (reference:
src/deepgraphpose/dataset.py).  Only the per-step feeders are here; video decoding, batching and the label files stay with
the reference's Dataset class."""
import numpy as np


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
