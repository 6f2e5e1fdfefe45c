"""Drop-in for the hot-path helpers of deepgraphpose.models.fitdgp_util
(reference: src/deepgraphpose/models/fitdgp_util.py).

``argmax_2d_from_cm`` keeps the reference signature (:342) and return pair; it takes the (N,H,W,nj) float32 logit
tensor as a CUDA tensor (or a numpy array, copied to the current device) instead of a tf.Tensor.
"""
import numpy as np
import torch

from .engine import Engine

_ENGINES = {}


def _engine_for(nj, device):
    key = (int(nj), torch.device(device).index)
    if key not in _ENGINES:
        _ENGINES[key] = Engine(nj, location_refinement=False, device=key[1])
    return _ENGINES[key]


def argmax_2d_from_cm(tensor, nj, gamma=1, gauss_len=2, th=None):
    """fitdgp_util.py:342-402: spatial softmax -> Gaussian blur -> renormalise -> soft-argmax.

    Returns (spatial_soft_argmax (N,nj,2) [(row, col) in scoremap pixels], softmax_tensor0 (N,H,W,nj)).
    ``th`` (thresholding, unused by every reference caller) is not on the B200 path.
    """
    if th is not None:
        raise NotImplementedError("argmax_2d_from_cm(th=...) is unused by the reference callers and not implemented")
    as_numpy = not isinstance(tensor, torch.Tensor)
    t = torch.from_numpy(np.ascontiguousarray(tensor, dtype=np.float32)).cuda() if as_numpy else tensor
    if t.dim() != 4 or t.shape[-1] != nj:
        raise ValueError("tensor must be (N, H, W, nj)")
    eng = _engine_for(nj, t.device)
    mu = eng.softargmax(t, None, gamma, gauss_len, want=("mu",))["mu"]
    sm = eng.softmax_map(t, gamma, gauss_len)
    if as_numpy:
        return mu.cpu().numpy(), sm.cpu().numpy()
    return mu, sm


def make_2Dgrids(H, W, device="cuda"):
    """fitdgp_util.py:318-339: (H, W, 1, 2) grid of (row, col)."""
    r = torch.arange(H, dtype=torch.float32, device=device).view(H, 1).expand(H, W)
    c = torch.arange(W, dtype=torch.float32, device=device).view(1, W).expand(H, W)
    return torch.stack([r, c], dim=2).unsqueeze(2)


def learn_wt(all_data_batch):
    """fitdgp_util.py:454-467: optical-flow magnitude per consecutive frame pair, the ``vector_field_tf`` feed of the
    temporal clique (nt-1, H, W): OpenCV Farneback flow (pyr_scale 0.5, 3 levels, window 15, 3 iterations, poly_n 5,
    poly_sigma 1.2) on the BGR2GRAY-converted frames, |u| + |v|.  Host-side (cv2), exactly the reference's feeder; moving it
    to the GPU is SURVEY.md 8(f) rank 4."""
    import cv2
    frames = np.asarray(all_data_batch)
    gray = [cv2.cvtColor(f.astype(np.uint8), cv2.COLOR_BGR2GRAY) for f in frames]
    fields = [np.abs(cv2.calcOpticalFlowFarneback(a, b, None, 0.5, 3, 15, 3, 5, 1.2, 0)).sum(2) for a, b in zip(gray[:-1], gray[1:])]
    return np.array(fields)
