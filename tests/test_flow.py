"""learn_wt (fitdgp_util.py:454-467): OpenCV's dense Farneback flow restated.  OpenCV is an un-vendored dependency of the reference;
the CPU restatement below (numpy, slow) is pinned against cv2.calcOpticalFlowFarneback itself and against the golden field the
reference's own learn_wt produced (tests/golden/feeders.npz); the CUDA path (dgp_learn_wt) is checked against both."""
import os

import numpy as np
import pytest

from deepgraphpose_b200 import synthetic

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "feeders.npz")
BORDER = np.array([0.14, 0.14, 0.4472, 0.4472, 0.4472], np.float32)


def _poly_consts(n=5, sigma=1.2):
    x = np.arange(-n, n + 1)
    g = np.exp(-x * x / (2 * sigma * sigma)).astype(np.float32)
    g = (g * (1.0 / np.sum(g.astype(np.float64)))).astype(np.float32)
    xg, xxg = (x * g).astype(np.float32), (x * x * g).astype(np.float32)
    Gm = np.zeros((6, 6))
    for yy in range(-n, n + 1):
        for xx in range(-n, n + 1):
            w = float(g[yy + n]) * float(g[xx + n])
            Gm[0, 0] += w; Gm[1, 1] += w * xx * xx; Gm[3, 3] += w * xx ** 4; Gm[5, 5] += w * xx * xx * yy * yy
    Gm[2, 2] = Gm[0, 3] = Gm[0, 4] = Gm[3, 0] = Gm[4, 0] = Gm[1, 1]
    Gm[4, 4] = Gm[3, 3]
    Gm[3, 4] = Gm[4, 3] = Gm[5, 5]
    inv = np.linalg.inv(Gm)
    return g, xg, xxg, inv[1, 1], inv[0, 3], inv[3, 3], inv[5, 5]


def _poly_exp(src, n=5):
    g, xg, xxg, ig11, ig03, ig33, ig55 = _poly_consts(n)
    H, W = src.shape
    ys, xs = np.arange(H), np.arange(W)
    r0, r1, r2 = src * g[n], np.zeros_like(src), np.zeros_like(src)
    for k in range(1, n + 1):
        a, b = src[np.maximum(ys - k, 0)], src[np.minimum(ys + k, H - 1)]
        r0, r1, r2 = r0 + g[n + k] * (a + b), r1 + xg[n + k] * (b - a), r2 + xxg[n + k] * (a + b)
    sh = lambda r, k: r[:, np.clip(xs + k, 0, W - 1)].astype(np.float64)
    b1, b2, b3, b4, b5, b6 = r0.astype(np.float64) * g[n], 0, r1.astype(np.float64) * g[n], 0, r2.astype(np.float64) * g[n], 0
    for k in range(1, n + 1):
        tg = sh(r0, k) + sh(r0, -k)
        b1, b4 = b1 + tg * g[n + k], b4 + tg * xxg[n + k]
        b2 = b2 + (sh(r0, k) - sh(r0, -k)) * xg[n + k]
        b3 = b3 + (sh(r1, k) + sh(r1, -k)) * g[n + k]
        b6 = b6 + (sh(r1, k) - sh(r1, -k)) * xg[n + k]
        b5 = b5 + (sh(r2, k) + sh(r2, -k)) * g[n + k]
    return np.stack([b3 * ig11, b2 * ig11, b1 * ig03 + b5 * ig33, b1 * ig03 + b4 * ig33, b6 * ig55], -1).astype(np.float32)


def _update_matrices(R0, R1, flow):
    H, W, _ = R0.shape
    ys, xs = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    dx, dy = flow[..., 0], flow[..., 1]
    fx, fy = (xs + dx).astype(np.float32), (ys + dy).astype(np.float32)
    x1, y1 = np.floor(fx).astype(np.int64), np.floor(fy).astype(np.int64)
    fx, fy = fx - x1, fy - y1
    inside = (x1 >= 0) & (x1 < W - 1) & (y1 >= 0) & (y1 < H - 1)
    xc, yc = np.clip(x1, 0, W - 2), np.clip(y1, 0, H - 2)
    a00, a01, a10, a11 = (1 - fx) * (1 - fy), fx * (1 - fy), (1 - fx) * fy, fx * fy
    Rw = (a00[..., None] * R1[yc, xc] + a01[..., None] * R1[yc, xc + 1] + a10[..., None] * R1[yc + 1, xc]
          + a11[..., None] * R1[yc + 1, xc + 1]).astype(np.float32)
    r2, r3 = np.where(inside, Rw[..., 0], 0), np.where(inside, Rw[..., 1], 0)
    r4 = np.where(inside, (R0[..., 2] + Rw[..., 2]) * 0.5, R0[..., 2])
    r5 = np.where(inside, (R0[..., 3] + Rw[..., 3]) * 0.5, R0[..., 3])
    r6 = np.where(inside, (R0[..., 4] + Rw[..., 4]) * 0.25, R0[..., 4] * 0.5)
    r2, r3 = (R0[..., 0] - r2) * 0.5, (R0[..., 1] - r3) * 0.5
    r2, r3 = r2 + r4 * dy + r6 * dx, r3 + r6 * dy + r5 * dx
    sx, sy = np.ones(W, np.float32), np.ones(H, np.float32)
    for i in range(5):
        sx[i] *= BORDER[i]; sx[W - 1 - i] *= BORDER[i]; sy[i] *= BORDER[i]; sy[H - 1 - i] *= BORDER[i]
    sc = sy[:, None] * sx[None, :]
    r2, r3, r4, r5, r6 = [(v * sc).astype(np.float32) for v in (r2, r3, r4, r5, r6)]
    return np.stack([r4 * r4 + r6 * r6, (r4 + r5) * r6, r5 * r5 + r6 * r6, r4 * r2 + r6 * r3, r6 * r2 + r5 * r3], -1).astype(np.float32)


def _update_flow(M, block=15):
    m = block // 2
    H, W, _ = M.shape
    ys, xs = np.arange(H), np.arange(W)
    v = np.zeros(M.shape, np.float64)
    for k in range(-m, m + 1):
        v += M[np.clip(ys + k, 0, H - 1)]
    b = np.zeros_like(v)
    for k in range(-m, m + 1):
        b += v[:, np.clip(xs + k, 0, W - 1)]
    b *= 1.0 / (block * block)
    g11, g12, g22, h1, h2 = [b[..., i] for i in range(5)]
    idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3)
    return np.stack([(g11 * h2 - g12 * h1) * idet, (g22 * h1 - g12 * h2) * idet], -1).astype(np.float32)


def _gauss(f, sigma):
    ksize = max(int(round(sigma * 5)) | 1, 3)
    if sigma <= 0:
        k = np.array([0.25, 0.5, 0.25], np.float32)
    else:
        x = np.arange(ksize) - (ksize - 1) * 0.5
        k = np.exp(-0.5 / (sigma * sigma) * x * x)
        k = (k / k.sum()).astype(np.float32)
    r = ksize // 2
    refl = lambda i, n: np.where(np.abs(i) >= n, 2 * (n - 1) - np.abs(i), np.abs(i))
    H, W = f.shape
    t = sum(k[j] * f[:, refl(np.arange(W) + j - r, W)] for j in range(ksize))
    return sum(k[j] * t[refl(np.arange(H) + j - r, H)] for j in range(ksize)).astype(np.float32)


def _resize(a, w, h):
    sq = a.ndim == 2
    a = a[..., None] if sq else a

    def coeffs(dst, src):
        f = ((np.arange(dst) + 0.5) * (src / dst) - 0.5).astype(np.float32)
        s = np.floor(f).astype(np.int64)
        f = (f - s).astype(np.float32)
        f, s = np.where(s < 0, 0, f), np.where(s < 0, 0, s)
        f, s = np.where(s >= src - 1, 0, f), np.where(s >= src - 1, src - 1, s)
        return s, np.minimum(s + 1, src - 1), f.astype(np.float32)
    x0, x1, fx = coeffs(w, a.shape[1])
    y0, y1, fy = coeffs(h, a.shape[0])
    t = (a[:, x0] * (1 - fx)[None, :, None] + a[:, x1] * fx[None, :, None]).astype(np.float32)
    o = (t[y0] * (1 - fy)[:, None, None] + t[y1] * fy[:, None, None]).astype(np.float32)
    return o[..., 0] if sq else o


def gray_bgr2gray(frame):
    c = frame.astype(np.int64)
    return ((c[..., 0] * 3735 + c[..., 1] * 19235 + c[..., 2] * 9798 + 16384) >> 15).astype(np.uint8)


def farneback_numpy(prev, nxt, levels=3, iters=3):
    """cv2.calcOpticalFlowFarneback(prev, nxt, None, 0.5, 3, 15, 3, 5, 1.2, 0), restated (see csrc/flow_kernels.cu)."""
    H0, W0 = prev.shape
    k, scale = 0, 1.0
    while k < levels:
        scale *= 0.5
        if W0 * scale < 32 or H0 * scale < 32:
            break
        k += 1
    flow = None
    for k in range(k, -1, -1):
        scale = 0.5 ** k
        sigma = (1.0 / scale - 1) * 0.5
        w, h = int(round(W0 * scale)), int(round(H0 * scale))
        flow = np.zeros((h, w, 2), np.float32) if flow is None else _resize(flow, w, h) * np.float32(2.0)
        R = [_poly_exp(_resize(_gauss(img.astype(np.float32), sigma), w, h)) for img in (prev, nxt)]
        M = _update_matrices(R[0], R[1], flow)
        for i in range(iters):
            flow = _update_flow(M)
            if i < iters - 1:
                M = _update_matrices(R[0], R[1], flow)
    return flow


def test_numpy_restatement_matches_cv2_and_reference_golden():
    cv2 = pytest.importorskip("cv2")
    with np.load(G) as z:
        seed, golden = int(z["flow_seed"]), z["flow_field"]
    vid, _ = synthetic.make_video(3, 64, 96, 3, seed=seed)
    gray = [cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in vid]
    assert all(np.array_equal(gray_bgr2gray(f), g) for f, g in zip(vid, gray))
    for a, b, gold in zip(gray[:-1], gray[1:], golden):
        ref = cv2.calcOpticalFlowFarneback(a, b, None, 0.5, 3, 15, 3, 5, 1.2, 0)
        mine = farneback_numpy(a, b)
        assert np.abs(mine - ref).max() < 1e-4
        assert np.abs(np.abs(mine).sum(2) - gold).max() < 2e-4
    # an odd size with all four pyramid levels (53 x 37 would drop levels: min size 32)
    vid, _ = synthetic.make_video(2, 261, 333, 3, seed=5)
    a, b = [cv2.cvtColor(f, cv2.COLOR_BGR2GRAY) for f in vid]
    assert np.abs(farneback_numpy(a, b) - cv2.calcOpticalFlowFarneback(a, b, None, 0.5, 3, 15, 3, 5, 1.2, 0)).max() < 1e-4


@pytest.mark.gpu
def test_gpu_learn_wt_matches_cv2_and_golden():
    """dgp_learn_wt (all frame pairs of a batch per launch) vs the reference's learn_wt (cv2) and its golden field."""
    cv2 = pytest.importorskip("cv2")
    import torch
    from deepgraphpose_b200 import fitdgp_util
    from deepgraphpose_b200.engine import Engine
    eng = Engine(3, location_refinement=False)
    with np.load(G) as z:
        seed, golden = int(z["flow_seed"]), z["flow_field"]
    vid, _ = synthetic.make_video(3, 64, 96, 3, seed=seed)
    got = fitdgp_util.learn_wt(vid.astype(np.float64), engine=eng)
    assert got.is_cuda and got.dtype == torch.float32 and tuple(got.shape) == (2, 64, 96)
    assert np.abs(got.cpu().numpy() - golden).max() < 5e-4
    for (T, H, W, seed) in ((4, 235, 301, 1), (2, 747, 832, 2), (3, 40, 50, 3)):
        vid, _ = synthetic.make_video(T, H, W, 3, seed=seed)
        ref = fitdgp_util.learn_wt(vid)                       # the reference's cv2 loop
        got = fitdgp_util.learn_wt(vid, engine=eng).cpu().numpy()
        assert got.shape == ref.shape
        assert np.abs(got - ref).max() < 1e-3 + 1e-3 * np.abs(ref).max(), (T, H, W, np.abs(got - ref).max())
    assert tuple(eng.learn_wt(torch.zeros((1, 32, 32, 3), dtype=torch.uint8, device="cuda")).shape) == (0, 32, 32)
    eng.close()
