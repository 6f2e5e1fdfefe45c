"""Fused soft-argmax / peak / likelihood kernel and potentials vs the oracle (through the C ABI)."""
import numpy as np
import pytest
import torch

from oracle import dgp_ops, pose_net

pytestmark = pytest.mark.gpu

MU_TOL_PX = 1e-3   # scoremap pixels; BASELINE tolerance is 0.5 image px (= 0.0625 scoremap px)


@pytest.fixture(scope="module")
def eng():
    from deepgraphpose_b200.engine import Engine
    e = Engine(4)
    yield e
    e.close()


def make_logits(rng, B, H, W, nj, scale):
    logits = (rng.standard_normal((B, H, W, nj)) * scale).astype(np.float32)
    for b in range(B):
        for j in range(nj):
            r, c = rng.integers(0, H), rng.integers(0, W)
            if j == 0:
                r, c = 0, 0
            if j == 1:
                r, c = H - 1, W - 1
            logits[b, r, c, j] += 4 * scale
    return logits


def check(eng, logits, locref=None, gamma=1.0, gauss_len=1.0, nb=6):
    B, H, W, nj = logits.shape
    lt = torch.from_numpy(logits)
    mu_ref, _ = dgp_ops.argmax_2d_from_cm(lt, nj, gamma, gauss_len)
    out = eng.softargmax(lt.cuda(), torch.from_numpy(locref).cuda() if locref is not None else None, gamma, gauss_len)
    torch.cuda.synchronize()
    mu = out["mu"].cpu()
    ok = torch.isfinite(mu_ref)
    assert (mu[ok] - mu_ref[ok]).abs().max().item() < MU_TOL_PX
    prob = eng.sigmoid(lt.cuda()).cpu().numpy()
    assert np.abs(prob - torch.sigmoid(lt).numpy()).max() < 1e-6
    for b in range(min(B, nb)):
        # estimate_pose read-out (eval.py:331-343) from the SAME fp32 logits and the GPU's mu: bit-exact integers
        _, pk, lk = dgp_ops.estimate_pose_readout(mu[b:b + 1].numpy(), logits[b:b + 1])
        assert (pk == out["peak"][b].cpu().numpy()).all()
        assert np.allclose(lk, out["lik"][b].cpu().numpy(), atol=1e-6, equal_nan=True)
        # DLC argmax_pose_predict (predict.py:62-77) on the GPU's own fp32 scoremap: bit-exact integer peaks
        scm, loc = pose_net.extract_cnn_output(prob[b:b + 1], locref[b:b + 1] if locref is not None else None)
        pose, peaks = pose_net.argmax_pose_predict(scm, loc, 8.0)
        assert (peaks == out["dlc_peak"][b].cpu().numpy()).all()
        assert np.abs(pose - out["dlc_pose"][b].cpu().numpy()).max() < 1e-3
    return out


@pytest.mark.parametrize("shape", [(3, 30, 38, 4, 3.0), (2, 94, 104, 5, 4.0), (5, 60, 80, 20, 2.0), (1, 128, 160, 16, 6.0),
                                   (300, 30, 38, 4, 3.0), (2, 30, 38, 3, 30.0), (1, 2, 2, 1, 1.0), (2, 6, 4, 7, 2.0)])
def test_softargmax_parity(eng, shape):
    B, H, W, nj, scale = shape
    rng = np.random.default_rng(B * 1000 + H)
    logits = make_logits(rng, B, H, W, nj, scale)
    locref = rng.standard_normal((B, H, W, 2 * nj)).astype(np.float32)
    check(eng, logits, locref)


def test_softargmax_gamma_and_gauss_len(eng):
    rng = np.random.default_rng(5)
    logits = make_logits(rng, 2, 20, 24, 4, 2.0)
    check(eng, logits, None, gamma=0.5, gauss_len=1.0)
    check(eng, logits, None, gamma=2.0, gauss_len=2.0)


def test_softargmax_saturation_ties_and_delta(eng):
    H, W, nj = 12, 16, 4
    x = np.full((1, H, W, nj), -30.0, np.float32)
    x[0, 5, 7, 0] = 50.0                  # delta: soft-argmax == peak exactly
    x[0, :, :, 1] = 40.0                  # fully saturated: sigmoid == 1.0 everywhere -> first index wins
    x[0, 3, 3, 2] = 20.0
    x[0, 8, 9, 2] = 25.0                  # both saturate to 1.0f in fp32 -> tie -> first (3,3)
    x[0, 0, 0, 3] = 50.0                  # corner: biased by zero padding + renormalisation
    out = check(eng, x)
    assert torch.allclose(out["mu"][0, 0].cpu(), torch.tensor([5.0, 7.0]), atol=1e-4)
    assert out["dlc_peak"][0, 1].tolist() == [0, 0]
    assert out["dlc_peak"][0, 2].tolist() == [3, 3]
    k = np.array([0.27406862, 0.45186276, 0.27406862])
    assert abs(out["mu"][0, 3, 0].item() - k[2] / (k[1] + k[2])) < 1e-4


def test_softargmax_large_logits_no_overflow(eng):
    rng = np.random.default_rng(7)
    logits = make_logits(rng, 1, 10, 12, 2, 200.0)   # |x| up to ~1000: online softmax must not overflow
    lt = torch.from_numpy(logits)
    mu_ref, _ = dgp_ops.argmax_2d_from_cm(lt, 2, 1.0, 1.0)
    out = eng.softargmax(lt.cuda(), None, 1.0, 1.0, want=("mu", "dlc_peak"))
    assert torch.isfinite(out["mu"]).all()
    assert (out["mu"].cpu() - mu_ref).abs().max().item() < MU_TOL_PX


def test_softmax_map_matches_oracle(eng):
    rng = np.random.default_rng(9)
    logits = make_logits(rng, 2, 20, 24, 5, 2.0)
    lt = torch.from_numpy(logits)
    _, sm_ref = dgp_ops.argmax_2d_from_cm(lt, 5, 1.0, 1.0)
    sm = eng.softmax_map(lt.cuda(), 1.0, 1.0).cpu()
    assert (sm - sm_ref).abs().max().item() < 1e-6 + 1e-4 * sm_ref.max().item()
    # the reference-shaped shim returns the same pair
    from deepgraphpose_b200 import fitdgp_util
    mu, sm2 = fitdgp_util.argmax_2d_from_cm(lt.cuda(), 5, 1, 1)
    mu_ref, _ = dgp_ops.argmax_2d_from_cm(lt, 5, 1, 1)
    assert (mu.cpu() - mu_ref).abs().max().item() < MU_TOL_PX and torch.equal(sm2.cpu(), sm)


def test_softargmax_argument_errors(eng):
    from deepgraphpose_b200._lib import DgpError
    with pytest.raises(DgpError):
        eng.softargmax(torch.zeros(1, 5, 6, 2, device="cuda"))          # odd height
    with pytest.raises(ValueError):
        eng.softargmax(torch.zeros(1, 6, 6, 2))                          # not on the GPU
    assert eng.softargmax(torch.zeros(0, 6, 6, 2, device="cuda"))["mu"].shape == (0, 2, 2)   # empty batch


def test_potentials_and_halo(eng):
    from deepgraphpose_b200 import synthetic
    rng = np.random.default_rng(2)
    T, nj = 1000, 16
    mu = torch.from_numpy(rng.uniform(0, 100, (T, nj, 2)).astype(np.float32))
    for edges in (synthetic.chain_skeleton(nj), synthetic.dense_skeleton(nj)):
        S0 = dgp_ops.skeleton_matrix(edges, nj)
        d_ref = dgp_ops.skeleton_distances(mu, S0)
        t_ref = dgp_ops.temporal_distances(mu)
        ws = rng.uniform(1, 2, len(edges)).astype(np.float32)
        ws_max = rng.uniform(50, 400, len(edges)).astype(np.float32)
        out = eng.potentials(mu.cuda(), edges, ws=ws, ws_max=ws_max, wt_max=3.0)
        torch.cuda.synchronize()
        assert (out["skel"].cpu() - d_ref).abs().max().item() < 1e-3
        assert (out["temporal"].cpu() - t_ref).abs().max().item() < 1e-3
        e_ref = (torch.from_numpy(ws)[:, None] * (torch.relu(d_ref - torch.from_numpy(ws_max)[:, None]) + torch.from_numpy(ws_max)[:, None])).sum(0)
        assert (out["e_skel"].cpu() - e_ref).abs().max().item() < 1e-3 * e_ref.max().item()
        # contiguous shards + one-frame halo == unsharded, bit for bit
        a = eng.potentials(mu[:400].cuda(), edges, halo_next=mu[400].cuda())
        b = eng.potentials(mu[400:].cuda(), edges)
        assert torch.equal(torch.cat([a["temporal"], b["temporal"]]), out["temporal"])
        assert torch.equal(torch.cat([a["skel"], b["skel"]], dim=1), out["skel"])
