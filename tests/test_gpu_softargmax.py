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


@pytest.mark.parametrize("shape", [(3, 30, 38, 4, 3.0), (2, 94, 104, 5, 4.0), (2, 60, 80, 20, 2.0), (1, 128, 160, 16, 6.0),
                                   (2, 20, 24, 8, 3.0), (2, 20, 24, 12, 3.0), (1, 16, 20, 32, 3.0), (1, 8, 12, 64, 3.0),
                                   (1, 8, 12, 128, 3.0), (2, 30, 38, 3, 30.0), (1, 2, 2, 1, 1.0), (2, 6, 4, 7, 2.0)])
def test_softargmax_estimate_pose_readout_only(eng, shape):
    """The path estimate_pose takes (no DLC peak wanted): speculative single pass against the fixed reference 0, every
    joint-count family of the kernel (shuffle classes 4..128, scratch reduction for 12 / 20, generic 1 / 3 / 5 / 7)."""
    B, H, W, nj, scale = shape
    rng = np.random.default_rng(B * 77 + nj)
    logits = make_logits(rng, B, H, W, nj, scale)
    lt = torch.from_numpy(logits)
    mu_ref, _ = dgp_ops.argmax_2d_from_cm(lt, nj, 1.0, 1.0)
    out = eng.softargmax(lt.cuda(), None, 1.0, 1.0, want=("mu", "peak", "lik"))
    mu = out["mu"].cpu()
    assert (mu - mu_ref).abs().max().item() < MU_TOL_PX
    for b in range(B):
        _, pk, lk = dgp_ops.estimate_pose_readout(mu[b:b + 1].numpy(), logits[b:b + 1])
        assert (pk == out["peak"][b].cpu().numpy()).all()
        assert np.allclose(lk, out["lik"][b].cpu().numpy(), atol=1e-6, equal_nan=True)
    # and the two kernel variants agree on mu
    mu_dlc = eng.softargmax(lt.cuda(), None, 1.0, 1.0, want=("mu", "dlc_peak"))["mu"].cpu()
    assert (mu - mu_dlc).abs().max().item() < MU_TOL_PX


@pytest.mark.parametrize("gauss_len", [1.0, 2.0, 3.0, 4.0])
def test_softargmax_blur_radius_border(eng, gauss_len):
    """Peaks sitting on the edges: the zero-padded blur + renormalisation bias must match for every radius."""
    rng = np.random.default_rng(int(gauss_len))
    H, W, nj = 20, 26, 4
    x = (rng.standard_normal((3, H, W, nj)) * 0.5).astype(np.float32)
    x[0, 0, 0, 0] += 9; x[0, H - 1, W - 1, 1] += 9; x[0, 0, W - 1, 2] += 9; x[0, H - 1, 0, 3] += 9
    x[1, 1, 5, 0] += 9; x[1, H - 2, 7, 1] += 9; x[1, 9, 1, 2] += 9; x[1, 11, W - 2, 3] += 9
    lt = torch.from_numpy(x)
    mu_ref, _ = dgp_ops.argmax_2d_from_cm(lt, nj, 1.0, gauss_len)
    for want in (("mu", "peak", "lik"), ("mu", "dlc_peak")):
        mu = eng.softargmax(lt.cuda(), None, 1.0, gauss_len, want=want)["mu"].cpu()
        assert (mu - mu_ref).abs().max().item() < MU_TOL_PX


def test_softargmax_out_of_range_exponents_take_the_exact_pass(eng):
    """|logit * gamma * log2(e)| > 90 leaves the speculative pass's safe range: the kernel must redo the chunk against
    the true maximum (and an all -inf joint / a NaN must come out as NaN, like the reference)."""
    rng = np.random.default_rng(3)
    logits = make_logits(rng, 2, 10, 12, 4, 200.0)   # |x| up to ~1000
    logits[1, :, :, 2] = -500.0 + rng.standard_normal((10, 12)).astype(np.float32)   # everything far below 2^-90
    lt = torch.from_numpy(logits)
    mu_ref, _ = dgp_ops.argmax_2d_from_cm(lt, 4, 1.0, 1.0)
    out = eng.softargmax(lt.cuda(), None, 1.0, 1.0, want=("mu", "peak", "lik"))
    assert torch.isfinite(out["mu"]).all()
    assert (out["mu"].cpu() - mu_ref).abs().max().item() < MU_TOL_PX
    bad = logits.copy()
    bad[0, :, :, 1] = -np.inf
    bad[1, 3, 4, 0] = np.nan
    mu = eng.softargmax(torch.from_numpy(bad).cuda(), None, 1.0, 1.0, want=("mu",))["mu"].cpu()
    assert torch.isnan(mu[0, 1]).all() and torch.isnan(mu[1, 0]).all()
    assert torch.isfinite(mu[0, 0]).all() and torch.isfinite(mu[1, 1]).all()


def test_softargmax_is_bitwise_batch_invariant(eng):
    """The chunking depends on the map shape only: a frame's result does not depend on its batch or its position in it."""
    rng = np.random.default_rng(11)
    logits = torch.from_numpy(make_logits(rng, 9, 94, 104, 4, 3.0)).cuda()
    full = eng.softargmax(logits, None, 1.0, 1.0, want=("mu", "peak", "lik"))
    for b in (0, 4, 8):
        one = eng.softargmax(logits[b:b + 1].contiguous(), None, 1.0, 1.0, want=("mu", "peak", "lik"))
        assert torch.equal(one["mu"][0], full["mu"][b]) and torch.equal(one["lik"][0], full["lik"][b])
    part = eng.softargmax(logits[3:8].contiguous(), None, 1.0, 1.0, want=("mu",))
    assert torch.equal(part["mu"], full["mu"][3:8])


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
    with pytest.raises(DgpError):
        eng.softargmax(torch.zeros(1, 6, 6, 33, device="cuda"))         # 33 joints: no lane mapping with a fixed joint
    with pytest.raises(DgpError):
        eng.softargmax(torch.zeros(1, 6, 6, 132, device="cuda"))        # > 128 joints


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


@pytest.mark.parametrize("shape", [(2, 12, 16, 3), (1, 30, 38, 5), (3, 94, 104, 4)])
def test_soft_pose_dgp_branch_matches_oracle(shape):
    """evaluate_dgp 'dgp' locref read-out (eval.py:751-785) on the device vs the literal numpy restatement, including the
    reference's quirk (locref's first component lands on the row coordinate) and the DLC-ordered variant."""
    from deepgraphpose_b200.engine import Engine
    from oracle import dgp_ops
    B, H, W, nj = shape
    rng = np.random.default_rng(B * 100 + nj)
    logits = (rng.standard_normal(shape) * 3).astype(np.float32)
    locref = rng.standard_normal((B, H, W, 2 * nj)).astype(np.float32)
    eng = Engine(nj)
    got = eng.soft_pose(torch.from_numpy(logits).cuda(), torch.from_numpy(locref).cuda()).cpu().numpy()
    got_swapped = eng.soft_pose(torch.from_numpy(logits).cuda(), torch.from_numpy(locref).cuda(), swap_offsets=True).cpu().numpy()
    for b in range(B):
        _, st = dgp_ops.argmax_2d_from_cm(torch.from_numpy(logits[b:b + 1]), nj, 1.0, 1.0)
        ref = dgp_ops.evaluate_dgp_pose_dgp_branch(st.numpy(), locref[b:b + 1])
        assert np.abs(got[b] - ref).max() < 2e-3, np.abs(got[b] - ref).max()
        lr2 = locref[b:b + 1].reshape(1, H, W, nj, 2)[..., ::-1].reshape(1, H, W, 2 * nj)
        ref2 = dgp_ops.evaluate_dgp_pose_dgp_branch(st.numpy(), np.ascontiguousarray(lr2))
        assert np.abs(got_swapped[b] - ref2).max() < 2e-3
    eng.close()
