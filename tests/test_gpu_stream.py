"""Streaming estimate_pose (dgp_estimate_pose_stream): frames pulled from a source through a pinned
ring, bit-identical to the in-memory path, bounded host memory, errors of the source surfaced."""
import os

import numpy as np
import pytest
import torch

from deepgraphpose_b200 import synthetic

pytestmark = pytest.mark.gpu
NJ, H, W = 4, 96, 128


@pytest.fixture(scope="module")
def eng():
    from deepgraphpose_b200.engine import Engine
    e = Engine(NJ, location_refinement=False)
    e.load_weights(synthetic.make_weights(NJ, seed=0, location_refinement=False))
    yield e
    e.close()


def test_stream_from_iterator_equals_in_memory_path(eng):
    from deepgraphpose_b200 import eval as dgp_eval
    frames, _ = synthetic.make_video(23, H, W, NJ, seed=4)
    ref = dgp_eval.estimate_pose_frames(eng, frames, batch=5)
    got = dgp_eval.estimate_pose_stream(eng, iter(frames), H, W, n_frames=None, batch=5)
    for k in ("x", "y", "likelihoods", "mu_likelihoods", "markers"):
        assert np.array_equal(got[k], ref[k]), k
    # a declared length shorter than the source stops early; a longer one returns what the source had
    short = dgp_eval.estimate_pose_stream(eng, iter(frames), H, W, n_frames=7, batch=5)
    assert short["x"].shape == (7, NJ) and np.array_equal(short["x"], ref["x"][:7])
    longer = dgp_eval.estimate_pose_stream(eng, iter(frames), H, W, n_frames=100, batch=5)
    assert longer["x"].shape == (23, NJ)


def test_stream_cyclic_pinned_pool_and_shard_offsets(eng):
    """A pinned pool cycled as the video (frame t = pool[t % P]) is served zero-copy; `start` shifts the range a rank reads."""
    from deepgraphpose_b200 import eval as dgp_eval
    frames, _ = synthetic.make_video(6, H, W, NJ, seed=5)
    pool = torch.from_numpy(frames).pin_memory()
    video = np.stack([frames[t % 6] for t in range(20)])
    ref = dgp_eval.estimate_pose_frames(eng, video, batch=4)
    mu, peak, lik = eng.estimate_pose_stream(pool, H, W, 20, batch=4)
    assert np.array_equal(mu.numpy().astype(np.float64), ref["markers"]) and np.array_equal(peak.numpy(), ref["mu_likelihoods"])
    mu2, _, lik2 = eng.estimate_pose_stream(pool, H, W, 9, batch=4, start=7)
    assert np.array_equal(mu2.numpy().astype(np.float64), ref["markers"][7:16])
    assert np.array_equal(lik2.numpy().astype(np.float64), ref["likelihoods"][7:16])


def test_sharded_entry_on_one_rank_equals_forward_plus_potentials(eng):
    from deepgraphpose_b200 import eval as dgp_eval
    frames, _ = synthetic.make_video(11, H, W, NJ, seed=6)
    pool = torch.from_numpy(frames).pin_memory()
    edges = synthetic.chain_skeleton(NJ)
    ws, ws_max = np.full(3, 12.0, np.float32), np.full(3, 90.0, np.float32)
    res = dgp_eval.estimate_pose_sharded(eng, pool, 11, H, W, edges, ws, ws_max, 0.0, batch=4)
    logits, _ = eng.forward(torch.from_numpy(frames).cuda())
    # the same read-outs the streaming entry requests: asking for the DLC global peak as well selects the soft-argmax variant
    # with an exact max pass, whose sums differ from the fixed-reference pass in the last ulp
    out = eng.softargmax(logits, want=("mu", "peak", "lik"))
    pot = eng.potentials(out["mu"], edges, ws=ws, ws_max=ws_max)
    assert res["shard"] == (0, 11)
    assert np.array_equal(res["markers"], out["mu"].cpu().numpy().astype(np.float64))
    assert np.array_equal(res["temporal"][:-1], pot["temporal"].cpu().numpy()) and (res["temporal"][-1] == 0).all()
    assert np.array_equal(res["skel"], pot["skel"].t().cpu().numpy())
    assert np.array_equal(res["e_skel"], pot["e_skel"].cpu().numpy())
    # the same through a callable source (how a rank opens its own range of a real video)
    res2 = dgp_eval.estimate_pose_sharded(eng, lambda a, b: iter(frames[a:b]), 11, H, W, edges, ws, ws_max, 0.0, batch=4)
    assert np.array_equal(res2["markers"], res["markers"]) and np.array_equal(res2["skel"], res["skel"])


def test_stream_keeps_host_memory_bounded(eng):
    """10 000 frames from a generator: resident memory must not grow with the length of the video (the in-memory path would
    hold 10 000 * 36 KB = 370 MB here; a 10 k-frame 747x832 clip 18.6 GB)."""
    import psutil
    from deepgraphpose_b200 import eval as dgp_eval
    base, _ = synthetic.make_video(8, H, W, NJ, seed=7)
    n = 10000

    def gen():
        for t in range(n):
            yield base[t % 8]

    dgp_eval.estimate_pose_stream(eng, gen(), H, W, n_frames=64, batch=32)       # warm-up: ring, plans
    proc = psutil.Process(os.getpid())
    rss0 = proc.memory_info().rss
    res = dgp_eval.estimate_pose_stream(eng, gen(), H, W, n_frames=None, batch=32)
    grown = proc.memory_info().rss - rss0
    assert res["x"].shape == (n, NJ) and np.isfinite(res["x"]).all()
    assert np.array_equal(res["x"][:8], res["x"][8:16])                           # the clip repeats every 8 frames
    assert grown < 64 * 2 ** 20, "resident memory grew by %.1f MB" % (grown / 2 ** 20)


def test_stream_surfaces_reader_errors(eng):
    from deepgraphpose_b200 import eval as dgp_eval
    frames, _ = synthetic.make_video(6, H, W, NJ, seed=8)

    def bad_shape():
        yield frames[0]
        yield frames[1][:, :-1]

    with pytest.raises(ValueError):
        dgp_eval.estimate_pose_stream(eng, bad_shape(), H, W, n_frames=None, batch=4)

    def raises():
        yield frames[0]
        raise IOError("decoder died")

    with pytest.raises(IOError):
        dgp_eval.estimate_pose_stream(eng, raises(), H, W, n_frames=None, batch=4)
    # the engine is still usable afterwards
    ok = dgp_eval.estimate_pose_stream(eng, iter(frames), H, W, n_frames=None, batch=4)
    assert ok["x"].shape == (6, NJ)
