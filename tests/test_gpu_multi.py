"""Frame-sharded inference on 2 GPUs (NCCL halo) == single-GPU inference, bit for bit (SURVEY.md 8e exactness test).
Skipped on boxes with fewer than 2 GPUs."""
import os

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, T, H, W, nj, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from deepgraphpose_b200 import sharding, synthetic
    from deepgraphpose_b200.engine import Engine
    frames, _ = synthetic.make_video(T, H, W, nj, seed=21)
    eng = Engine(nj, location_refinement=False, device=rank)
    eng.load_weights(synthetic.make_weights(nj, seed=0, location_refinement=False))
    edges = synthetic.chain_skeleton(nj)
    a, b = sharding.shard_range(T, rank, world)
    logits, _ = eng.forward(torch.from_numpy(frames[a:b]).cuda(rank))
    out = eng.softargmax(logits)
    halo = sharding.exchange_halo(out["mu"][0])
    pot = eng.potentials(out["mu"], edges, halo_next=halo)
    temporal = pot["temporal"]
    if halo is None:  # last shard has no frame after its last one
        temporal = torch.cat([temporal, torch.zeros(1, nj, device=temporal.device)])
    full = {k: sharding.gather_frames(v, T) for k, v in
            dict(mu=out["mu"], peak=out["peak"], lik=out["lik"], temporal=temporal, skel=pot["skel"].t().contiguous()).items()}
    if rank == 0:
        np.savez(os.path.join(out_dir, "sharded.npz"), **{k: v.cpu().numpy() for k, v in full.items()})
    dist.barrier()
    dist.destroy_process_group()


def test_two_gpu_shards_match_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from deepgraphpose_b200 import synthetic
    from deepgraphpose_b200.engine import Engine
    T, H, W, nj = 9, 96, 128, 4
    mp.spawn(_worker, args=(2, 29533, T, H, W, nj, str(tmp_path)), nprocs=2, join=True)
    got = np.load(os.path.join(str(tmp_path), "sharded.npz"))
    frames, _ = synthetic.make_video(T, H, W, nj, seed=21)
    eng = Engine(nj, location_refinement=False, device=0)
    eng.load_weights(synthetic.make_weights(nj, seed=0, location_refinement=False))
    logits, _ = eng.forward(torch.from_numpy(frames).cuda(0))
    out = eng.softargmax(logits)
    pot = eng.potentials(out["mu"], synthetic.chain_skeleton(nj))
    assert np.array_equal(got["mu"], out["mu"].cpu().numpy())
    assert np.array_equal(got["peak"], out["peak"].cpu().numpy())
    assert np.array_equal(got["lik"], out["lik"].cpu().numpy())
    assert np.array_equal(got["temporal"][: T - 1], pot["temporal"].cpu().numpy())
    assert np.array_equal(got["skel"], pot["skel"].t().cpu().numpy())
    eng.close()
