"""Wall-clock phases of eval.estimate_pose_sharded per rank (run under torchrun): where the end-to-end time of a sharded
video goes -- the streamed forward of the shard, the halo + potentials, the all-gather of the per-frame results."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lr = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
from deepgraphpose_b200.eval import estimate_pose_sharded
wl = bench.Workload("b", lr, rank, "fp16")
T = world * 10000
run = lambda n, tm=None: estimate_pose_sharded(wl.eng, wl.pool_host, n, wl.H, wl.W, wl.edges, wl.ws_vec, wl.ws_max, 0.0, batch=wl.B, timings=tm)
run(world * 2 * wl.B)
for rep in range(3):
    tm = {}
    torch.cuda.synchronize(); bench.barrier(world)
    t0 = time.perf_counter()
    run(T, tm)
    tm["total"] = time.perf_counter() - t0
    tm["stream_fps_this_rank"] = 10000 / tm["stream"]
    print(json.dumps({"rank": rank, "rep": rep, **tm}), flush=True)
    bench.barrier(world)
if world > 1:
    dist.destroy_process_group()
