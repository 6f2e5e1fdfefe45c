"""NCCL all-reduce of a gradient-sized fp32 buffer (94 MB), alone: 1 bucket vs dp.BUCKETS buckets.  torchrun, one rank per GPU."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepgraphpose_b200 import dp  # noqa: E402

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
buf = torch.randn(23_600_000, device="cuda")
for buckets in (1, dp.BUCKETS):
    for _ in range(5):
        dp.allreduce_flat_(buf, buckets=buckets)
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        dp.allreduce_flat_(buf, buckets=buckets)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    w = dist.get_world_size()
    if rank == 0:
        print("allreduce %d MB, world %d, buckets %d: %.3f ms (bus bw %.0f GB/s)" % (buf.numel() * 4 // 2 ** 20, w, buckets, ms,
                                                                                  2 * (w - 1) / w * buf.numel() * 4 / ms / 1e6), flush=True)
dist.barrier()
dist.destroy_process_group()
