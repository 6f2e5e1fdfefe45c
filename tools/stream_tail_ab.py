"""A/B of the short-batch rule of the streaming entry (csrc/stream.cu): a shard whose start is not a multiple of the pool
wraps the cyclic source once with a short batch.  Prints frames/s for an aligned and a misaligned 10 000-frame shard.
DGP_STREAM_NO_PAD=1 restores the old behaviour (a second plan for the short batch)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench

wl = bench.Workload("b", 0, 0, "fp16")
run = lambda start, n: wl.eng.estimate_pose_stream(wl.pool_host, wl.H, wl.W, n, wl.B, 1.0, 1.0, start=start)
run(0, 4 * wl.B)
out = {}
for name, start in (("aligned", 0), ("misaligned", 10000), ("aligned_again", 0), ("misaligned_again", 10000)):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mu, peak, lik = run(start, 10000)
    torch.cuda.synchronize()
    out[name] = 10000 / (time.perf_counter() - t0)
a = run(10000, 200)
os.environ["DGP_STREAM_NO_PAD"] = "1"
b = run(10000, 200)
out["padded_equals_unpadded"] = all(bool(torch.equal(x, y)) for x, y in zip(a, b))
out["no_pad_env"] = bool(os.environ.get("DGP_STREAM_NO_PAD_AT_START"))
print(json.dumps(out))
