"""One soft-argmax call on a configs[1]-shaped scoremap batch (for `ncu -k regex:softargmax`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from deepgraphpose_b200.engine import Engine  # noqa: E402

B, H, W, nj = (int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (4096, 94, 104, 4)))
want = ("mu", "peak", "lik") if len(sys.argv) > 5 and sys.argv[5] == "nodlc" else ("mu", "peak", "lik", "dlc_peak", "dlc_pose")
eng = Engine(nj)
x = torch.randn(B, H, W, nj, device="cuda") * 3
for _ in range(3):
    r = eng.softargmax(x, None, 1.0, 1.0, want=want)
torch.cuda.synchronize()
