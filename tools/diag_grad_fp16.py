"""Gradient parity of the training step in fp16 storage mode (8x less rounding noise than bf16) vs the fp32 oracle."""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import test_gpu_train as T  # noqa: E402
from deepgraphpose_b200 import fitdgp  # noqa: E402
from deepgraphpose_b200.engine import Engine  # noqa: E402

W, frames, batch, edges, S0, cfg, ws, ws_max = T._setup()
loss, ref, heads = T._oracle_grads(W, frames, batch, S0, cfg, ws, ws_max)
for prec in ("bf16", "fp16"):
    eng = Engine(T.NJ, precision=prec)
    eng.load_weights(W)
    got = fitdgp.train_forward_backward(eng, torch.from_numpy(frames).cuda(), batch, cfg, edges, ws, ws_max, 200, 20)
    rel, cos = [], []
    for name, g_ref in sorted(ref.items()):
        g = eng.get_variable(name, "grad")
        nr = np.linalg.norm(g_ref)
        rel.append(np.linalg.norm(g - g_ref) / nr)
        cos.append((g * g_ref).sum() / (np.linalg.norm(g) * nr + 1e-30))
    i = int(np.argmax(rel))
    print("%s: loss %.6f (oracle %.6f) | grad rel-L2 median %.4f max %.4f (%s) | min cos %.5f" % (
        prec, float(got["total_loss"]), loss["total_loss"], np.median(rel), max(rel), sorted(ref)[i], min(cos)), flush=True)
    eng.close()
