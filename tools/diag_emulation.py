"""Where does the GPU forward leave the bf16-emulating oracle?  Per end_point: max-abs / rms difference relative to max."""
import sys
from unittest import mock

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
import test_gpu_train as T  # noqa: E402
from deepgraphpose_b200.engine import Engine  # noqa: E402
from oracle import pose_net, resnet_v1  # noqa: E402

W, frames, batch, edges, S0, cfg, ws, ws_max = T._setup()
rb = T._RoundBF16.apply
Wn = {k: (rb(torch.from_numpy(v)) if k.endswith("/weights") else torch.from_numpy(v)) for k, v in W.items()}
conv_bn0, bottleneck0, resnet0 = resnet_v1._conv_bn, resnet_v1.bottleneck, resnet_v1.resnet_v1_50
inner = {}


def conv_bn(x, Wd, scope, **kw):
    y = conv_bn0(x, Wd, scope, **kw)
    y = y if scope.endswith("/conv3") else rb(y)
    inner[scope] = y
    return y


ep_e, ep_f = {}, {}
with torch.no_grad():
    with mock.patch.object(resnet_v1, "_conv_bn", conv_bn), \
            mock.patch.object(resnet_v1, "bottleneck", lambda *a, **k: rb(bottleneck0(*a, **k))), \
            mock.patch.object(resnet_v1, "resnet_v1_50", lambda im, *a, **k: resnet0(rb(im), *a, **k)):
        pose_net.extract_features(torch.from_numpy(frames.astype(np.float32)), Wn, ep_e)
    pose_net.extract_features(torch.from_numpy(frames.astype(np.float32)), {k: torch.from_numpy(v) for k, v in W.items()}, ep_f)
eng = Engine(T.NJ)
eng.load_weights(W)
eng.keep_activations(True)
eng.forward(torch.from_numpy(frames).cuda())
for name in ep_e:
    g = torch.from_numpy(eng.get_activation(name))
    e, f = ep_e[name], ep_f[name]
    mx = f.abs().max().item()
    print("%-50s gpu-emul max %.4f rms %.5f | gpu-fp32 max %.4f rms %.5f | emul-fp32 rms %.5f  (rel to max %.3g)" % (
        name[-50:], (g - e).abs().max().item() / mx, (g - e).pow(2).mean().sqrt().item() / mx,
        (g - f).abs().max().item() / mx, (g - f).pow(2).mean().sqrt().item() / mx, (e - f).pow(2).mean().sqrt().item() / mx, mx))
for name in ("resnet_v1_50/block1/unit_1/bottleneck_v1/shortcut", "resnet_v1_50/block1/unit_1/bottleneck_v1/conv1",
             "resnet_v1_50/block1/unit_1/bottleneck_v1/conv2"):
    g = torch.from_numpy(eng.get_activation(name))
    e = inner[name]
    mx = e.abs().max().item()
    print("%-50s gpu-emul max %.4f rms %.5f" % (name[-50:], (g - e).abs().max().item() / mx, (g - e).pow(2).mean().sqrt().item() / mx))
eng.close()
