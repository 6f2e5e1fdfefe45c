"""Known-answer tests that pin the oracle's TF/slim op semantics (the reference has no golden vectors)."""
import numpy as np
import torch

from oracle import resnet_v1, tf_ops


def test_same_pad_rule():
    # TF: pad_total = max((out-1)*s + k_eff - in, 0), extra pixel at the END
    assert tf_ops.same_pad(374, 3, 2) == (0, 1, 187)   # even input: (0,1)
    assert tf_ops.same_pad(187, 3, 2) == (1, 1, 94)    # odd input: (1,1)
    assert tf_ops.same_pad(47, 3, 1, rate=2) == (2, 2, 47)
    assert tf_ops.same_pad(52, 1, 1) == (0, 0, 52)


def test_max_pool_same_even_odd():
    x = torch.arange(6, dtype=torch.float32).view(1, 1, 6, 1).expand(1, 6, 6, 1).contiguous()
    y = tf_ops.max_pool2d_same(x, 3, 2)
    # even size 6 -> out 3, windows [0,1,2],[2,3,4],[4,5,(pad)]
    assert y.shape == (1, 3, 3, 1)
    assert y[0, 0, :, 0].tolist() == [2.0, 4.0, 5.0]
    x = torch.arange(5, dtype=torch.float32).view(1, 1, 5, 1).expand(1, 5, 5, 1).contiguous()
    y = tf_ops.max_pool2d_same(x, 3, 2)
    # odd size 5 -> pad (1,1), windows [(pad),0,1],[1,2,3],[3,4,(pad)]
    assert y[0, 0, :, 0].tolist() == [1.0, 3.0, 4.0]
    # padding never wins, even for all-negative inputs
    y = tf_ops.max_pool2d_same(-torch.ones(1, 4, 4, 1), 3, 2)
    assert torch.all(y == -1)


def test_conv2d_same_strided_is_explicit_pad():
    # conv2d_same(stride 2) on an EVEN size differs from plain SAME: pad (1,1) instead of (0,1)
    x = torch.zeros(1, 4, 4, 1)
    x[0, 0, 0, 0] = 1.0
    w = torch.zeros(3, 3, 1, 1)
    w[0, 0, 0, 0] = 1.0  # picks x[2i-1+0, 2j-1+0] with explicit padding
    y = tf_ops.conv2d_same(x, w, 2)
    assert y.shape == (1, 2, 2, 1)
    assert y.abs().sum() == 0  # tap (0,0) reads the padding for output (0,0); x[0,0] is never at an offset-0 tap
    w = torch.zeros(3, 3, 1, 1)
    w[1, 1, 0, 0] = 1.0
    assert tf_ops.conv2d_same(x, w, 2)[0, 0, 0, 0] == 1.0
    y_same = tf_ops.conv2d(x, w, 2, 1, "SAME")  # plain SAME: pad_beg 0 -> centre tap reads x[1,1]
    assert y_same[0, 0, 0, 0] == 0.0


def test_conv2d_transpose_alignment():
    # out[2i+k] += x[i] * w[k], cropped to 2n  (== gradient of the SAME stride-2 forward conv)
    x = torch.zeros(1, 3, 3, 1)
    x[0, 1, 1, 0] = 1.0
    w = torch.arange(9, dtype=torch.float32).view(3, 3, 1, 1)
    y = tf_ops.conv2d_transpose_same_s2(x, w)
    assert y.shape == (1, 6, 6, 1)
    exp = torch.zeros(6, 6)
    exp[2:5, 2:5] = torch.arange(9, dtype=torch.float32).view(3, 3)
    assert torch.equal(y[0, :, :, 0], exp)
    # crop: impulse at the last input pixel loses its k=2 row/col
    x = torch.zeros(1, 3, 3, 1)
    x[0, 2, 2, 0] = 1.0
    y = tf_ops.conv2d_transpose_same_s2(x, w)
    assert torch.equal(y[0, 4:, 4:, 0], torch.tensor([[0.0, 1.0], [3.0, 4.0]]))
    # it is the adjoint of the SAME stride-2 conv: <conv(a), b> == <a, conv_T(b)>
    g = torch.Generator().manual_seed(0)
    a = torch.randn(1, 8, 10, 2, generator=g)
    b = torch.randn(1, 4, 5, 3, generator=g)
    wf = torch.randn(3, 3, 2, 3, generator=g)  # forward HWIO (in=2, out=3) == transpose layout [kh,kw,cout=2,cin=3]
    lhs = (tf_ops.conv2d(a, wf, 2, 1, "SAME") * b).sum()
    rhs = (a * tf_ops.conv2d_transpose_same_s2(b, wf)).sum()
    assert abs(lhs - rhs) < 1e-3


def test_shape_identities():
    assert resnet_v1.output_dims(747, 832) == ((47, 52), (94, 104))
    assert resnet_v1.output_dims(470, 640) == ((30, 40), (60, 80))
    assert resnet_v1.output_dims(1024, 1280) == ((64, 80), (128, 160))


def test_unit_plan_output_stride_16():
    plan = resnet_v1.unit_plan(16)
    assert len(plan) == 16
    strides = [p[3] for p in plan]
    rates = [p[4] for p in plan]
    assert strides == [1, 1, 2, 1, 1, 1, 2] + [1] * 9       # only block1/unit_3 and block2/unit_4 stride
    assert rates == [1] * 13 + [2, 2, 2]                      # block4 runs dilated
    assert [p[1] for p in plan][-1] == 2048


def test_resnet_small_forward_shapes():
    from deepgraphpose_b200 import synthetic
    from oracle import pose_net
    W = {k: torch.from_numpy(v) for k, v in synthetic.make_weights(3, seed=1).items()}
    x = torch.zeros(1, 70, 90, 3)
    ep = {}
    with torch.no_grad():
        net = pose_net.extract_features(x, W, ep)
        pred = pose_net.prediction_layer(net, W, "part_pred")
        loc = pose_net.prediction_layer(net, W, "locref_pred")
    assert net.shape == (1, 5, 6, 2048)
    assert pred.shape == (1, 10, 12, 3) and loc.shape == (1, 10, 12, 6)
    assert ep["resnet_v1_50/conv1"].shape == (1, 35, 45, 64)
    assert ep["resnet_v1_50/pool1"].shape == (1, 18, 23, 64)


def test_bn_and_weighted_loss():
    x = torch.tensor([[1.0, 2.0]])
    y = tf_ops.batch_norm_inference(x, torch.tensor([2.0, 1.0]), torch.tensor([0.5, 0.0]), torch.tensor([1.0, 0.0]),
                                    torch.tensor([4.0 - 1e-5, 1.0 - 1e-5]))
    assert torch.allclose(y, torch.tensor([[0.5, 2.0]]), atol=1e-6)
    # SUM_BY_NONZERO_WEIGHTS counts broadcast weights != 0
    losses = torch.ones(2, 3, 4)
    w = torch.tensor([1.0, 0.0]).view(2, 1, 1)
    assert tf_ops.compute_weighted_loss(losses, w) == 1.0           # 12 / 12
    assert tf_ops.compute_weighted_loss(losses * 2, 1.0) == 2.0
    assert tf_ops.compute_weighted_loss(losses, torch.zeros(2, 1, 1)) == 0.0
    # sigmoid CE closed form
    z, xl = torch.tensor([0.3]), torch.tensor([-1.2])
    ref = -(z * torch.log(torch.sigmoid(xl)) + (1 - z) * torch.log(1 - torch.sigmoid(xl)))
    assert torch.allclose(tf_ops.sigmoid_cross_entropy_with_logits(z, xl), ref, atol=1e-6)
