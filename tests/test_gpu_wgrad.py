"""Backward GEMMs through the C ABI vs torch autograd of the oracle's TF-semantics conv:
  * dgp_conv2d_wgrad -- the tcgen05 weight-gradient kernel (MN-major TMA tiles, reduction over pixels, split + fixed-order
    reduce), every conv geometry of the ResNet-50 path;
  * the data gradient expressed as a stride-1 conv with flipped / transposed weights through dgp_conv2d (stride-2 convs:
    zero-inserted dy), which is how train.cu builds its dgrad steps."""
import numpy as np
import pytest
import torch

from oracle import tf_ops

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["fp16", "bf16"])
def eng(request):
    """Both 16-bit storage modes of the same kernels: fp16 (the default of the shims / bench.py) and bf16."""
    from deepgraphpose_b200.engine import Engine
    e = Engine(4, precision=request.param)
    yield e
    e.close()


def ref_conv(x, w, stride, dil, pm):
    if pm == 0:
        return tf_ops.conv2d(x, w, stride, dil, "SAME")
    if pm == 1:
        return tf_ops.conv2d_same(x, w, stride, dil)
    return tf_ops.conv2d(x, w, stride, dil, "VALID")


CASES = [
    # N, H, W, Cin, Cout, R, stride, dil, pad_mode
    (1, 8, 16, 64, 64, 1, 1, 1, 0),       # one pixel block, Cout < 128 (zero-filled M rows)
    (2, 13, 17, 128, 128, 1, 1, 1, 0),    # pixel tail
    (2, 13, 17, 256, 512, 1, 1, 1, 0),    # 4 m-blocks
    (2, 13, 17, 64, 64, 3, 1, 1, 1),      # 3x3: Kw = 576 -> n-block tail of one 64-column chunk
    (2, 13, 17, 128, 128, 3, 2, 1, 1),    # conv2d_same stride 2 (odd size)
    (1, 20, 18, 128, 128, 3, 2, 1, 1),    # conv2d_same stride 2 (even size)
    (1, 15, 19, 512, 512, 3, 1, 2, 1),    # dilation 2 (block4)
    (3, 47, 52, 256, 64, 1, 1, 1, 0),     # many pixel blocks -> many splits
    (2, 30, 40, 2048, 64, 1, 1, 1, 0),    # 8 n-blocks
    (4, 47, 52, 64, 64, 1, 1, 1, 0),      # single output tile, > 32 splits (split-parallel reduce)
]


@pytest.mark.parametrize("case", CASES, ids=[str(i) for i in range(len(CASES))])
def test_wgrad_matches_autograd(eng, case):
    N, H, W, Cin, Cout, R, stride, dil, pm = case
    rng = np.random.default_rng(abs(hash(case)) % (2 ** 31))
    x = torch.from_numpy(rng.standard_normal((N, H, W, Cin)).astype(np.float32)).to(eng.act_dtype)
    w = torch.zeros((R, R, Cin, Cout), requires_grad=True)
    y = ref_conv(x.float(), w, stride, dil, pm)
    dy = torch.from_numpy(rng.standard_normal(tuple(y.shape)).astype(np.float32)).to(eng.act_dtype)
    (y * dy.float()).sum().backward()
    ref = w.grad.permute(3, 0, 1, 2).reshape(Cout, R * R * Cin)  # kernel layout [Cout][tap][Cin]
    got = eng.conv2d_wgrad(x.cuda(), dy.cuda(), R, stride, dil, pm).cpu()
    # 16-bit x 16-bit products are exact in fp32; only the fp32 accumulation order differs
    assert (got - ref).abs().max().item() <= 2e-5 * ref.abs().max().item()


def test_wgrad_is_bitwise_reproducible(eng):
    rng = np.random.default_rng(1)
    x = torch.from_numpy(rng.standard_normal((3, 47, 52, 256)).astype(np.float32)).to(eng.act_dtype).cuda()
    dy = torch.from_numpy(rng.standard_normal((3, 47, 52, 256)).astype(np.float32)).to(eng.act_dtype).cuda()
    a = eng.conv2d_wgrad(x, dy, 3, 1, 1, 1)
    b = eng.conv2d_wgrad(x, dy, 3, 1, 1, 1)
    assert torch.equal(a, b)


@pytest.mark.parametrize("case", CASES[:7], ids=[str(i) for i in range(7)])
def test_dgrad_as_flipped_conv_matches_autograd(eng, case):
    N, H, W, Cin, Cout, R, stride, dil, pm = case
    rng = np.random.default_rng(abs(hash(case)) % (2 ** 31) + 1)
    x = torch.zeros((N, H, W, Cin), requires_grad=True)
    w = (rng.standard_normal((R, R, Cin, Cout)) * np.sqrt(1.0 / (R * R * Cin))).astype(np.float32)
    wq = torch.from_numpy(w).to(eng.act_dtype).float()
    y = ref_conv(x, wq, stride, dil, pm)
    dy = torch.from_numpy(rng.standard_normal(tuple(y.shape)).astype(np.float32)).to(eng.act_dtype)
    (y * dy.float()).sum().backward()
    wd = np.ascontiguousarray(w[::-1, ::-1].transpose(0, 1, 3, 2))
    if stride == 1:
        dyu = dy
    else:
        dyu = torch.zeros((N, H, W, Cout), dtype=eng.act_dtype)
        dyu[:, ::stride, ::stride, :][:, :dy.shape[1], :dy.shape[2]] = dy
    got = eng.conv2d(dyu.cuda(), wd, 1, dil, 1 if R > 1 else 0, None, None, None, 1, False, True, 0).cpu()
    assert (got - x.grad).abs().max().item() <= 2e-5 * x.grad.abs().max().item()
