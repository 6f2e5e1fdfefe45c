"""tcgen05 implicit-GEMM conv kernel vs the oracle's TF-semantics conv (through the C ABI: dgp_conv2d)."""
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


def q(a, dtype):
    return torch.from_numpy(np.asarray(a, np.float32)).to(dtype).float()


CASES = [
    # N,H,W,Cin,Cout,R,stride,dil,pad_mode,bn,res,res_sub,relu,block_n,out_f32
    (1, 8, 16, 64, 64, 1, 1, 1, 0, False, False, 1, False, 0, True),      # M == 128 exactly
    (2, 13, 17, 64, 64, 1, 1, 1, 0, False, False, 1, False, 0, True),     # M tail
    (2, 13, 17, 256, 128, 1, 1, 1, 0, True, False, 1, True, 0, False),
    (2, 13, 17, 128, 512, 1, 1, 1, 0, True, True, 1, True, 0, False),     # residual + relu, 2 n-blocks
    (2, 13, 17, 128, 512, 1, 1, 1, 0, True, True, 1, True, 128, False),
    (1, 9, 11, 512, 48, 1, 1, 1, 0, False, False, 1, False, 48, True),    # small odd N (head-like GEMM)
    (1, 20, 20, 2048, 256, 1, 1, 1, 0, True, False, 1, True, 0, False),   # 32 k-blocks: ring wraps 4x
    (2, 13, 17, 64, 64, 3, 1, 1, 1, True, False, 1, True, 0, False),      # 3x3 SAME (im2col TMA)
    (1, 21, 19, 128, 128, 3, 1, 1, 1, False, False, 1, False, 0, True),
    (2, 13, 17, 64, 64, 3, 2, 1, 1, True, False, 1, True, 0, False),      # conv2d_same stride 2, odd size
    (1, 20, 18, 128, 128, 3, 2, 1, 1, True, False, 1, True, 0, False),    # conv2d_same stride 2, even size
    (1, 15, 19, 512, 512, 3, 1, 2, 1, True, False, 1, True, 0, False),    # dilation 2 (block4)
    (2, 7, 9, 64, 256, 1, 1, 1, 0, True, True, 2, True, 0, False),        # identity shortcut subsampled by 2
    (3, 47, 52, 256, 1024, 1, 1, 1, 0, True, True, 1, True, 0, False),    # block3 shapes, > 148 tiles (persistent loop)
    (3, 47, 52, 256, 256, 3, 1, 1, 1, True, False, 1, True, 0, False),
    (1, 5, 5, 64, 64, 3, 1, 1, 2, False, False, 1, False, 0, True),       # VALID
    (3, 47, 52, 64, 64, 3, 1, 1, 1, True, False, 1, True, 0, False),      # shared-memory resident patch path: many tiles,
    (2, 101, 75, 64, 64, 3, 1, 1, 0, True, False, 1, False, 0, False),    # ragged tile edges, no ReLU
    (1, 33, 40, 64, 64, 3, 1, 2, 1, True, False, 1, True, 0, False),      # ... and dilation 2
]


@pytest.mark.parametrize("case", CASES, ids=[str(i) for i in range(len(CASES))])
def test_conv_matches_oracle(eng, case):
    N, H, W, Cin, Cout, R, stride, dil, pm, bn, res, res_sub, relu, block_n, out_f32 = case
    rng = np.random.default_rng(hash(case) % (2 ** 31))
    dt = eng.act_dtype
    x = torch.from_numpy(rng.standard_normal((N, H, W, Cin)).astype(np.float32)).to(dt)
    w = (rng.standard_normal((R, R, Cin, Cout)) * np.sqrt(1.0 / (R * R * Cin))).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, Cout).astype(np.float32) if bn else None
    shift = rng.normal(0, 0.2, Cout).astype(np.float32) if bn else None
    wq = q(w, dt)
    if pm == 0:
        ref = tf_ops.conv2d(x.float(), wq, stride, dil, "SAME")
    elif pm == 1:
        ref = tf_ops.conv2d_same(x.float(), wq, stride, dil)
    else:
        ref = tf_ops.conv2d(x.float(), wq, stride, dil, "VALID")
    if bn:
        ref = ref * torch.from_numpy(scale) + torch.from_numpy(shift)
    residual = None
    if res:
        Ho, Wo = ref.shape[1], ref.shape[2]
        residual = torch.from_numpy(rng.standard_normal((N, Ho * res_sub, Wo * res_sub, Cout)).astype(np.float32)).to(dt)
        ref = ref + residual.float()[:, ::res_sub, ::res_sub, :]
    if relu:
        ref = torch.relu(ref)
    got = eng.conv2d(x.cuda(), w, stride, dil, pm, scale, shift, residual.cuda() if res else None, res_sub, relu, out_f32, block_n)
    torch.cuda.synchronize()
    assert got.shape == ref.shape
    err = (got.float().cpu() - ref).abs().max().item() / ref.abs().max().item()
    # fp32 output: accumulation-order noise only; 16-bit output: one rounding (2^-12 relative in fp16, 2^-9 in bf16)
    assert err < (2e-5 if out_f32 else (8e-4 if dt == torch.float16 else 6e-3)), err


def test_conv_fp16_storage():
    from deepgraphpose_b200.engine import Engine
    e = Engine(4, precision="fp16")
    rng = np.random.default_rng(5)
    x = torch.from_numpy(rng.standard_normal((2, 13, 17, 128)).astype(np.float32)).half()
    w = (rng.standard_normal((3, 3, 128, 256)) * np.sqrt(1.0 / (9 * 128))).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, 256).astype(np.float32)
    shift = rng.normal(0, 0.2, 256).astype(np.float32)
    res = torch.from_numpy(rng.standard_normal((2, 13, 17, 256)).astype(np.float32)).half()
    wq = torch.from_numpy(w).half().float()
    ref = torch.relu(tf_ops.conv2d(x.float(), wq, 1, 1, "SAME") * torch.from_numpy(scale) + torch.from_numpy(shift) + res.float())
    got = e.conv2d(x.cuda(), w, 1, 1, 1, scale, shift, res.cuda(), 1, True, False, 0)
    assert got.dtype == torch.float16
    assert (got.float().cpu() - ref).abs().max().item() / ref.abs().max().item() < 1e-3
    e.close()


def test_conv_rejects_bad_channels(eng):
    from deepgraphpose_b200._lib import DgpError
    x = torch.zeros(1, 8, 8, 48, dtype=eng.act_dtype, device="cuda")
    with pytest.raises(DgpError):
        eng.conv2d(x, np.zeros((1, 1, 48, 64), np.float32))
