"""TF 1.15 / tf.contrib.slim op semantics restated with torch CPU ops (oracle only).

Everything here takes and returns NHWC float32 torch tensors so that the code
reads like the TF graph it restates.  Weight layouts are TF's:
conv ``[kh, kw, cin, cout]`` (HWIO), transposed conv ``[kh, kw, cout, cin]``.

Third-party semantics restated (tensorflow==1.15, pinned at /root/reference/README.md:37-38):
* ``tf.nn.conv2d`` / ``slim.conv2d`` SAME padding:
  out = ceil(in/stride); pad_total = max((out-1)*stride + (k-1)*rate + 1 - in, 0);
  pad_beg = pad_total // 2 (the extra pixel goes to the END).
* ``resnet_utils.conv2d_same``: stride 1 -> SAME; stride > 1 -> explicit
  symmetric-ish zero pad of k_eff-1 (beg = (k_eff-1)//2) followed by VALID.
* ``slim.max_pool2d(padding='SAME')``: same pad rule, padding never wins the max.
* ``slim.conv2d_transpose(padding='SAME', stride=2, kernel=3)``: the gradient of
  the SAME forward conv, i.e. out[2i+k] += x[i] * w[k], cropped to 2n.
* ``slim.batch_norm(is_training=False)``: gamma*(x-mean)/sqrt(var+eps)+beta, eps=1e-5
  (resnet_arg_scope defaults).
"""
import math

import torch
import torch.nn.functional as F

BN_EPS = 1e-5  # resnet_utils.resnet_arg_scope(batch_norm_epsilon=1e-5)


def same_pad(in_size, k, stride, rate=1):
    """TF SAME padding amounts (beg, end) and output size for one spatial dim."""
    out = -(-in_size // stride)
    k_eff = (k - 1) * rate + 1
    pad_total = max((out - 1) * stride + k_eff - in_size, 0)
    beg = pad_total // 2
    return beg, pad_total - beg, out


def _nhwc_to_nchw(x):
    return x.permute(0, 3, 1, 2)


def _nchw_to_nhwc(x):
    return x.permute(0, 2, 3, 1)


def conv2d(x, w_hwio, stride=1, rate=1, padding="SAME"):
    """tf.nn.conv2d on NHWC input with HWIO weights."""
    kh, kw = w_hwio.shape[0], w_hwio.shape[1]
    xc = _nhwc_to_nchw(x)
    if padding == "SAME":
        pt, pb, _ = same_pad(x.shape[1], kh, stride, rate)
        pl, pr, _ = same_pad(x.shape[2], kw, stride, rate)
        xc = F.pad(xc, (pl, pr, pt, pb))
    elif padding != "VALID":
        raise ValueError(padding)
    w = w_hwio.permute(3, 2, 0, 1).contiguous()
    y = F.conv2d(xc, w, stride=stride, dilation=rate)
    return _nchw_to_nhwc(y)


def conv2d_same(x, w_hwio, stride, rate=1):
    """resnet_utils.conv2d_same (explicit padding for strided convs)."""
    if stride == 1:
        return conv2d(x, w_hwio, 1, rate, "SAME")
    k = w_hwio.shape[0]
    k_eff = k + (k - 1) * (rate - 1)
    pad_total = k_eff - 1
    beg = pad_total // 2
    end = pad_total - beg
    xp = F.pad(x, (0, 0, beg, end, beg, end))
    return conv2d(xp, w_hwio, stride, rate, "VALID")


def max_pool2d_same(x, k, stride):
    """slim.max_pool2d(..., padding='SAME') on NHWC (pads with -inf)."""
    pt, pb, _ = same_pad(x.shape[1], k, stride)
    pl, pr, _ = same_pad(x.shape[2], k, stride)
    xc = F.pad(_nhwc_to_nchw(x), (pl, pr, pt, pb), value=-math.inf)
    return _nchw_to_nhwc(F.max_pool2d(xc, k, stride))


def subsample(x, factor):
    """resnet_utils.subsample: 1x1 max-pool with stride `factor` (== x[:, ::f, ::f])."""
    if factor == 1:
        return x
    return x[:, ::factor, ::factor, :]


def batch_norm_inference(x, gamma, beta, mean, var, eps=BN_EPS):
    return (x - mean) * (gamma / torch.sqrt(var + eps)) + beta


def conv2d_transpose_same_s2(x, w_hwoi, bias=None):
    """slim.conv2d_transpose(kernel 3x3, stride 2, SAME): out[2i+k] += x[i]*w[k], crop to 2n.

    w_hwoi: [kh, kw, cout, cin] (TF layout for conv2d_transpose filters).
    """
    n, h, w_, c = x.shape
    wt = w_hwoi.permute(3, 2, 0, 1).contiguous()  # torch: [cin, cout, kh, kw]
    y = F.conv_transpose2d(_nhwc_to_nchw(x), wt, stride=2, padding=0)
    y = y[:, :, : 2 * h, : 2 * w_]
    y = _nchw_to_nhwc(y)
    if bias is not None:
        y = y + bias
    return y


def sigmoid_cross_entropy_with_logits(labels, logits):
    """tf.nn.sigmoid_cross_entropy_with_logits: max(x,0) - x*z + log1p(exp(-|x|))."""
    return torch.clamp(logits, min=0) - logits * labels + torch.log1p(torch.exp(-torch.abs(logits)))


def compute_weighted_loss(losses, weights):
    """tf.losses.compute_weighted_loss, reduction=SUM_BY_NONZERO_WEIGHTS.

    sum(losses * w) / #{elements of broadcast(w) that are != 0}  (0 if that count is 0).
    """
    w = torch.as_tensor(weights, dtype=losses.dtype)
    wl = losses * w
    present = torch.broadcast_to((w != 0).to(losses.dtype), losses.shape).sum()
    total = wl.sum()
    return torch.where(present > 0, total / torch.clamp(present, min=1.0), torch.zeros_like(total))
