"""slim ``resnet_v1_50`` (output_stride=16, is_training=False) restated on torch CPU (oracle only).

Follows tensorflow==1.15 ``tensorflow/contrib/slim/python/slim/nets/resnet_v1.py`` and
``resnet_utils.py`` as called at
/root/reference/src/DeepLabCut/deeplabcut/pose_estimation_tensorflow/nnet/pose_net.py:14-16,50-52
(``net_fun(im_centered, global_pool=False, output_stride=16, is_training=False)`` under
``resnet_arg_scope()``).  Weights are a ``dict`` keyed by the TF variable names
(``resnet_v1_50/conv1/weights``, ``.../BatchNorm/{gamma,beta,moving_mean,moving_variance}``,
``resnet_v1_50/block{b}/unit_{u}/bottleneck_v1/{shortcut,conv1,conv2,conv3}/...``).
"""
import torch

from . import tf_ops

# (scope, base_depth, num_units, stride)  -- resnet_v1.resnet_v1_50 in TF 1.15
RESNET50_BLOCKS = (
    ("block1", 64, 3, 2),
    ("block2", 128, 4, 2),
    ("block3", 256, 6, 2),
    ("block4", 512, 3, 1),
)


def _conv_bn(x, W, scope, stride=1, rate=1, relu=True, same_explicit=False):
    w = W[scope + "/weights"]
    if same_explicit:
        y = tf_ops.conv2d_same(x, w, stride, rate)
    else:
        y = tf_ops.conv2d(x, w, stride, rate, "SAME")
    y = tf_ops.batch_norm_inference(
        y,
        W[scope + "/BatchNorm/gamma"],
        W[scope + "/BatchNorm/beta"],
        W[scope + "/BatchNorm/moving_mean"],
        W[scope + "/BatchNorm/moving_variance"],
    )
    return torch.relu(y) if relu else y


def bottleneck(x, W, scope, depth, depth_bottleneck, stride, rate):
    """resnet_v1.bottleneck (v1: post-activation, stride on the 3x3)."""
    depth_in = x.shape[-1]
    if depth == depth_in:
        shortcut = tf_ops.subsample(x, stride)
    else:
        shortcut = _conv_bn(x, W, scope + "/shortcut", stride=stride, relu=False)
    r = _conv_bn(x, W, scope + "/conv1")
    r = _conv_bn(r, W, scope + "/conv2", stride=stride, rate=rate, same_explicit=True)
    r = _conv_bn(r, W, scope + "/conv3", relu=False)
    return torch.relu(shortcut + r)


def unit_plan(output_stride=16, blocks=RESNET50_BLOCKS):
    """resnet_utils.stack_blocks_dense bookkeeping -> list of (scope, depth, depth_bn, stride, rate)."""
    target = output_stride // 4  # root block (conv1 s2 + pool s2) already has stride 4
    current_stride, rate = 1, 1
    plan = []
    for name, base, units, bstride in blocks:
        for u in range(units):
            ustride = bstride if u == units - 1 else 1
            scope = "resnet_v1_50/%s/unit_%d/bottleneck_v1" % (name, u + 1)
            if current_stride == target:
                plan.append((scope, base * 4, base, 1, rate))
                rate *= ustride
            else:
                plan.append((scope, base * 4, base, ustride, 1))
                current_stride *= ustride
    return plan


def resnet_v1_50(im_centered, W, output_stride=16, end_points=None):
    """Returns block4 output `net` (N, ceil(H/16), ceil(W/16), 2048)."""
    x = _conv_bn(im_centered, W, "resnet_v1_50/conv1", stride=2, same_explicit=True)
    if end_points is not None:
        end_points["resnet_v1_50/conv1"] = x
    x = tf_ops.max_pool2d_same(x, 3, 2)
    if end_points is not None:
        end_points["resnet_v1_50/pool1"] = x
    for scope, depth, dbn, stride, rate in unit_plan(output_stride):
        x = bottleneck(x, W, scope, depth, dbn, stride, rate)
        if end_points is not None:
            end_points[scope] = x
    return x


def output_dims(h, w):
    """Closed form replacing Dataset._compute_pred_dims (/root/reference/src/deepgraphpose/dataset.py:348-371)."""
    def c2(v):
        return -(-v // 2)
    fh, fw = c2(c2(c2(c2(h)))), c2(c2(c2(c2(w))))
    return (fh, fw), (2 * fh, 2 * fw)
