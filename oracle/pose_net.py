"""DeepLabCut PoseNet restated (oracle only).

Follows /root/reference/src/DeepLabCut/deeplabcut/pose_estimation_tensorflow/nnet/pose_net.py:
``prediction_layer`` :18-26, ``PoseNet.extract_features`` :36-54, ``test`` :84-90,
``inference`` :92-163; nnet/predict.py ``extract_cnn_output`` :45-60,
``argmax_pose_predict`` :62-77; default_config.py :16-59 for the constants.
"""
import numpy as np
import torch

from . import resnet_v1, tf_ops

MEAN_PIXEL = (123.68, 116.779, 103.939)  # default_config.py:23 (RGB)
STRIDE = 8.0  # default_config.py:18
LOCREF_STDEV = 7.2801  # default_config.py:29


def extract_features(inputs, W, end_points=None):
    """pose_net.py:36-54. inputs: (N,H,W,3) float32 holding 0..255 pixel values."""
    mean = torch.tensor(MEAN_PIXEL, dtype=torch.float32).view(1, 1, 1, 3)
    im_centered = inputs - mean
    return resnet_v1.resnet_v1_50(im_centered, W, 16, end_points)


def prediction_layer(net, W, name):
    """pose_net.py:18-26 under variable_scope('pose'): 3x3 stride-2 SAME deconv + bias."""
    scope = "pose/%s/block4" % name
    return tf_ops.conv2d_transpose_same_s2(net, W[scope + "/weights"], W[scope + "/biases"])


def get_net(inputs, W, location_refinement=True):
    net = extract_features(inputs, W)
    out = {"part_pred": prediction_layer(net, W, "part_pred")}
    if location_refinement:
        out["locref"] = prediction_layer(net, W, "locref_pred")
    return out


def test(inputs, W, location_refinement=True):
    """pose_net.py:84-90."""
    heads = get_net(inputs, W, location_refinement)
    out = {"part_prob": torch.sigmoid(heads["part_pred"])}
    if location_refinement:
        out["locref"] = heads["locref"]
    return out


def extract_cnn_output(scmap_np, locref_np, locref_stdev=LOCREF_STDEV):
    """predict.py:45-60 (numpy). scmap (1,H,W,nj) ; locref (1,H,W,2nj) or None."""
    if locref_np is not None:
        locref = np.squeeze(locref_np).copy()
        shape = locref.shape
        locref = np.reshape(locref, (shape[0], shape[1], -1, 2))
        locref *= locref_stdev
    else:
        locref = None
    scmap = np.squeeze(scmap_np)
    if len(scmap.shape) == 2:
        scmap = np.expand_dims(scmap, axis=2)
    return scmap, locref


def argmax_pose_predict(scmap, offmat, stride=STRIDE):
    """predict.py:62-77 (numpy): global first-max peak + locref offset -> (x, y, likelihood).

    Also returns the integer peaks (row, col) per joint, which is what the GPU path
    must reproduce bit-exactly from the same fp32 scoremap.
    """
    num_joints = scmap.shape[2]
    pose, peaks = [], []
    for joint_idx in range(num_joints):
        maxloc = np.unravel_index(np.argmax(scmap[:, :, joint_idx]), scmap[:, :, joint_idx].shape)
        if offmat is None:
            offset = 0
        else:
            offset = np.array(offmat[maxloc][joint_idx])[::-1]
        pos_f8 = np.array(maxloc).astype("float") * stride + 0.5 * stride + offset
        pose.append(np.hstack((pos_f8[::-1], [scmap[maxloc][joint_idx]])))
        peaks.append(maxloc)
    return np.array(pose), np.array(peaks, dtype=np.int64)
