"""Drop-in for deeplabcut.pose_estimation_tensorflow.nnet.pose_net on the DGP path
(reference: src/DeepLabCut/deeplabcut/pose_estimation_tensorflow/nnet/pose_net.py).

``PoseNet(cfg)`` keeps the reference's method names.  Where the reference returns TF graph nodes, these return
CUDA tensors computed by the engine: ``get_net`` / ``test`` / ``inference`` take uint8 frames (N,H,W,3).
The ResNet trunk and the two deconv heads are fused into one C-ABI call (``dgp_forward``), so
``extract_features`` + ``prediction_layers`` are exposed together as ``get_net``.
"""
import numpy as np
import torch

from .engine import Engine, LOCREF_STDEV, MEAN_PIXEL, STRIDE


def _get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


class PoseNet:
    def __init__(self, cfg, variables=None, device=None):
        """pose_net.py:28-34 (defaults output_stride=16, deconvolutionstride=2 are the only supported values)."""
        self.cfg = cfg
        if _get(cfg, "output_stride", 16) != 16 or _get(cfg, "deconvolutionstride", 2) != 2:
            raise ValueError("the B200 path implements output_stride=16, deconvolutionstride=2 (pose_net.py:31-34)")
        if _get(cfg, "net_type", "resnet_50") != "resnet_50":
            raise ValueError("only resnet_50 is on the B200 path")
        self.num_joints = int(_get(cfg, "num_joints"))
        self.location_refinement = bool(_get(cfg, "location_refinement", True))
        self.engine = Engine(self.num_joints, self.location_refinement, device,
                             float(_get(cfg, "stride", STRIDE)), float(_get(cfg, "locref_stdev", LOCREF_STDEV)),
                             tuple(_get(cfg, "mean_pixel", MEAN_PIXEL)), _get(cfg, "precision", "bf16"))
        if variables is not None:
            self.restore(variables)

    def restore(self, variables):
        """Replaces Saver.restore: {tf_var_name: ndarray} with slim names (resnet_v1_50/..., pose/...)."""
        self.engine.load_weights(variables)

    def _frames(self, inputs):
        if isinstance(inputs, torch.Tensor):
            t = inputs
        else:
            t = torch.from_numpy(np.ascontiguousarray(inputs))
        if t.dtype != torch.uint8:
            t = t.to(torch.uint8)
        return t.to(self.engine.device)

    def get_net(self, inputs):
        """pose_net.py:80-82 -> {'part_pred': logits (N,2h,2w,nj), 'locref': (N,2h,2w,2nj)}."""
        logits, locref = self.engine.forward(self._frames(inputs))
        out = {"part_pred": logits}
        if self.location_refinement:
            out["locref"] = locref
        return out

    def test(self, inputs):
        """pose_net.py:84-90."""
        heads = self.get_net(inputs)
        out = {"part_prob": self.engine.sigmoid(heads["part_pred"])}
        if self.location_refinement:
            out["locref"] = heads["locref"]
        return out

    def inference(self, inputs):
        """pose_net.py:92-163: {'pose': (N*nj, 3)} rows (x, y, likelihood), frame-major like the batched TF version."""
        heads = self.get_net(inputs)
        r = self.engine.softargmax(heads["part_pred"], heads.get("locref"), want=("dlc_pose",))
        return {"pose": r["dlc_pose"].reshape(-1, 3)}
