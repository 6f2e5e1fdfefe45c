"""A TF1-``Session``-shaped front for the engine.

The reference's boundary is ``sess.run(fetches, feed_dict)`` on graph handles
(/root/reference/src/deepgraphpose/models/eval.py:214,328; fitdgp.py:1130-1144,817-818).  ``Handle`` objects
stand in for the TF tensors / placeholders and ``Session.run`` maps a fetch list onto C-ABI calls, returning numpy
arrays of the same shape, dtype and axis order (NHWC, (row, col)) as the reference.
"""
import numpy as np
import torch


class Handle:
    """Stand-in for a tf.Tensor / tf.placeholder of the reference graph."""

    def __init__(self, name, kind):
        self.name = name
        self.kind = kind

    def __repr__(self):
        return "<dgp_b200.Handle %s>" % self.name

    def __hash__(self):
        return id(self)


class Session:
    """Evaluates handles created by ``setup_dgp_eval_graph`` (inference graph)."""

    def __init__(self, engine, handles, gamma=1.0, gauss_len=1.0):
        self.engine = engine
        self.h = handles
        self.gamma = float(gamma)
        self.gauss_len = float(gauss_len)
        self.closed = False

    def run(self, fetches, feed_dict=None):
        if self.closed:
            raise RuntimeError("Attempted to use a closed Session.")
        single = not isinstance(fetches, (list, tuple))
        flist = [fetches] if single else list(fetches)
        feed_dict = feed_dict or {}
        inputs = None
        for k, v in feed_dict.items():
            if k is self.h["inputs"]:
                inputs = v
        if inputs is None:
            raise ValueError("feed_dict must provide the `inputs` placeholder")
        arr = np.asarray(inputs)
        if arr.ndim != 4 or arr.shape[-1] != 3:
            raise ValueError("inputs must have shape [N, H, W, 3]")
        # The reference feeds uint8 pixel values into a float32 placeholder; the kernels take the uint8 directly.
        if arr.dtype != np.uint8:
            if np.any(arr != np.round(arr)) or arr.min() < 0 or arr.max() > 255:
                raise ValueError("inputs must hold 0..255 integer pixel values (uint8 frames)")
            arr = arr.astype(np.uint8)
        dev = self.engine.device
        frames = torch.from_numpy(np.ascontiguousarray(arr)).to(dev, non_blocking=True)
        kinds = {f.kind for f in flist}
        want_locref = "locref" in kinds
        if want_locref and not self.engine.location_refinement:
            raise ValueError("locref was not built (loc_ref=False)")
        logits, locref = self.engine.forward(frames, want_locref=want_locref)
        out = {}
        if "mu_n" in kinds:
            out["mu_n"] = self.engine.softargmax(logits, None, self.gamma, self.gauss_len, want=("mu",))["mu"]
        if "softmax_tensor" in kinds:
            out["softmax_tensor"] = self.engine.softmax_map(logits, self.gamma, self.gauss_len)
        if "scmap" in kinds:
            out["scmap"] = logits
        if "locref" in kinds:
            out["locref"] = locref
        res = [out[f.kind].cpu().numpy() for f in flist]
        return res[0] if single else res

    def close(self):
        self.closed = True

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
