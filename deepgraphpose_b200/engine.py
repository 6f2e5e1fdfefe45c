"""Thin host-side wrapper of the C ABI: PyTorch supplies device memory and streams, nothing else.

``Engine`` is the object behind the reference-shaped shims (``pose_net.PoseNet``, ``fitdgp_util.argmax_2d_from_cm``,
``eval.setup_dgp_eval_graph`` / ``estimate_pose``).  Every method ends in a call into libdgp_b200.so.
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import DgpConfig, DgpError, check

MEAN_PIXEL = (123.68, 116.779, 103.939)  # PTF/default_config.py:23
STRIDE = 8.0                             # PTF/default_config.py:18
LOCREF_STDEV = 7.2801                    # PTF/default_config.py:29


def output_dims(H, W):
    """(h_feat, w_feat), (h_out, w_out) -- closed form of Dataset._compute_pred_dims (dataset.py:348-371)."""
    lib = _lib.load()
    a, b, c, d = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    check(lib.dgp_output_dims(int(H), int(W), C.byref(a), C.byref(b), C.byref(c), C.byref(d)))
    return (a.value, b.value), (c.value, d.value)


def suggest_batch(H, W, lo=24, hi=48, num_sms=148):
    """Frames per batch that minimise wave quantisation of the persistent conv-GEMM grid (one CTA per SM).

    Every GEMM layer runs ceil(B*pixels/tile_rows) * n_blocks tiles on `num_sms` CTAs; a batch size for which the
    dominant layers' tile counts are (just under) a multiple of the SM count avoids a nearly empty last wave.  Layers are
    weighted by their roofline bound (max of bf16 tensor time and HBM time), tile shapes follow csrc/capi.cu."""
    c2 = lambda v: -(-v // 2)
    layers = []  # (pixels per frame, tile rows, n_blocks, weight)
    hh, ww = c2(H), c2(W)
    def add(px, cin_k, cout, bytes_per_px):
        bn = min(cout, 256)
        rows = 256 if bn <= 128 else 128
        wgt = max(2.0 * px * cin_k * cout / 1414e12, px * bytes_per_px / 6464e9)
        layers.append((px, rows, -(-cout // bn), wgt))
    add(hh * ww, 256, 64, 32 + 128)
    hh, ww = c2(hh), c2(ww)
    cin = 64
    for base, units, bstride in ((64, 3, 2), (128, 4, 2), (256, 6, 1), (512, 3, 1)):
        for u in range(units):
            s_ = bstride if u == units - 1 else 1
            depth = 4 * base
            ho, wo = (c2(hh), c2(ww)) if s_ == 2 else (hh, ww)
            if cin != depth:
                add(hh * ww, cin, depth, 2 * (cin + depth))
            add(hh * ww, cin, base, 2 * (cin + base))
            add(ho * wo, 9 * base, base, 4 * base)
            add(ho * wo, base, depth, 2 * (base + 2 * depth))
            hh, ww, cin = ho, wo, depth
    best, best_b = None, lo
    for B in range(lo, hi + 1):
        t = 0.0
        for px, rows, nb, wgt in layers:
            tiles = -(-(px * B) // rows) * nb
            waves = -(-tiles // num_sms)
            t += wgt * waves * num_sms / tiles
        t /= B ** 0.0  # weights are per frame already
        if best is None or t < best - 1e-12:
            best, best_b = t, B
    return best_b


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


class _Raw:
    """__cuda_array_interface__ carrier for memory owned by the handle."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"data": (int(ptr), False), "shape": tuple(shape), "typestr": typestr, "version": 2}


def _device_view(ptr, shape, typestr, device):
    with torch.cuda.device(device):
        return torch.as_tensor(_Raw(ptr, shape, typestr), device=device)


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class Engine:
    def __init__(self, num_joints, location_refinement=True, device=None, stride=STRIDE, locref_stdev=LOCREF_STDEV,
                 mean_pixel=MEAN_PIXEL, precision="fp16"):
        if not torch.cuda.is_available():
            raise DgpError(-2, "no CUDA device: deepgraphpose_b200 has no CPU fallback")
        self.lib = _lib.load()
        if device is None:
            device = torch.cuda.current_device()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        self.nj = int(num_joints)
        self.location_refinement = bool(location_refinement)
        cfg = DgpConfig()
        cfg.num_joints = self.nj
        cfg.location_refinement = int(self.location_refinement)
        cfg.device = self.device.index
        cfg.stride = stride
        cfg.locref_stdev = locref_stdev
        cfg.mean_pixel = (C.c_float * 3)(*mean_pixel)
        cfg.bn_epsilon = 1e-5
        if precision not in ("bf16", "fp16"):
            raise ValueError("precision must be 'bf16' or 'fp16'")
        cfg.precision = 1 if precision == "fp16" else 0
        self.precision = precision
        self.act_dtype = torch.float16 if precision == "fp16" else torch.bfloat16
        self.stride = float(stride)
        self.locref_stdev = float(locref_stdev)
        h = C.c_void_p()
        check(self.lib.dgp_create(C.byref(cfg), C.byref(h)))
        self.h = h
        self.weights_loaded = False

    def close(self):
        if getattr(self, "h", None):
            self.lib.dgp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, status):
        check(status, self.h)

    # ------------------------------------------------------------------ weights (Saver.restore replacement)
    def load_weights(self, variables):
        """variables: {tf_var_name: float32 ndarray in TF layout}."""
        for name, arr in variables.items():
            a = np.ascontiguousarray(arr, dtype=np.float32)
            shape = (C.c_int64 * a.ndim)(*a.shape)
            self._check(self.lib.dgp_load_weights(self.h, name.encode(), a.ctypes.data_as(C.c_void_p), shape, a.ndim, 0))
        self._check(self.lib.dgp_finalize_weights(self.h))
        self.weights_loaded = True
        # the frozen moving statistics are not trainable state of the handle; keep them for checkpoint export
        self._frozen = {k: np.array(v, dtype=np.float32) for k, v in variables.items()
                        if k.endswith("/moving_mean") or k.endswith("/moving_variance")}

    # ------------------------------------------------------------------ forward
    def forward(self, frames, want_locref=None):
        """frames: uint8 cuda tensor (B,H,W,3). Returns (logits (B,2h,2w,nj) f32, locref (B,2h,2w,2nj) f32 | None)."""
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3 or not frames.is_cuda:
            raise ValueError("frames must be a uint8 CUDA tensor of shape (B,H,W,3)")
        frames = frames.contiguous()
        B, H, W, _ = frames.shape
        _, (ho, wo) = output_dims(H, W)
        if want_locref is None:
            want_locref = self.location_refinement
        logits = torch.empty((B, ho, wo, self.nj), dtype=torch.float32, device=frames.device)
        locref = torch.empty((B, ho, wo, 2 * self.nj), dtype=torch.float32, device=frames.device) if want_locref else None
        self._check(self.lib.dgp_forward(self.h, _ptr(frames), B, H, W, _ptr(logits), _ptr(locref), _stream(frames.device)))
        return logits, locref

    def extract_features(self, frames):
        """PoseNet.extract_features (pose_net.py:36-54): uint8 cuda frames (B,H,W,3) -> net (B,hf,wf,2048) float32."""
        if frames.dtype != torch.uint8 or frames.dim() != 4 or frames.shape[-1] != 3 or not frames.is_cuda:
            raise ValueError("frames must be a uint8 CUDA tensor of shape (B,H,W,3)")
        frames = frames.contiguous()
        B, H, W, _ = frames.shape
        (hf, wf), _ = output_dims(H, W)
        net = torch.empty((B, hf, wf, 2048), dtype=torch.float32, device=frames.device)
        self._check(self.lib.dgp_extract_features(self.h, _ptr(frames), B, H, W, _ptr(net), _stream(frames.device)))
        return net

    def prediction_layers(self, net, want_locref=None):
        """PoseNet.prediction_layers (pose_net.py:56-78) on a float32 cuda ``net`` (B,hf,wf,2048) -> (logits, locref | None)."""
        if net.dtype != torch.float32 or net.dim() != 4 or net.shape[-1] != 2048 or not net.is_cuda:
            raise ValueError("net must be a float32 CUDA tensor (B,hf,wf,2048)")
        net = net.contiguous()
        B, hf, wf, _ = net.shape
        if want_locref is None:
            want_locref = self.location_refinement
        logits = torch.empty((B, 2 * hf, 2 * wf, self.nj), dtype=torch.float32, device=net.device)
        locref = torch.empty((B, 2 * hf, 2 * wf, 2 * self.nj), dtype=torch.float32, device=net.device) if want_locref else None
        self._check(self.lib.dgp_prediction_layers(self.h, _ptr(net), B, hf, wf, _ptr(logits), _ptr(locref), _stream(net.device)))
        return logits, locref

    def deconv2d(self, x, w, bias=None):
        """slim.conv2d_transpose(x, Cout, [3,3], stride=2, 'SAME') + bias: x float32 cuda (N,H,W,Cin), Cin % 64 == 0;
        w float32 ndarray in the TF layout [3,3,Cout,Cin]; returns float32 (N,2H,2W,Cout)."""
        if x.dtype != torch.float32 or x.dim() != 4 or not x.is_cuda:
            raise ValueError("x must be a float32 CUDA tensor (N,H,W,Cin)")
        x = x.contiguous()
        w = np.ascontiguousarray(w, dtype=np.float32)
        if w.ndim != 4 or w.shape[0] != 3 or w.shape[1] != 3 or w.shape[3] != x.shape[3]:
            raise ValueError("w must be [3,3,Cout,Cin] with Cin = %d" % x.shape[3])
        N, H, W, Cin = x.shape
        Cout = w.shape[2]
        b = np.ascontiguousarray(bias, dtype=np.float32).reshape(-1) if bias is not None else None
        if b is not None and b.size != Cout:
            raise ValueError("bias must have %d entries" % Cout)
        out = torch.empty((N, 2 * H, 2 * W, Cout), dtype=torch.float32, device=x.device)
        self._check(self.lib.dgp_deconv2d(self.h, _ptr(x), N, H, W, Cin, w.ctypes.data_as(C.c_void_p),
                                          b.ctypes.data_as(C.c_void_p) if b is not None else None, Cout, _ptr(out),
                                          _stream(x.device)))
        return out

    def softargmax(self, logits, locref=None, gamma=1.0, gauss_len=1.0, want=("mu", "peak", "lik", "dlc_peak", "dlc_pose")):
        """Fused argmax_2d_from_cm + estimate_pose read-out + DLC argmax pose. Returns a dict of cuda tensors."""
        if logits.dtype != torch.float32 or logits.dim() != 4 or not logits.is_cuda:
            raise ValueError("logits must be a float32 CUDA tensor (B,H,W,nj)")
        logits = logits.contiguous()
        B, H, W, nj = logits.shape
        dev = logits.device
        out = {}
        out["mu"] = torch.empty((B, nj, 2), dtype=torch.float32, device=dev) if "mu" in want else None
        out["peak"] = torch.empty((B, nj, 2), dtype=torch.int32, device=dev) if "peak" in want else None
        out["lik"] = torch.empty((B, nj), dtype=torch.float32, device=dev) if "lik" in want else None
        out["dlc_peak"] = torch.empty((B, nj, 2), dtype=torch.int32, device=dev) if "dlc_peak" in want else None
        out["dlc_pose"] = torch.empty((B, nj, 3), dtype=torch.float32, device=dev) if "dlc_pose" in want else None
        if locref is not None:
            locref = locref.contiguous()
            if locref.shape != (B, H, W, 2 * nj) or locref.dtype != torch.float32:
                raise ValueError("locref must be float32 (B,H,W,2*nj)")
        self._check(self.lib.dgp_softargmax(self.h, _ptr(logits), _ptr(locref), B, H, W, nj, float(gamma), float(gauss_len),
                                            _ptr(out["mu"]), _ptr(out["peak"]), _ptr(out["lik"]), _ptr(out["dlc_peak"]),
                                            _ptr(out["dlc_pose"]), _stream(dev)))
        return {k: v for k, v in out.items() if v is not None}

    def softmax_map(self, logits, gamma=1.0, gauss_len=1.0):
        """Second output of argmax_2d_from_cm: blurred, renormalised spatial softmax (B,H,W,nj)."""
        logits = logits.contiguous()
        B, H, W, nj = logits.shape
        out = torch.empty_like(logits)
        self._check(self.lib.dgp_softmax_map(self.h, _ptr(logits), B, H, W, nj, float(gamma), float(gauss_len), _ptr(out),
                                             _stream(logits.device)))
        return out

    def soft_pose(self, logits, locref, gamma=1.0, gauss_len=1.0, swap_offsets=False):
        """evaluate_dgp's 'dgp' locref branch (eval.py:751-785) on the device: (B,nj,3) = (x, y, 1)."""
        logits, locref = logits.contiguous(), locref.contiguous()
        B, H, W, nj = logits.shape
        if locref.shape != (B, H, W, 2 * nj) or locref.dtype != torch.float32 or logits.dtype != torch.float32:
            raise ValueError("logits (B,H,W,nj) and locref (B,H,W,2*nj) must be float32")
        ws = torch.empty_like(logits)
        pose = torch.empty((B, nj, 3), dtype=torch.float32, device=logits.device)
        self._check(self.lib.dgp_soft_pose(self.h, _ptr(logits), _ptr(locref), B, H, W, nj, float(gamma), float(gauss_len),
                                           int(swap_offsets), _ptr(ws), _ptr(pose), _stream(logits.device)))
        return pose

    def sigmoid(self, logits):
        logits = logits.contiguous()
        out = torch.empty_like(logits)
        self._check(self.lib.dgp_sigmoid(self.h, _ptr(logits), _ptr(out), logits.numel(), _stream(logits.device)))
        return out

    def potentials(self, mu, edges, halo_next=None, ws=None, ws_max=None, wt_max=0.0, out=None):
        """mu (T,nj,2) f32 cuda; edges int32 (nl,2). Returns dict(skel (nl,T), temporal (T or T-1, nj), e_skel (T), e_temp (T)).
        ``out`` may carry the dict of a previous call to reuse its buffers (steady-state loops allocate nothing)."""
        mu = mu.contiguous()
        T, nj, _ = mu.shape
        dev = mu.device
        key = (tuple(map(tuple, np.asarray(edges, dtype=np.int64).reshape(-1, 2).tolist())), str(dev))
        cache = self.__dict__.setdefault("_edge_cache", {})
        if key not in cache:
            cache[key] = torch.as_tensor(np.asarray(edges, dtype=np.int32).reshape(-1, 2), device=dev)
        edges_t = cache[key]
        nl = edges_t.shape[0]
        if out is not None and out["_T"] == T and out["_nl"] == nl:
            skel, temporal, e_skel, e_temp = out["skel"], out["_temporal_full"], out["e_skel"], out["e_temp"]
        else:
            skel = torch.empty((nl, T), dtype=torch.float32, device=dev)
            temporal = torch.empty((T, nj), dtype=torch.float32, device=dev)
            e_skel = torch.empty((T,), dtype=torch.float32, device=dev) if ws is not None else None
            e_temp = torch.empty((T,), dtype=torch.float32, device=dev)
        wkey = ("ws", None if ws is None else tuple(np.asarray(ws, dtype=np.float32).tolist()),
                None if ws_max is None else tuple(np.asarray(ws_max, dtype=np.float32).tolist()), str(dev))
        if wkey not in cache:
            cache[wkey] = (torch.as_tensor(np.asarray(ws, dtype=np.float32), device=dev) if ws is not None else None,
                           torch.as_tensor(np.asarray(ws_max, dtype=np.float32), device=dev) if ws_max is not None else None)
        ws_t, wsm_t = cache[wkey]
        if halo_next is not None:
            halo_next = halo_next.contiguous()
        self._check(self.lib.dgp_potentials(self.h, _ptr(mu), _ptr(halo_next), T, nj, _ptr(edges_t), nl, _ptr(ws_t), _ptr(wsm_t),
                                            float(wt_max), _ptr(skel), _ptr(temporal), _ptr(e_skel), _ptr(e_temp), _stream(dev)))
        n_t = T if halo_next is not None else T - 1
        return {"skel": skel, "temporal": temporal[:n_t], "e_skel": e_skel, "e_temp": e_temp, "_temporal_full": temporal,
                "_T": T, "_nl": nl}

    def estimate_pose_host(self, frames_host, batch=16, gamma=1.0, gauss_len=1.0):
        """End-to-end with HOST buffers (H2D + forward + soft-argmax + D2H inside). frames_host: uint8 (T,H,W,3) CPU tensor."""
        if frames_host.is_cuda or frames_host.dtype != torch.uint8:
            raise ValueError("frames_host must be a uint8 CPU tensor (pinned for async copies)")
        frames_host = frames_host.contiguous()
        T, H, W, _ = frames_host.shape
        mu = torch.empty((T, self.nj, 2), dtype=torch.float32).pin_memory()
        peak = torch.empty((T, self.nj, 2), dtype=torch.int32).pin_memory()
        lik = torch.empty((T, self.nj), dtype=torch.float32).pin_memory()
        self._check(self.lib.dgp_estimate_pose_host(self.h, _ptr(frames_host), T, H, W, int(batch), float(gamma), float(gauss_len),
                                                    _ptr(mu), _ptr(peak), _ptr(lik)))
        return mu, peak, lik

    def estimate_pose_stream(self, source, H, W, max_frames, batch=16, gamma=1.0, gauss_len=1.0, start=0):
        """Streaming estimate_pose (dgp_estimate_pose_stream): host memory is bounded by a ring of 3 pinned slots of `batch`
        frames, whatever the length of the video.  ``source`` is either

        * an iterator / generator yielding uint8 (H,W,3) RGB frames (a video decoder): frames are written into the pinned
          slot by a reader thread while the GPU works on the previous slots, or
        * a uint8 CPU tensor (P,H,W,3) (pinned for asynchronous copies): a video of ``max_frames`` frames that cycles over
          these P frames, served zero-copy by the library's own ``dgp_cyclic_reader`` (synthetic videos, bench.py);
          ``start`` is the index of the first frame to serve (frame t of the video is ``source[t % P]``), which lets every
          rank of a sharded run read its own contiguous range of the same video.

        Returns (mu (T,nj,2) f32, peak (T,nj,2) i32, lik (T,nj) f32) as CPU tensors, T = frames actually delivered."""
        T = int(max_frames)
        mu = torch.empty((T, self.nj, 2), dtype=torch.float32).pin_memory()
        peak = torch.empty((T, self.nj, 2), dtype=torch.int32).pin_memory()
        lik = torch.empty((T, self.nj), dtype=torch.float32).pin_memory()
        done = C.c_int64(0)
        frame_bytes = int(H) * int(W) * 3
        if isinstance(source, torch.Tensor):
            if source.is_cuda or source.dtype != torch.uint8 or source.dim() != 4 or tuple(source.shape[1:]) != (H, W, 3):
                raise ValueError("a tensor source must be a uint8 CPU tensor (P,%d,%d,3)" % (H, W))
            pool = source.contiguous()
            src = _lib.DgpCyclicSource(pool.data_ptr(), pool.shape[0], int(start) + T, int(start), frame_bytes)
            reader, user, keep = C.cast(self.lib.dgp_cyclic_reader, C.c_void_p), C.cast(C.pointer(src), C.c_void_p), (pool, src)
        else:
            it = iter(source)
            state = {"error": None}

            def _read(user_, slot, want, direct):
                try:
                    dst = np.ctypeslib.as_array(C.cast(slot, C.POINTER(C.c_uint8)), shape=(want, H, W, 3))
                    n = 0
                    for n in range(1, want + 1):
                        try:
                            fr = next(it)
                        except StopIteration:
                            n -= 1
                            break
                        fr = np.asarray(fr)
                        if fr.shape != (H, W, 3) or fr.dtype != np.uint8:
                            raise ValueError("the source must yield uint8 frames of shape (%d, %d, 3), got %s %s"
                                             % (H, W, fr.dtype, fr.shape))
                        dst[n - 1] = fr
                    return n
                except Exception as ex:  # never unwind through the C frame
                    state["error"] = ex
                    return -1

            cb = _lib.FRAME_READER(_read)
            reader, user, keep = C.cast(cb, C.c_void_p), C.c_void_p(0), (cb, state)
        status = self.lib.dgp_estimate_pose_stream(self.h, reader, user, int(H), int(W), int(batch), float(gamma),
                                                   float(gauss_len), T, _ptr(mu), _ptr(peak), _ptr(lik), C.byref(done))
        if not isinstance(source, torch.Tensor) and keep[1]["error"] is not None:
            raise keep[1]["error"]
        self._check(status)
        n = int(done.value)
        return mu[:n], peak[:n], lik[:n]

    # ------------------------------------------------------------------ training (fit_dgp's train_op)
    def train_enable(self):
        if not getattr(self, "_train", False):
            self._check(self.lib.dgp_train_enable(self.h))
            self._train = True
            if self.precision == "fp16":
                # keeps the smallest activation gradients out of fp16's subnormal range; a power of two, divided out again by
                # dgp_optimizer_step, so the update is unchanged wherever nothing underflowed
                self._check(self.lib.dgp_train_set_loss_scale(self.h, 1024.0))

    def optimizer_step(self, lr=0.005, momentum=0.9, clip_norm=10.0, grad_scale=1.0):
        """clip_by_global_norm(clip_norm) + MomentumOptimizer(lr, momentum) (fitdgp.py:706-713) on the gradient buffer."""
        self._check(self.lib.dgp_optimizer_step(self.h, float(lr), float(momentum), float(clip_norm), float(grad_scale),
                                                _stream(self.device)))

    def grad_buffer(self):
        """The flat float32 gradient buffer as a CUDA tensor VIEW (what a data-parallel caller all-reduces)."""
        self.train_enable()
        p, n = C.c_void_p(), C.c_size_t()
        self._check(self.lib.dgp_get_grad_buffer(self.h, C.byref(p), C.byref(n)))
        return _device_view(p.value, (n.value // 4,), "<f4", self.device)

    def use_graphs(self, enable=True):
        """Replay the network backward as a CUDA graph from the second step on (default) or launch it eagerly."""
        self.train_enable()
        self._check(self.lib.dgp_train_use_graphs(self.h, int(enable)))

    def set_loss_scale(self, loss_scale):
        """Loss scaling for fp16-storage training (no effect on the update; see dgp_train_set_loss_scale)."""
        self.train_enable()
        self._check(self.lib.dgp_train_set_loss_scale(self.h, float(loss_scale)))

    def early_bucket(self):
        """(offset, count) in floats of the gradient-buffer slice that is final early in the backward pass (block4 + heads)."""
        self.train_enable()
        o, c = C.c_size_t(), C.c_size_t()
        self._check(self.lib.dgp_train_early_bucket(self.h, C.byref(o), C.byref(c)))
        return int(o.value), int(c.value)

    def wait_early_bucket(self, stream):
        """Make the torch.cuda.Stream `stream` wait until the early bucket of the last training step is final."""
        self._check(self.lib.dgp_train_wait_early_bucket(self.h, C.c_void_p(stream.cuda_stream)))

    def allreduce_exposed_ms(self):
        """Milliseconds between the end of the last backward pass and the end of its C-side gradient all-reduce."""
        v = C.c_float()
        self._check(self.lib.dgp_allreduce_exposed_ms(self.h, C.byref(v)))
        return float(v.value)

    def grad_norm(self):
        v = C.c_float()
        self._check(self.lib.dgp_get_grad_norm(self.h, C.byref(v)))
        return float(v.value)

    def train_outputs(self, nt, H, W):
        """Views of the head outputs (logits, locref) the last training step at this input shape produced."""
        a, b = C.c_void_p(), C.c_void_p()
        self._check(self.lib.dgp_train_outputs(self.h, int(nt), int(H), int(W), C.byref(a), C.byref(b)))
        _, (ho, wo) = output_dims(H, W)
        logits = _device_view(a.value, (nt, ho, wo, self.nj), "<f4", self.device)
        locref = _device_view(b.value, (nt, ho, wo, 2 * self.nj), "<f4", self.device) if b.value else None
        return logits, locref

    def get_variable(self, name, what="value"):
        """A trainable variable (or its gradient / momentum accumulator) under its TF name, in TF layout (float32 ndarray)."""
        code = {"value": 0, "grad": 1, "momentum": 2}[what]
        shape = (C.c_int64 * 4)()
        nd = C.c_int()
        self._check(self.lib.dgp_get_variable(self.h, name.encode(), code, None, 0, shape, C.byref(nd)))
        shp = tuple(int(shape[i]) for i in range(nd.value))
        out = np.empty(shp, np.float32)
        self._check(self.lib.dgp_get_variable(self.h, name.encode(), code, out.ctypes.data_as(C.c_void_p), out.size, shape,
                                              C.byref(nd)))
        return out

    def locref_targets(self, joint_loc, visible_frame_within_batch, nt, H, W, pos_dist_thresh=17.0, locref_stdev=7.2801):
        """Device-side coord2map + batch scatter (dataset.py:246-271, fitdgp.py:781-795): joint_loc (n_vis,nj,2) float64
        scoremap (row,col) labels (NaN = missing).  Returns float32 CUDA tensors (locref_map, locref_mask) (nt,H,W,2nj)."""
        jl = np.ascontiguousarray(np.asarray(joint_loc, dtype=np.float64).reshape(-1, self.nj, 2))
        idx = np.ascontiguousarray(np.asarray(visible_frame_within_batch, dtype=np.int32).reshape(-1))
        if jl.shape[0] != idx.shape[0]:
            raise ValueError("one batch position per labelled frame is required")
        dev = self.device
        jl_d = torch.from_numpy(jl).to(dev) if jl.size else None
        idx_d = torch.from_numpy(idx).to(dev) if idx.size else None
        lmap = torch.empty((nt, H, W, 2 * self.nj), dtype=torch.float32, device=dev)
        lmask = torch.empty_like(lmap)
        self._check(self.lib.dgp_locref_targets(self.h, _ptr(jl_d), _ptr(idx_d), int(idx.shape[0]), int(nt), int(H), int(W),
                                                float(pos_dist_thresh), float(locref_stdev), _ptr(lmap), _ptr(lmask),
                                                _stream(dev)))
        return lmap, lmask

    def set_variable(self, name, array, what="value"):
        """Overwrite a trainable variable ("value") or its Momentum accumulator ("momentum") from a TF-layout array."""
        code = {"value": 0, "momentum": 2}[what]
        if code == 2:
            self.train_enable()
        a = np.ascontiguousarray(array, dtype=np.float32)
        self._check(self.lib.dgp_set_variable(self.h, name.encode(), code, a.ctypes.data_as(C.c_void_p), a.size))

    def variable_names(self):
        """TF names of every trainable variable of the graph (slim resnet_v1_50 + the deconv heads)."""
        from .synthetic import resnet50_conv_specs
        names = []
        for scope, *_ in resnet50_conv_specs():
            names += [scope + "/weights", scope + "/BatchNorm/gamma", scope + "/BatchNorm/beta"]
        for head in ["part_pred"] + (["locref_pred"] if self.location_refinement else []):
            names += ["pose/%s/block4/weights" % head, "pose/%s/block4/biases" % head]
        return names

    def save_checkpoint(self, path):
        """All trainable variables and their Momentum accumulators under TF names -> .npz (Saver.save, fitdgp.py:830-839)."""
        out = {}
        for n in self.variable_names():
            out[n] = self.get_variable(n)
            if getattr(self, "_train", False):
                out["momentum::" + n] = self.get_variable(n, "momentum")
        np.savez(path, **out)

    def save_tf_checkpoint(self, prefix, with_momentum=True, global_step=None):
        """``saver.save(sess, prefix)`` (fitdgp.py:830-839) as a TensorFlow checkpoint bundle: every trainable variable, the
        frozen moving statistics and (like the reference's full Saver) the MomentumOptimizer slots ``<var>/Momentum``."""
        from . import tf_checkpoint
        out = dict(getattr(self, "_frozen", {}))
        for n in self.variable_names():
            out[n] = self.get_variable(n)
            if with_momentum and getattr(self, "_train", False):
                out[n + "/Momentum"] = self.get_variable(n, "momentum")
        if global_step is not None:
            out["global_step"] = np.array(int(global_step), dtype=np.int64)
        tf_checkpoint.write_checkpoint(prefix, out)

    def load_checkpoint(self, path):
        """Resume from save_checkpoint (the frozen moving statistics stay as loaded by load_weights)."""
        with np.load(path) as z:
            for k in z.files:
                if k.startswith("momentum::"):
                    self.set_variable(k[len("momentum::"):], z[k], "momentum")
                else:
                    self.set_variable(k, z[k], "value")

    # ------------------------------------------------------------------ test hooks
    def keep_activations(self, enable=True):
        self._check(self.lib.dgp_debug_keep_activations(self.h, int(enable)))

    def get_activation(self, end_point):
        shape = (C.c_int64 * 4)()
        self._check(self.lib.dgp_debug_get_activation(self.h, end_point.encode(), None, 0, shape))
        n = int(np.prod(list(shape)))
        out = np.empty(n, np.float32)
        self._check(self.lib.dgp_debug_get_activation(self.h, end_point.encode(), out.ctypes.data_as(C.c_void_p), n, shape))
        return out.reshape(tuple(shape))

    def conv2d(self, x, w_hwio, stride=1, dilation=1, pad_mode=0, scale=None, shift=None, residual=None, res_sub=1,
               relu=False, out_f32=False, block_n=0):
        """One conv through the tcgen05 implicit-GEMM kernel. x: bf16 cuda NHWC; w_hwio: float32 ndarray HWIO."""
        x = x.contiguous()
        N, H, W, Cin = x.shape
        w = np.ascontiguousarray(w_hwio, dtype=np.float32)
        R, S, _, Cout = w.shape
        if pad_mode == 2:
            Ho = (H - ((R - 1) * dilation + 1)) // stride + 1
            Wo = (W - ((S - 1) * dilation + 1)) // stride + 1
        else:
            Ho, Wo = -(-H // stride), -(-W // stride)
        if x.dtype != self.act_dtype:
            raise ValueError("x must be %s for this engine" % self.act_dtype)
        out = torch.empty((N, Ho, Wo, Cout), dtype=torch.float32 if out_f32 else self.act_dtype, device=x.device)
        sc = np.ascontiguousarray(scale, dtype=np.float32) if scale is not None else None
        sh = np.ascontiguousarray(shift, dtype=np.float32) if shift is not None else None
        res_H = res_W = 0
        if residual is not None:
            residual = residual.contiguous()
            res_H, res_W = residual.shape[1], residual.shape[2]
        self._check(self.lib.dgp_conv2d(
            self.h, _ptr(x), N, H, W, Cin, w.ctypes.data_as(C.c_void_p), R, S, Cout, stride, dilation, pad_mode,
            sc.ctypes.data_as(C.c_void_p) if sc is not None else None,
            sh.ctypes.data_as(C.c_void_p) if sh is not None else None,
            _ptr(residual), res_sub, res_H, res_W, int(relu), _ptr(out), int(out_f32), int(block_n), _stream(x.device)))
        return out

    def conv2d_wgrad(self, x, dy, R, stride=1, dilation=1, pad_mode=0, dbg=None):
        """Weight gradient of one conv through the tcgen05 wgrad GEMM: x (N,H,W,Cin), dy (N,P,Q,Cout) 16-bit NHWC ->
        float32 (Cout, R*R*Cin) in the kernel's weight layout (tap-major, then input channel)."""
        x = x.contiguous()
        dy = dy.contiguous()
        N, H, W, Cin = x.shape
        Cout = dy.shape[-1]
        if x.dtype != self.act_dtype or dy.dtype != self.act_dtype:
            raise ValueError("x and dy must be %s for this engine" % self.act_dtype)
        dw = torch.empty((Cout, R * R * Cin), dtype=torch.float32, device=x.device)
        d3 = (C.c_int32 * 3)(*[int(v) for v in dbg]) if dbg is not None else None
        self._check(self.lib.dgp_conv2d_wgrad(self.h, _ptr(x), N, H, W, Cin, _ptr(dy), R, R, Cout, stride, dilation, pad_mode,
                                              _ptr(dw), d3, _stream(x.device)))
        return dw

    def marker_indices(self, visible_frames, hidden_frames, joint_loc, nt):
        """gen_idx_chunk on the device (dgp_marker_indices).  Returns int32 CUDA tensors (visible_marker, hidden_marker,
        visible_marker_in_targets), already cut to their lengths (one 8-byte read-back of the two counts)."""
        dev = self.device
        vf = torch.as_tensor(np.asarray(visible_frames, dtype=np.int32)).to(dev)
        hf = torch.as_tensor(np.asarray(hidden_frames, dtype=np.int32)).to(dev)
        jl = torch.as_tensor(np.ascontiguousarray(np.asarray(joint_loc, dtype=np.float64).reshape(-1, self.nj, 2))).to(dev)
        cap = max(int(nt) * self.nj, 1)
        vm = torch.empty(cap, dtype=torch.int32, device=dev)
        hm = torch.empty(cap, dtype=torch.int32, device=dev)
        vit = torch.empty(cap, dtype=torch.int32, device=dev)
        cnt = torch.zeros(2, dtype=torch.int32, device=dev)
        self._check(self.lib.dgp_marker_indices(self.h, _ptr(vf) if vf.numel() else None, vf.numel(), _ptr(hf) if hf.numel() else None,
                                                hf.numel(), _ptr(jl) if jl.numel() else None, int(nt), _ptr(vm), _ptr(hm), _ptr(vit),
                                                _ptr(cnt), _stream(dev)))
        nv, nh = [int(x) for x in cnt.cpu().tolist()]
        return vm[:nv], hm[:nh], vit[:nv]

    def side_stream(self):
        """A second stream on the engine's device for feeders that may run beside the training step (learn_wt)."""
        if getattr(self, "_side_stream", None) is None:
            self._side_stream = torch.cuda.Stream(device=self.device)
        return self._side_stream

    def learn_wt(self, frames):
        """Farneback flow magnitude |u| + |v| per consecutive frame pair (dgp_learn_wt): frames uint8 cuda (T,H,W,3) ->
        float32 cuda (T-1,H,W), the `vector_field_tf` feed of the temporal clique."""
        if frames.dtype != torch.uint8 or not frames.is_cuda or frames.dim() != 4 or frames.shape[-1] != 3:
            raise ValueError("frames must be a uint8 CUDA tensor (T,H,W,3)")
        frames = frames.contiguous()
        T, H, W, _ = frames.shape
        out = torch.empty((max(T - 1, 0), H, W), dtype=torch.float32, device=frames.device)
        if T > 1:
            self._check(self.lib.dgp_learn_wt(self.h, _ptr(frames), T, H, W, _ptr(out), _stream(frames.device)))
        return out

    def motion_energy_sums(self, frames):
        """Exact byte sums of (frames[t] - frames[t-1]) mod 256 per frame (dgp_motion_energy): frames uint8 cuda (T,...)."""
        if frames.dtype != torch.uint8 or not frames.is_cuda:
            raise ValueError("frames must be a uint8 CUDA tensor")
        frames = frames.contiguous()
        T = frames.shape[0]
        sums = torch.zeros((T,), dtype=torch.int64, device=frames.device)
        if T:
            self._check(self.lib.dgp_motion_energy(self.h, _ptr(frames), T, frames[0].numel(), _ptr(sums), _stream(frames.device)))
        return sums

    PROFILE_KINDS = ("prep_s2d", "conv_gemm", "maxpool", "deconv_col2im", "softargmax", "dgrad_gemm", "wgrad_gemm",
                     "bwd_bandwidth")

    def set_profiling(self, enable=True):
        """True / 1: CUDA events around every launch; 2: one event pair around every run of same-kind launches."""
        self._check(self.lib.dgp_set_profiling(self.h, int(enable)))

    def get_profile(self):
        """{kind: (total_ms, launches)} since the last call (synchronises the device)."""
        n = len(self.PROFILE_KINDS)
        ms = (C.c_double * n)()
        cnt = (C.c_int64 * n)()
        self._check(self.lib.dgp_get_profile(self.h, ms, cnt, n))
        return {k: (ms[i], int(cnt[i])) for i, k in enumerate(self.PROFILE_KINDS)}

    def get_profile_records(self, max_records=65536):
        """[(kind, ms)] per launch, in launch order, since the last call (synchronises the device)."""
        ms = (C.c_float * max_records)()
        kind = (C.c_int32 * max_records)()
        n = C.c_int(0)
        self._check(self.lib.dgp_get_profile_records(self.h, ms, kind, max_records, C.byref(n)))
        return [(self.PROFILE_KINDS[kind[i]] if 0 <= kind[i] < len(self.PROFILE_KINDS) else str(kind[i]), ms[i]) for i in range(n.value)]

    def launch_count(self):
        return int(self.lib.dgp_launch_count(self.h))

    def num_sms(self):
        return int(self.lib.dgp_num_sms(self.h))
