"""A minimal TF1-API emulation (lazy graph over torch CPU tensors) -- TEST INFRASTRUCTURE ONLY.

Purpose: execute the reference's OWN, UNMODIFIED Python for the hot path
(/root/reference/src/deepgraphpose/models/{fitdgp_util,fitdgp,eval}.py and
/root/reference/src/DeepLabCut/deeplabcut/pose_estimation_tensorflow/nnet/{pose_net,losses,predict}.py)
in this container, where TensorFlow 1.x cannot be installed, so that golden vectors can be generated from the
reference's graph-building code itself (tests/golden/make_golden.py) and the oracle's restatement of the
*composition* of TF ops is pinned against them.

What is emulated: the few dozen ``tf.*`` / ``tf.compat.v1.*`` / ``slim.*`` symbols those files touch, each with its
TF 1.15 semantics restated on torch CPU ops (the single-op semantics live in oracle/tf_ops.py and are covered by the
known-answer tests in tests/test_oracle_tf_ops.py).  slim's ``resnet_v1_50`` -- third-party code that is NOT in the
reference tree -- is provided by oracle/resnet_v1.py.

``install()`` registers fake ``tensorflow`` modules plus permissive dummies for the unrelated imports of the reference
files (imgaug, moviepy, easydict, matplotlib, skimage, h5py ...) and returns a handle to the variable store.
Nothing here is imported by the product or by the GPU tests.
"""
import contextlib
import importlib
import importlib.abc
import importlib.machinery
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

from . import resnet_v1 as _oracle_resnet
from . import tf_ops as _tfops

REF_ROOT = "/root/reference"
float32, int32, int64, float64 = torch.float32, torch.int32, torch.int64, torch.float64
newaxis = None
_DT = {"float32": float32, "int32": int32, "int64": int64, "float64": float64, float: float32, int: int32}

VARIABLES = {}      # name -> torch tensor (the "checkpoint")
_SCOPE = []         # variable_scope stack


# ------------------------------------------------------------------------------------------- lazy graph
class Node:
    def __init__(self, fn, inputs=(), name=None, static_shape=None):
        self.fn, self.inputs, self.name, self._static = fn, tuple(inputs), name, static_shape

    # tf.Tensor surface used by the reference
    def get_shape(self):
        return _Shape(self._static)

    @property
    def shape(self):
        return _Shape(self._static)

    def __add__(self, o): return _op(lambda a, b: a + b, self, o)
    def __radd__(self, o): return _op(lambda a, b: b + a, self, o)
    def __sub__(self, o): return _op(lambda a, b: a - b, self, o)
    def __rsub__(self, o): return _op(lambda a, b: b - a, self, o)
    def __mul__(self, o): return _op(lambda a, b: a * b, self, o)
    def __rmul__(self, o): return _op(lambda a, b: b * a, self, o)
    def __truediv__(self, o): return _op(_div, self, o)
    def __rtruediv__(self, o): return _op(lambda a, b: _div(b, a), self, o)
    def __neg__(self): return _op(lambda a: -a, self)
    def __lt__(self, o): return _op(lambda a, b: a < b, self, o)
    def __gt__(self, o): return _op(lambda a, b: a > b, self, o)
    def __getitem__(self, idx): return _op(lambda a: a[_fix_index(idx)], self)
    def __hash__(self): return id(self)


class _Shape:
    def __init__(self, s): self.s = s
    def as_list(self): return list(self.s) if self.s is not None else None
    def __len__(self): return len(self.s)
    def assert_is_compatible_with(self, other): return True


def _fix_index(idx):
    return idx


def _div(a, b):
    if isinstance(a, torch.Tensor) and not a.is_floating_point() or isinstance(a, int) and isinstance(b, int):
        return a / b
    return a / b


def _t(x, like=None):
    """python / numpy / torch -> torch value (scalars stay python numbers so that dtype promotion follows the tensor)."""
    if isinstance(x, torch.Tensor):
        return x
    if isinstance(x, np.ndarray):
        t = torch.from_numpy(np.ascontiguousarray(x))
        return t.float() if t.dtype == torch.float64 else t
    if isinstance(x, (list, tuple)):
        return torch.as_tensor(np.asarray(x))
    return x


def _op(fn, *args, static_shape=None):
    return Node(lambda *vals: fn(*vals), args, static_shape=static_shape)


def evaluate(x, feed, memo):
    if isinstance(x, Node):
        if x in memo:
            return memo[x]
        if x in feed:
            v = feed[x]
        else:
            v = x.fn(*[evaluate(i, feed, memo) for i in x.inputs])
        memo[x] = v
        return v
    if isinstance(x, (list, tuple)):
        return type(x)(evaluate(i, feed, memo) for i in x)
    if isinstance(x, dict):
        return {k: evaluate(v, feed, memo) for k, v in x.items()}
    return _t(x)


def _static(x):
    if isinstance(x, Node):
        return x._static
    if hasattr(x, "shape"):
        return tuple(x.shape)
    return None


# ------------------------------------------------------------------------------------------- tf.* ops
def placeholder(dtype, shape=None, name=None):
    n = Node(None, (), name=name, static_shape=tuple(shape) if shape is not None else None)
    n.dtype = dtype
    n.is_placeholder = True
    return n


def constant(value, dtype=None, shape=None, name=None):
    v = torch.as_tensor(np.asarray(value), dtype=_DT.get(dtype, dtype) if dtype is not None else None)
    if v.dtype == torch.float64:
        v = v.float()
    if shape is not None:
        v = v.reshape(tuple(shape)) if v.numel() == int(np.prod(shape)) else v.expand(tuple(shape)).clone()
    return Node(lambda: v, (), name=name, static_shape=tuple(v.shape))


def _as_int(v):
    return int(v.item()) if isinstance(v, torch.Tensor) else int(v)


def _tensor(a):
    if isinstance(a, torch.Tensor):
        return a
    if isinstance(a, (list, tuple)):
        if any(isinstance(v, torch.Tensor) for v in a):
            return torch.stack([torch.as_tensor(v) for v in a])
        t = torch.as_tensor(np.asarray(a))
        return t.float() if t.dtype == torch.float64 else t
    t = torch.as_tensor(a)
    return t.float() if t.dtype == torch.float64 else t


def shape(x):
    return _op(lambda a: torch.tensor(list(a.shape), dtype=torch.int32), x)


def reshape(x, shp):
    def f(a, s):
        if isinstance(s, torch.Tensor):
            s = s.tolist()
        return _tensor(a).reshape([_as_int(v) for v in s])
    st = None
    if isinstance(shp, (list, tuple)) and all(isinstance(v, int) for v in shp):
        st = tuple(shp)
    return _op(f, x, list(shp) if isinstance(shp, tuple) else shp, static_shape=st)


def transpose(x, perm=None):
    st = _static(x)
    return _op(lambda a: a.permute(*perm) if perm is not None else a.permute(*reversed(range(a.dim()))), x,
               static_shape=tuple(st[p] for p in perm) if (st is not None and perm is not None) else (tuple(reversed(st)) if st else None))


def cast(x, dtype):
    return _op(lambda a: _tensor(a).to(_DT.get(dtype, dtype)), x, static_shape=_static(x))


def to_int32(x): return cast(x, int32)
def to_float(x): return cast(x, float32)


def range_(start, limit=None, delta=1, dtype=None):
    def f(a, b):
        if b is None:
            a, b = 0, a
        r = torch.arange(_as_int(a), _as_int(b), delta)
        return r.to(_DT.get(dtype, dtype)) if dtype is not None else r.to(torch.int32)
    return _op(f, start, limit)


def exp(x): return _op(torch.exp, x, static_shape=_static(x))
def log(x): return _op(torch.log, x, static_shape=_static(x))
def sqrt(x): return _op(torch.sqrt, x, static_shape=_static(x))
def square(x): return _op(torch.square, x, static_shape=_static(x))
def sigmoid(x): return _op(torch.sigmoid, x, static_shape=_static(x))
def sign(x): return _op(torch.sign, x, static_shape=_static(x))
def abs_(x): return _op(torch.abs, x, static_shape=_static(x))
def is_nan(x): return _op(torch.isnan, x, static_shape=_static(x))
def ones_like(x): return _op(torch.ones_like, x, static_shape=_static(x))
def zeros_like(x): return _op(torch.zeros_like, x, static_shape=_static(x))
def multiply(a, b): return _op(lambda u, v: u * v, a, b)
def subtract(a, b): return _op(lambda u, v: u - v, a, b)
def maximum(a, b): return _op(lambda u, v: torch.maximum(torch.as_tensor(u, dtype=torch.float32), torch.as_tensor(v, dtype=torch.float32)), a, b)
def minimum(a, b): return _op(lambda u, v: torch.minimum(torch.as_tensor(u, dtype=torch.float32), torch.as_tensor(v, dtype=torch.float32)), a, b)
def relu(x): return _op(torch.relu, x, static_shape=_static(x))
def where(c, a, b): return _op(lambda cc, u, v: torch.where(cc, u, v), c, a, b)


def _axes(axis):
    if axis is None:
        return None
    if isinstance(axis, (list, tuple)):
        return [int(a) for a in axis]
    return [int(axis)]


def _reduce(fn_all, fn_dim):
    def red(x, axis=None, keepdims=False, reduction_indices=None, **kw):
        ax = _axes(axis if axis is not None else reduction_indices)
        def f(a):
            if ax is None:
                return fn_all(a)
            out = a
            for d in sorted([d % a.dim() for d in ax], reverse=True):
                out = fn_dim(out, d)
            return out
        return _op(f, x)
    return red


reduce_sum = _reduce(lambda a: a.sum(), lambda a, d: a.sum(dim=d))
reduce_max = _reduce(lambda a: a.max(), lambda a, d: a.max(dim=d).values)
reduce_min = _reduce(lambda a: a.min(), lambda a, d: a.min(dim=d).values)
reduce_mean = _reduce(lambda a: a.mean(), lambda a, d: a.mean(dim=d))


def expand_dims(x, axis):
    ax = axis[0] if isinstance(axis, (list, tuple)) else axis
    return _op(lambda a: a.unsqueeze(ax), x)


def squeeze(x, axis=None):
    return _op(lambda a: a.squeeze() if axis is None else a.squeeze(axis), x)


def tile(x, multiples):
    return _op(lambda a, m: a.repeat(*[_as_int(v) for v in (m.tolist() if isinstance(m, torch.Tensor) else m)]), x, multiples)


def zeros(shp, dtype=float32):
    return _op(lambda s: torch.zeros([_as_int(v) for v in s], dtype=_DT.get(dtype, dtype)), list(shp))


def ones(shp, dtype=float32):
    return _op(lambda s: torch.ones([_as_int(v) for v in s], dtype=_DT.get(dtype, dtype)), list(shp))


def eye(n, batch_shape=None, dtype=float32):
    e = torch.eye(int(n))
    if batch_shape:
        e = e.reshape(*([1] * len(batch_shape)), int(n), int(n)).expand(*batch_shape, int(n), int(n)).clone()
    return constant(e.numpy())


def concat(values, axis):
    return _op(lambda vs: torch.cat([torch.as_tensor(v) for v in vs], dim=axis), list(values))


def stack(values, axis=0):
    return _op(lambda vs: torch.stack([torch.as_tensor(v) for v in vs], dim=axis), list(values))


def pad(x, paddings, mode="CONSTANT"):
    def f(a, p):
        p = p.tolist() if isinstance(p, torch.Tensor) else p
        flat = []
        for lo, hi in reversed(p):
            flat += [int(lo), int(hi)]
        return F.pad(a, flat)
    return _op(f, x, paddings)


def gather(params, indices, axis=0):
    return _op(lambda a, i: a.index_select(axis, torch.as_tensor(i).long().reshape(-1)).reshape(
        a.shape[:axis] + tuple(torch.as_tensor(i).shape) + a.shape[axis + 1:]), params, indices)


def gather_nd(params, indices):
    def f(a, i):
        i = i.long()
        return a[tuple(i[..., k] for k in range(i.shape[-1]))]
    return _op(f, params, indices)


def scatter_nd(indices, updates, shp):
    def f(i, u, s):
        n = _as_int(s[0] if not isinstance(s, int) else s)
        out = torch.zeros((n,) + tuple(u.shape[1:]), dtype=u.dtype)
        return out.index_add(0, i.long().reshape(-1), u)   # duplicate indices accumulate, as in TF
    return _op(f, indices, updates, shp)


def matmul(a, b): return _op(lambda u, v: u @ v, a, b)
def norm(x, ord=2): return _op(lambda a: torch.sqrt(torch.sum(a * a)), x)
def argmax(x, axis=0): return _op(lambda a: torch.argmax(a, dim=axis), x)


def unravel_index(indices, dims):
    def f(i, d):
        d = [_as_int(v) for v in d]
        rows = torch.div(i, d[1], rounding_mode="floor")
        return torch.stack([rows, i - rows * d[1]])
    return _op(f, indices, list(dims))


def softmax(x, axis=-1): return _op(lambda a: torch.softmax(a, dim=axis), x, static_shape=_static(x))


def separable_conv2d(x, depthwise_filter, pointwise_filter, strides, padding):
    def f(a, dw, pw):
        C = a.shape[-1]
        w = dw.permute(2, 3, 0, 1).reshape(C, 1, dw.shape[0], dw.shape[1])
        y = F.conv2d(a.permute(0, 3, 1, 2), w, groups=C)          # VALID
        y = torch.einsum("nchw,cd->ndhw", y, pw.reshape(pw.shape[-2], pw.shape[-1]))
        return y.permute(0, 2, 3, 1)
    assert padding == "VALID"
    return _op(f, x, depthwise_filter, pointwise_filter)


def crop_and_resize(image, boxes, box_ind, crop_size):
    from .dgp_loss import crop_and_resize_mean  # noqa: F401  (same bilinear rule; full crops are materialised here)
    def f(img, bx, bi, cs):
        ch, cw = _as_int(cs[0]), _as_int(cs[1])
        B, H, W, C = img.shape
        outs = []
        for k in range(bx.shape[0]):
            y1, x1, y2, x2 = [bx[k, i] for i in range(4)]
            im = img[int(bi[k])]
            ys = y1 * (H - 1) + torch.arange(ch, dtype=img.dtype) * ((y2 - y1) * (H - 1) / max(ch - 1, 1))
            xs = x1 * (W - 1) + torch.arange(cw, dtype=img.dtype) * ((x2 - x1) * (W - 1) / max(cw - 1, 1))
            vy, vx = (ys >= 0) & (ys <= H - 1), (xs >= 0) & (xs <= W - 1)
            y0, y1i = torch.floor(ys).clamp(0, H - 1).long(), torch.ceil(ys).clamp(0, H - 1).long()
            x0, x1i = torch.floor(xs).clamp(0, W - 1).long(), torch.ceil(xs).clamp(0, W - 1).long()
            ly, lx = (ys - torch.floor(ys))[:, None, None], (xs - torch.floor(xs))[None, :, None]
            tl, tr, bl, br = im[y0][:, x0], im[y0][:, x1i], im[y1i][:, x0], im[y1i][:, x1i]
            top, bot = tl + (tr - tl) * lx, bl + (br - bl) * lx
            outs.append((top + (bot - top) * ly) * (vy[:, None] & vx[None, :]).to(img.dtype)[..., None])
        return torch.stack(outs)
    return _op(f, image, boxes, box_ind, list(crop_size))


# ------------------------------------------------------------------------------------------- losses / slim / resnet
def sigmoid_cross_entropy(multi_class_labels, logits, weights=1.0, **kw):
    return _op(lambda z, x, w: _tfops.compute_weighted_loss(_tfops.sigmoid_cross_entropy_with_logits(z, x), w),
               multi_class_labels, logits, weights)


def mean_squared_error(labels, predictions, weights=1.0, **kw):
    return _op(lambda z, x, w: _tfops.compute_weighted_loss(torch.square(x - z), w), labels, predictions, weights)


def compute_weighted_loss(losses, weights=1.0, **kw):
    return _op(lambda l, w: _tfops.compute_weighted_loss(l, w), losses, weights)


@contextlib.contextmanager
def variable_scope(name, reuse=None, **kw):
    _SCOPE.append(name)
    try:
        yield
    finally:
        _SCOPE.pop()


@contextlib.contextmanager
def _null_scope(*a, **kw):
    yield


def _scoped(name):
    return "/".join(_SCOPE + [name])


class _ConstInit:
    """tf.constant_initializer(value): the variable starts as `value` (reshaped to the variable's shape)."""
    def __init__(self, value):
        self.value = np.asarray(value, dtype=np.float32)


def slim_conv2d_transpose(inputs, num_outputs, kernel_size, stride=1, scope=None, weights_initializer=None,
                          biases_initializer=None, **kw):
    full = _scoped(scope)
    assert list(kernel_size) == [3, 3] and stride == 2
    if isinstance(weights_initializer, _ConstInit):
        # dgp_prediction_layer(init_flag=True) (fitdgp_util.py:57-66): the variables are created from constants, nothing is restored
        w = torch.from_numpy(np.ascontiguousarray(weights_initializer.value))
        b = torch.from_numpy(np.ascontiguousarray(biases_initializer.value.reshape(-1)))
        assert tuple(w.shape[:3]) == (3, 3, num_outputs) and b.numel() == num_outputs
        return _op(lambda a: _tfops.conv2d_transpose_same_s2(a, w, b), inputs, static_shape=(None, None, None, num_outputs))
    return _op(lambda a: _tfops.conv2d_transpose_same_s2(a, VARIABLES[full + "/weights"], VARIABLES[full + "/biases"]), inputs,
               static_shape=(None, None, None, num_outputs))


def resnet_v1_50(inputs, num_classes=None, is_training=True, global_pool=True, output_stride=None, **kw):
    assert not global_pool and num_classes is None and is_training is False
    end_points = {}
    return _op(lambda a: _oracle_resnet.resnet_v1_50(a, VARIABLES, output_stride), inputs,
               static_shape=(None, None, None, 2048)), end_points


# ------------------------------------------------------------------------------------------- session & misc
class Session:
    def __init__(self, *a, **kw): pass
    def run(self, fetches, feed_dict=None):
        if isinstance(fetches, _Dummy) or fetches is None:
            return None
        feed = {}
        for k, v in (feed_dict or {}).items():
            if isinstance(k, Node):
                t = torch.as_tensor(np.asarray(v))
                dt = getattr(k, "dtype", None)
                feed[k] = t.to(_DT.get(dt, dt)) if dt is not None else (t.float() if t.dtype == torch.float64 else t)
        with torch.no_grad():
            out = evaluate(fetches, feed, {})
        def conv(o):
            if isinstance(o, torch.Tensor):
                return o.numpy()
            if isinstance(o, (list, tuple)):
                return type(o)(conv(i) for i in o)
            if isinstance(o, dict):
                return {k: conv(v) for k, v in o.items()}
            return o
        return conv(out)
    def close(self): pass
    def __enter__(self): return self
    def __exit__(self, *a): pass


class _Dummy:
    """Permissive stand-in for everything unrelated (Saver, ConfigProto, imgaug, moviepy ...)."""
    def __init__(self, *a, **kw): pass
    def __call__(self, *a, **kw): return _Dummy()
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Dummy()
    def __iter__(self): return iter(())
    def __enter__(self): return self
    def __exit__(self, *a): return False
    def __mro_entries__(self, bases): return (object,)


class _DummyModule(types.ModuleType):
    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Dummy()


def _ns(name, **attrs):
    m = _DummyModule(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


def _build_tf():
    api = dict(
        float32=float32, int32=int32, int64=int64, float64=float64, newaxis=None,
        placeholder=placeholder, constant=constant, shape=shape, reshape=reshape, transpose=transpose, cast=cast,
        to_int32=to_int32, to_float=to_float, range=range_, exp=exp, log=log, sqrt=sqrt, square=square, sigmoid=sigmoid,
        sign=sign, abs=abs_, is_nan=is_nan, ones_like=ones_like, zeros_like=zeros_like, multiply=multiply,
        subtract=subtract, where=where, reduce_sum=reduce_sum, reduce_max=reduce_max, reduce_min=reduce_min,
        reduce_mean=reduce_mean, expand_dims=expand_dims, squeeze=squeeze, tile=tile, zeros=zeros, ones=ones, eye=eye,
        concat=concat, stack=stack, pad=pad, gather=gather, gather_nd=gather_nd, scatter_nd=scatter_nd, matmul=matmul,
        norm=norm, argmax=argmax, unravel_index=unravel_index, variable_scope=variable_scope, Session=Session,
        AUTO_REUSE=True, reset_default_graph=lambda: None,
        global_variables_initializer=lambda: None, local_variables_initializer=lambda: None,
        constant_initializer=lambda value=0, *a, **k: _ConstInit(value),
    )
    nn = _ns("tensorflow.nn", softmax=softmax, separable_conv2d=separable_conv2d, relu=relu, sigmoid=sigmoid)
    math = _ns("tensorflow.math", multiply=multiply, maximum=maximum, minimum=minimum)
    losses = _ns("tensorflow.losses", sigmoid_cross_entropy=sigmoid_cross_entropy,
                 mean_squared_error=mean_squared_error, compute_weighted_loss=compute_weighted_loss)
    image = _ns("tensorflow.image", crop_and_resize=crop_and_resize)
    v1 = _ns("tensorflow.compat.v1", nn=nn, math=math, losses=losses, image=image, **api)
    compat = _ns("tensorflow.compat", v1=v1)
    tf = _ns("tensorflow", __version__="1.15.0", compat=compat, nn=nn, math=math, losses=losses, image=image, **api)
    resnet = _ns("tensorflow.contrib.slim.nets.resnet_v1", resnet_v1_50=resnet_v1_50, resnet_arg_scope=lambda *a, **k: {})
    nets = _ns("tensorflow.contrib.slim.nets", resnet_v1=resnet)
    slim = _ns("tensorflow.contrib.slim", arg_scope=_null_scope, conv2d_transpose=slim_conv2d_transpose,
               l2_regularizer=lambda *a, **k: None, nets=nets, conv2d=_Dummy(), get_variables_to_restore=lambda *a, **k: [])
    contrib = _ns("tensorflow.contrib", slim=slim)
    tf.contrib = contrib
    math_ops = _ns("tensorflow.python.ops.math_ops", to_float=to_float, subtract=subtract)
    ops_mod = _ns("tensorflow.python.framework.ops", name_scope=lambda *a, **k: _NameScope())
    py_ops = _ns("tensorflow.python.ops", math_ops=math_ops)
    py_fw = _ns("tensorflow.python.framework", ops=ops_mod)
    py = _ns("tensorflow.python", ops=py_ops, framework=py_fw)
    tf.python = py
    mods = {"tensorflow": tf, "tensorflow.compat": compat, "tensorflow.compat.v1": v1, "tensorflow.contrib": contrib,
            "tensorflow.contrib.slim": slim, "tensorflow.contrib.slim.nets": nets,
            "tensorflow.contrib.slim.nets.resnet_v1": resnet, "tensorflow.python": py, "tensorflow.python.ops": py_ops,
            "tensorflow.python.ops.math_ops": math_ops, "tensorflow.python.framework": py_fw,
            "tensorflow.python.framework.ops": ops_mod, "tensorflow.nn": nn, "tensorflow.losses": losses}
    return mods


class _NameScope:
    def __enter__(self): return "scope"
    def __exit__(self, *a): return False


class _DummyFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """Last-resort finder: unrelated third-party imports of the reference files resolve to permissive dummies."""
    PREFIXES = ("imgaug", "moviepy", "easydict", "matplotlib", "skimage", "h5py", "tensorpack", "ruamel", "wx",
                "statsmodels", "tables", "deeplabcut.utils", "deeplabcut.pose_estimation_tensorflow.train",
                "deeplabcut.pose_estimation_tensorflow.config", "deeplabcut.pose_estimation_tensorflow.dataset.factory",
                "deeplabcut.pose_estimation_tensorflow.dataset.pose_defaultdataset",
                "deeplabcut.pose_estimation_tensorflow.util", "deepgraphpose.dataset", "deepgraphpose.utils_model",
                "deepgraphpose.utils_data", "tqdm")

    def find_spec(self, fullname, path=None, target=None):
        if any(fullname == p or fullname.startswith(p + ".") for p in self.PREFIXES):
            return importlib.machinery.ModuleSpec(fullname, self, is_package=True)
        return None

    def create_module(self, spec):
        m = _DummyModule(spec.name)
        m.__path__ = []
        return m

    def exec_module(self, module):
        pass


def _pkg(name, path):
    m = types.ModuleType(name)
    m.__path__ = [path]
    return m


def install():
    """Register the fake modules; afterwards ``import deepgraphpose.models.fitdgp`` etc. load the REAL reference files."""
    if "tensorflow" in sys.modules and not isinstance(sys.modules["tensorflow"], _DummyModule):
        raise RuntimeError("a real tensorflow is importable: use it instead of the shim")
    sys.modules.update(_build_tf())
    src = REF_ROOT + "/src"
    dlc = src + "/DeepLabCut/deeplabcut"
    # package shells with the real __path__ but WITHOUT executing the reference's heavy __init__ files
    for name, path in (("deeplabcut", dlc), ("deeplabcut.pose_estimation_tensorflow", dlc + "/pose_estimation_tensorflow"),
                       ("deeplabcut.pose_estimation_tensorflow.nnet", dlc + "/pose_estimation_tensorflow/nnet"),
                       ("deeplabcut.pose_estimation_tensorflow.dataset", dlc + "/pose_estimation_tensorflow/dataset"),
                       ("deepgraphpose", src + "/deepgraphpose"), ("deepgraphpose.models", src + "/deepgraphpose/models")):
        sys.modules.setdefault(name, _pkg(name, path))
    if not any(isinstance(f, _DummyFinder) for f in sys.meta_path):
        sys.meta_path.insert(0, _DummyFinder())
    return VARIABLES


def set_variables(weights):
    VARIABLES.clear()
    for k, v in weights.items():
        VARIABLES[k] = torch.from_numpy(np.ascontiguousarray(v, dtype=np.float32))
