"""A minimal stand-in for the TensorFlow / Keras API surface the reference uses - TEST INFRASTRUCTURE.

Why it exists.  The reference (/root/reference, pure Python on TensorFlow 2.x) ships no tests or golden vectors, and
TensorFlow is neither installed nor installable in the build container.  With this module registered under the names
``tensorflow``, ``tensorflow.keras`` ... the UNMODIFIED reference files ``model_library.py`` and ``data_utils.py``
import and run here: the reference's own code decides every wiring question - which skip feeds which block, the
concatenation order, the reshape / tile / reduce axes of the filter synthesis, the two softmax axes, the tap order of
``Convolve``, the ``* initial_W`` of ``Convolve_perlayer``, the transposes in ``invert_preproc``, the crop / jitter /
noise sequence of ``preprocess_image`` - and only the PRIMITIVE ops below are ours.  ``tests/golden/make_ref_golden.py``
uses it to write ``tests/golden/ref_*.npz``; the oracle (``oracle/``) and the CUDA path are then checked against those.

What it does not prove: that the primitives below mean what TensorFlow means.  Each one states the TensorFlow
semantics it assumes (the same list as DESIGN.md section 6); with a real TensorFlow the same generator script runs
unchanged (it prefers ``import tensorflow`` when that works) and the assumption disappears.

Tensors are plain ``torch.Tensor`` (CPU, fp32/fp64); ``install()`` adds the two tensor methods the reference calls
that torch lacks (``get_shape().as_list()``).  Random ops draw from ``RNG`` - a seeded numpy generator - and append
every draw to ``DRAW_LOG`` so a test can replay the same draws through another implementation.
"""
from __future__ import annotations

import math
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

RNG = np.random.default_rng(1234)
DRAW_LOG = []            # (op name, value) in call order
LAYER_TAPS = None        # set to a dict to record every Keras layer's output by its attribute path


def reseed(seed):
    global RNG
    RNG = np.random.default_rng(seed)
    DRAW_LOG.clear()


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype if dtype is not None else None)


def _f(x):
    """Promote python scalars / numpy to a float tensor (tf converts them to float32 constants)."""
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(x, dtype=torch.float32)


class _Shape(tuple):
    def as_list(self):
        return list(self)


# ------------------------------------------------------------------ tf.* functions
def shape(x):
    return tuple(int(s) for s in x.shape)                    # eager tf.shape values are used as python ints


def reshape(x, shp):
    return x.reshape(tuple(int(s) for s in shp))


def expand_dims(x, axis):
    return x.unsqueeze(axis)


def tile(x, multiples):
    return x.repeat(*[int(m) for m in multiples])            # tf.tile: multiples[i] copies along axis i


def concat(values, axis):
    return torch.cat(list(values), dim=axis)


def stack(values, axis=0):
    return torch.stack(list(values), dim=axis)


def transpose(x, perm=None):
    if perm is None:                                         # tf.transpose default: reverse all axes
        perm = list(range(x.dim()))[::-1]
    return x.permute(*perm)


def pad(x, paddings):
    """tf.pad CONSTANT (zeros); paddings = [[before, after]] per axis."""
    flat = []
    for before, after in reversed([tuple(int(v) for v in p) for p in paddings]):
        flat += [before, after]
    return F.pad(x, flat)


def reduce_mean(x, axis=None, keepdims=False, name=None):
    return x.mean() if axis is None else x.mean(dim=axis, keepdim=keepdims)


def reduce_sum(x, axis=None, keepdims=False, name=None):
    return x.sum() if axis is None else x.sum(dim=axis, keepdim=keepdims)


def reduce_max(x, axis=None, keepdims=False):
    return x.max() if axis is None else x.amax(dim=axis, keepdim=keepdims)


def square(x):
    return x * x


def sqrt(x):
    return torch.sqrt(_f(x))


def tf_abs(x):
    return torch.abs(x)


def tf_pow(x, y):
    return torch.pow(_f(x), _f(y))


def maximum(x, y):
    return torch.maximum(_f(x), _f(y).to(_f(x).dtype) if isinstance(x, torch.Tensor) else _f(y))


def minimum(x, y):
    return torch.minimum(_f(x), _f(y).to(_f(x).dtype) if isinstance(x, torch.Tensor) else _f(y))


def where(cond, a, b):
    return torch.where(cond, _f(a), _f(b))


def log(x):
    return torch.log(_f(x))


def exp(x):
    return torch.exp(_f(x))


def cast(x, dtype):
    return _t(x).to(dtype)


def zeros(shp, dtype=torch.float32):
    return torch.zeros(tuple(int(s) for s in shp), dtype=dtype)


def ones(shp, dtype=torch.float32):
    return torch.ones(tuple(int(s) for s in shp), dtype=dtype)


def zeros_like(x):
    return torch.zeros_like(x)


def cond(pred, true_fn, false_fn):
    return true_fn() if bool(pred) else false_fn()


def convert_to_tensor(x, dtype=None):
    return _t(x, dtype)


def softmax(x, axis=-1, name=None):
    """tf.nn.softmax."""
    return torch.softmax(x, dim=axis)


# ------------------------------------------------------------------ random ops (seeded, logged)
def random_uniform(shp, minval=0.0, maxval=1.0, dtype=torch.float32, seed=None):
    shp = tuple(int(s) for s in shp)
    if dtype in (torch.int32, torch.int64):
        v = RNG.integers(int(minval), int(maxval), size=shp)
        out = torch.as_tensor(v, dtype=dtype)
    else:
        v = RNG.uniform(float(minval), float(maxval), size=shp).astype(np.float32)
        out = torch.as_tensor(v)
    DRAW_LOG.append(("random_uniform", out.clone()))
    return out


def random_normal(shp, mean=0.0, stddev=1.0, dtype=torch.float32, seed=None):
    shp = tuple(int(s) for s in shp)
    out = torch.as_tensor(RNG.standard_normal(size=shp).astype(np.float32)) * stddev + mean
    DRAW_LOG.append(("random_normal", out.clone()))
    return out


def random_poisson(lam, shp, dtype=torch.float32, seed=None):
    shp = tuple(int(s) for s in shp)
    out = torch.as_tensor(np.asarray(RNG.poisson(lam, size=shp)), dtype=dtype)
    DRAW_LOG.append(("random_poisson", out.clone()))
    return out


def random_crop(value, size, seed=None, name=None):
    """tf.image.random_crop: a uniformly random offset per axis, offset in [0, dim - size]."""
    size = [int(s) for s in size]
    off = [int(RNG.integers(0, d - s + 1)) for d, s in zip(value.shape, size)]
    DRAW_LOG.append(("random_crop", tuple(off)))
    sl = tuple(slice(o, o + s) for o, s in zip(off, size))
    return value[sl]


def resize(images, size, method="bilinear"):
    """tf.image.resize.  AREA with an integer shrink factor == the exact box mean (the only use: data_utils.py:459);
    BILINEAR == half-pixel-centre sampling (TF2 semantics)."""
    h, w = int(size[0]), int(size[1])
    x = images
    squeeze = x.dim() == 3
    if squeeze:
        x = x.unsqueeze(0)
    n, hh, ww, c = x.shape
    if method == "area":
        assert hh % h == 0 and ww % w == 0, "the stand-in implements AREA for integer factors only"
        out = x.reshape(n, h, hh // h, w, ww // w, c).mean(dim=(2, 4))
    else:
        out = F.interpolate(x.permute(0, 3, 1, 2), size=(h, w), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    return out[0] if squeeze else out


# ------------------------------------------------------------------ Keras
class Variable:
    def __init__(self, value):
        self.value = value

    def assign(self, v):
        v = _t(v, self.value.dtype)
        assert tuple(v.shape) == tuple(self.value.shape), (tuple(v.shape), tuple(self.value.shape))
        self.value = v.clone()
        return self

    def numpy(self):
        return self.value.numpy()

    @property
    def shape(self):
        return _Shape(self.value.shape)


class Layer:
    def __init__(self, name=None, **kwargs):
        self.name = name
        self._path = None

    def _children(self):
        for k, v in vars(self).items():
            if isinstance(v, Layer):
                yield k, v

    def _name_paths(self, prefix=""):
        for k, v in self._children():
            v._path = prefix + k
            v._name_paths(prefix + k + ".")

    def __call__(self, *args, **kwargs):
        if self._path is None:
            self._name_paths()
        out = self.call(*args, **kwargs)
        if LAYER_TAPS is not None and self._path is not None and isinstance(out, torch.Tensor):
            LAYER_TAPS[self._path] = out
        return out


class Model(Layer):
    pass


class Conv2D(Layer):
    """keras.layers.Conv2D, NHWC, stride 1: cross-correlation with an HWIO kernel; 'same' = symmetric zero padding of
    (k - 1) // 2 (odd k), 'valid' = none; glorot_uniform kernel / zero bias when first called."""

    def __init__(self, filters, kernel_size, padding="valid", activation=None, kernel_regularizer=None, input_shape=None,
                 name=None, **kwargs):
        super().__init__(name=name)
        self.filters = int(filters)
        self.k = int(kernel_size) if not isinstance(kernel_size, (tuple, list)) else int(kernel_size[0])
        self.padding = padding
        self.activation = activation
        self.kernel = None
        self.bias = None

    def build(self, cin, dtype):
        fan_in, fan_out = self.k * self.k * cin, self.k * self.k * self.filters
        lim = math.sqrt(6.0 / (fan_in + fan_out))
        w = RNG.uniform(-lim, lim, size=(self.k, self.k, cin, self.filters)).astype(np.float32)
        self.kernel = Variable(torch.as_tensor(w, dtype=dtype))
        self.bias = Variable(torch.zeros(self.filters, dtype=dtype))

    def call(self, x):
        if self.kernel is None:
            self.build(x.shape[-1], x.dtype)
        w = self.kernel.value.to(x.dtype).permute(3, 2, 0, 1)         # HWIO -> OIHW
        if self.padding == "same":
            assert self.k % 2 == 1
            p = self.k // 2
        else:
            p = 0
        y = F.conv2d(x.permute(0, 3, 1, 2), w, self.bias.value.to(x.dtype), padding=p).permute(0, 2, 3, 1)
        if self.activation == "relu":
            y = torch.relu(y)
        elif self.activation is not None:
            raise NotImplementedError(self.activation)
        return y


class Dense(Layer):
    def __init__(self, units, activation=None, name=None, **kwargs):
        super().__init__(name=name)


class MaxPooling2D(Layer):
    def __init__(self, pool_size=(2, 2), strides=None, padding="valid", name=None, **kwargs):
        super().__init__(name=name)
        self.pool, self.strides = tuple(pool_size), tuple(strides or pool_size)
        assert padding == "valid"

    def call(self, x):
        return F.max_pool2d(x.permute(0, 3, 1, 2), self.pool, self.strides).permute(0, 2, 3, 1)


class UpSampling2D(Layer):
    """keras.layers.UpSampling2D(interpolation='bilinear') in TF2 = tf.image.resize(BILINEAR) with half-pixel centres."""

    def __init__(self, size=(2, 2), interpolation="nearest", name=None, **kwargs):
        super().__init__(name=name)
        self.size, self.interpolation = tuple(size), interpolation

    def call(self, x):
        mode = {"bilinear": "bilinear", "nearest": "nearest"}[self.interpolation]
        kw = dict(align_corners=False) if mode == "bilinear" else {}
        return F.interpolate(x.permute(0, 3, 1, 2), scale_factor=self.size, mode=mode, **kw).permute(0, 2, 3, 1)


class GlobalAveragePooling2D(Layer):
    def call(self, x):
        return x.mean(dim=(1, 2))


def concatenate(inputs, axis=-1):
    return torch.cat(list(inputs), dim=axis)


class _Mean:
    """keras.metrics.Mean."""

    def __init__(self, name=None):
        self.total, self.count = 0.0, 0

    def __call__(self, v):
        self.total += float(v)
        self.count += 1

    def result(self):
        return self.total / max(self.count, 1)

    def reset_states(self):
        self.total, self.count = 0.0, 0


def _module(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    return m


def install():
    """Register the stand-in modules (idempotent).  Returns the fake ``tensorflow`` module."""
    if "tensorflow" in sys.modules and getattr(sys.modules["tensorflow"], "__standin__", False):
        return sys.modules["tensorflow"]
    if not hasattr(torch.Tensor, "get_shape"):
        torch.Tensor.get_shape = lambda self: _Shape(int(s) for s in self.shape)
    regularizers = _module("tensorflow.keras.regularizers", l2=lambda v: ("l2", v))
    layers = _module("tensorflow.keras.layers", Layer=Layer, Conv2D=Conv2D, Dense=Dense, MaxPooling2D=MaxPooling2D,
                     UpSampling2D=UpSampling2D, GlobalAveragePooling2D=GlobalAveragePooling2D, concatenate=concatenate)
    backend = _module("tensorflow.keras.backend", random_normal=lambda shape: random_normal(shape))
    metrics = _module("tensorflow.keras.metrics", Mean=_Mean)
    keras = _module("tensorflow.keras", layers=layers, Model=Model, regularizers=regularizers, backend=backend,
                    metrics=metrics)
    image = _module("tensorflow.image", random_crop=random_crop, resize=resize,
                    ResizeMethod=types.SimpleNamespace(AREA="area", BILINEAR="bilinear"))
    v1 = _module("tensorflow.compat.v1", random_uniform=random_uniform, random_normal=random_normal,
                 random_poisson=random_poisson, random_crop=random_crop)
    compat = _module("tensorflow.compat", v1=v1)
    data = _module("tensorflow.data", experimental=types.SimpleNamespace(AUTOTUNE=-1))
    tfmath = _module("tensorflow.math", log=log, square=square, exp=exp, reciprocal=lambda x: 1.0 / x)
    nn = _module("tensorflow.nn", softmax=softmax)
    rnd = _module("tensorflow.random", set_seed=reseed)
    tf = _module(
        "tensorflow", __standin__=True, keras=keras, image=image, compat=compat, data=data, math=tfmath, nn=nn,
        random=rnd, shape=shape, reshape=reshape, expand_dims=expand_dims, tile=tile, concat=concat, stack=stack,
        transpose=transpose, pad=pad, reduce_mean=reduce_mean, reduce_sum=reduce_sum, reduce_max=reduce_max,
        square=square, sqrt=sqrt, abs=tf_abs, pow=tf_pow, maximum=maximum, minimum=minimum, where=where, exp=exp,
        cast=cast, zeros=zeros, ones=ones, zeros_like=zeros_like, cond=cond, convert_to_tensor=convert_to_tensor,
        float32=torch.float32, float64=torch.float64, int32=torch.int32, int64=torch.int64, string=str)
    mods = {"tensorflow": tf, "tensorflow.keras": keras, "tensorflow.keras.layers": layers,
            "tensorflow.keras.regularizers": regularizers, "tensorflow.keras.backend": backend,
            "tensorflow.keras.metrics": metrics, "tensorflow.image": image, "tensorflow.compat": compat,
            "tensorflow.compat.v1": v1, "tensorflow.data": data, "tensorflow.math": tfmath, "tensorflow.nn": nn,
            "tensorflow.random": rnd}
    # the reference also imports these at module level without using them on the path
    for name in ("pydot", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                mods[name] = _module(name)
    if "matplotlib" in mods and "matplotlib.pyplot" in mods:
        mods["matplotlib"].pyplot = mods["matplotlib.pyplot"]
    sys.modules.update(mods)
    return tf
