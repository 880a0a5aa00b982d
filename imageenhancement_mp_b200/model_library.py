"""Drop-in for the model-builder entry points of the reference's ``model_library.py``.

``Simplemodel(params)`` / ``Basis_kpn(params)`` keep the reference constructors
(/root/reference/model_library.py:307-322, 180-195: a ``params`` dict with
``BURST_LENGTH, Kernel_size, Basis_num, regu, ps, layer_type``), the attributes
``burst_length, K, B`` and the call convention ``model(x)`` with ``x`` float32 NHWC
``[N,H,W,T+add]`` returning ``(output, Bas, originbasis)`` (:452) or ``(output, Bas)`` (:295).
Tensors are ``torch`` CUDA tensors; the forward runs the sm_100a kernels of csrc/ through
the C ABI - there is no CPU fallback.

Differences, all at the boundary: weights are random-initialised like Keras (glorot-uniform,
zero bias) from a seeded CPU generator, or loaded from a flat ``.npz`` (weights.py) instead
of a TF checkpoint; H and W that are not multiples of the network stride (8 / 32) are
zero-padded at the bottom/right and the output cropped - the reference raises at its first
``concatenate`` for such inputs (:96).
"""
from __future__ import annotations

import torch

from . import weights as _weights
from .engine import Engine
from ._lib import ImgEnhError


class _KpnModel:
    _arch = None
    _layers = None

    def __init__(self, params, name=None, weights=None, seed=1234, init="glorot", device="cuda", **kwargs):
        self.name = name
        self.burst_length = params["BURST_LENGTH"]          # :312 / :185
        self.K = params["Kernel_size"]                       # :313
        self.B = params["Basis_num"]                         # :316
        self.regu = params.get("regu", 0.0)
        self.ps = params.get("ps", False)
        self.params = dict(params)
        if weights is None:
            weights = _weights.init_weights(type(self)._layers(params), seed=seed, scheme=init)
        elif isinstance(weights, str):
            weights = _weights.load(weights, layers=type(self)._layers(params))   # .npz or a TensorFlow checkpoint
        self._engine = Engine(type(self)._arch, self.params, weights, device=device)

    # Keras-like conveniences
    def load_weights(self, weights):
        if isinstance(weights, str):
            weights = _weights.load(weights, layers=type(self)._layers(self.params))
        self._engine.load_weights(weights)

    @property
    def stride(self):
        return self._engine.stride

    def _forward(self, inputs, taps=None, conv_fn="ie_conv2d_nhwc_bf16"):
        if not isinstance(inputs, torch.Tensor) or not inputs.is_cuda:
            raise ImgEnhError("inputs must be a CUDA torch.Tensor [N,H,W,T+add] (no CPU fallback)")
        if inputs.dim() != 4:
            raise ImgEnhError(f"inputs must be [N,H,W,T+add], got shape {tuple(inputs.shape)}")
        if inputs.shape[0] == 0:
            # an empty batch (a rank whose shard of the last batch is empty): empty outputs, no launch - the reference's
            # Keras model returns empty tensors of the same trailing shapes
            e = self._engine
            n, hs, ws, _ = inputs.shape
            f = lambda *shape: torch.empty(shape, dtype=torch.float32, device=inputs.device)
            return f(0, hs, ws, e.T + 1), f(0, e.K, e.K, e.T, e.B), f(0, e.K, e.K, e.T * e.B)
        # sizes that are not multiples of the network stride are zero-padded implicitly by the engine
        if taps is None and conv_fn == "ie_conv2d_nhwc_bf16":
            return self._engine.forward_auto(inputs)          # CUDA-graph replay when the batch is launch-bound
        return self._engine.forward(inputs, taps=taps, conv_fn=conv_fn)

    def call(self, inputs):
        return self.__call__(inputs)

    def call_spatial_shard(self, x_slab, shard, total_rows, reduce=None):
        """One rank's part of a spatially sharded forward of ONE large image.

        ``x_slab``: rows ``shard["slab"]`` of the image(s) [N, s1-s0, W, T+add]; ``shard`` from ``dist.spatial_shards``;
        ``reduce``: SUM all-reduce over the ranks (default ``dist.all_reduce_sum``: NCCL under torchrun).
        Returns (output of the owned rows [N, b-a, W, T+1], Bas, originbasis) - Bas is identical on every rank."""
        from . import dist as _dist
        a, b = shard["own_in_slab"]
        out, bas, ob = self._engine.forward(
            x_slab, shard=dict(own=(a, b), total_rows=total_rows, reduce=reduce or _dist.all_reduce_sum))
        return out[:, a:b].contiguous(), bas, ob


class Simplemodel(_KpnModel):
    """model_library.py:306-452."""
    _arch = "simple"
    _layers = staticmethod(_weights.simplemodel_layers)

    def __init__(self, params, name='simple_kpn', **kwargs):
        super().__init__(params, name=name, **kwargs)

    def __call__(self, inputs, **kw):
        return self._forward(inputs, **kw)          # (output, Bas, originbasis)  :452


class Basis_kpn(_KpnModel):
    """model_library.py:179-295."""
    _arch = "basis_kpn"
    _layers = staticmethod(_weights.basis_kpn_layers)

    def __init__(self, params, name='basis_kpn', **kwargs):
        super().__init__(params, name=name, **kwargs)

    def __call__(self, inputs, **kw):
        out, bas, _ = self._forward(inputs, **kw)
        return out, bas                              # :295


class Convolve:
    """model_library.py:114-135: per-pixel filtering with materialised filters, summed over the frames.

    ``Convolve(K)(img_stack [N,H,W,T], filts [N,H,W,K,K,T]) -> [N,H,W]``.  The models of this package never
    materialise ``filts`` (they call the fused ``ops.kpn_apply``); this layer exists for drop-in callers."""

    def __init__(self, final_K, name='convolve', **kwargs):
        self.final_K = final_K
        self.name = name

    def __call__(self, img_stack, filts):
        from . import ops
        return ops.convolve_filts(img_stack, filts, self.final_K)[..., 0]

    call = __call__


def cus_convolve(img_stack, filts, final_K):
    """model_library.py:136-152 (the functional twin of ``Convolve.call``)."""
    return Convolve(final_K)(img_stack, filts)


class Convolve_perlayer:
    """model_library.py:153-168: the same filter applied frame by frame, each scaled by the number of frames.

    ``Convolve_perlayer(K, T)(conv_stack [N,H,W,T], filts [N,H,W,K,K,T]) -> [N,H,W,T]``."""

    def __init__(self, final_K, burst_length, name='convolve_perlayer', **kwargs):
        self.final_K = final_K
        self.burst_length = burst_length
        self.name = name

    def __call__(self, conv_stack, filts):
        from . import ops
        return ops.convolve_filts(conv_stack, filts, self.final_K)[..., 1:]

    call = __call__
