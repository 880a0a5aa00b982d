"""Layer tables and random initialisers for the reference networks.

Weights live in a flat dict keyed by the reference's Keras attribute paths
(/root/reference/model_library.py:323-368 for Simplemodel, :196-227 for Basis_kpn), each
entry ``(kernel [kh,kw,Cin,Cout] float32, bias [Cout] float32)`` - HWIO, like
``layers.Conv2D`` stores them.  ``save_npz`` / ``load_npz`` give the flat ``.npz`` wire
format that replaces the TF object-graph checkpoint of eval.py:112-118.
"""
from __future__ import annotations

import math

import numpy as np
import torch

ADD_LENGTHS = {"singlestd": 1, "dualparams": 2, "empty": 0}     # model_library.py:319-321


def _block(name, cin, cout, n=3):
    specs = [(f"{name}.conv2d1", 3, cin, cout, "same")]
    for i in range(2, n + 1):
        specs.append((f"{name}.conv2d{i}", 3, cout, cout, "same"))
    return specs


def simplemodel_layers(params):
    """(name, k, Cin, Cout, padding) for every conv of Simplemodel, in call order."""
    T, B = params["BURST_LENGTH"], params["Basis_num"]
    cin = T + ADD_LENGTHS[params["layer_type"]]
    L = [("layer0", 3, cin, 64, "same")]
    L += _block("down1", 64, 64, 2) + _block("down2", 64, 128, 2) + _block("down5", 128, 1024, 2)
    L += [("layer1_1", 3, 1024, 1024, "same")]
    L += _block("Coef_up1", 1024 + 1024, 512) + _block("Coef_up4", 512 + 128, 64) + _block("Coef_up5", 64 + 64, 64)
    L += [("layer2_1", 3, 64, 64, "same"), ("coef", 3, 64, B, "same")]
    L += _block("Basis_up1", 1024 + 1024, 512) + _block("Basis_up4", 512 + 128, 128)
    L += [("layer3_1", 2, 128, 128, "valid"), ("layer3_3", 3, 128, T * B, "same")]
    return L


def basis_kpn_layers(params):
    """Same for the five-level Basis_kpn (model_library.py:196-227)."""
    T, B = params["BURST_LENGTH"], params["Basis_num"]
    cin = T + ADD_LENGTHS[params["layer_type"]]
    L = [("layer0", 3, cin, 64, "same")]
    L += _block("down1", 64, 64, 2) + _block("down2", 64, 128, 2) + _block("down3", 128, 256, 2)
    L += _block("down4", 256, 512, 2) + _block("down5", 512, 1024, 2)
    L += [("layer1_1", 3, 1024, 1024, "same"), ("layer1_2", 3, 1024, 1024, "same")]
    L += _block("Coef_up1", 2048, 512) + _block("Coef_up2", 512 + 512, 256) + _block("Coef_up3", 256 + 256, 128)
    L += _block("Coef_up4", 128 + 128, 64) + _block("Coef_up5", 64 + 64, 64)
    L += [("layer2_1", 3, 64, 64, "same"), ("layer2_2", 3, 64, 64, "same"), ("coef", 3, 64, B, "same")]
    L += _block("Basis_up1", 2048, 512) + _block("Basis_up2", 512 + 512, 256)
    L += _block("Basis_up3", 256 + 256, 256) + _block("Basis_up4", 256 + 128, 128)
    L += [("layer3_1", 2, 128, 128, "valid"), ("layer3_2", 3, 128, 128, "same"),
          ("layer3_3", 3, 128, T * B, "same")]
    return L


def init_weights(layers, seed=1234, scheme="glorot"):
    """Random weights on CPU from a seeded generator.

    glorot: Keras' default for Conv2D - glorot_uniform kernel, zero bias (what the reference
            evaluates with when no checkpoint exists, eval.py:114-118).
    stress: He-normal kernels and N(0, 0.1) biases, so that the two softmaxes see O(1)
            logits and the output departs from the 15x15 box mean.
    """
    g = torch.Generator().manual_seed(seed)
    W = {}
    for name, k, cin, cout, _ in layers:
        fan_in, fan_out = k * k * cin, k * k * cout
        if scheme == "glorot":
            lim = math.sqrt(6.0 / (fan_in + fan_out))
            w = (torch.rand(k, k, cin, cout, generator=g) * 2 - 1) * lim
            b = torch.zeros(cout)
        elif scheme == "stress":
            w = torch.randn(k, k, cin, cout, generator=g) * math.sqrt(2.0 / fan_in)
            b = torch.randn(cout, generator=g) * 0.1
        elif scheme == "zeros":
            w = torch.zeros(k, k, cin, cout)
            b = torch.zeros(cout)
        else:
            raise ValueError(scheme)
        W[name] = (w, b)
    return W


def count_params(W):
    return sum(w.numel() + b.numel() for w, b in W.values())


def conv_flops(layers, params, H, Wd, full=True):
    """(FLOP per full-resolution pixel for the trunk, FLOP per image for the basis branch).

    FLOPs = 2*kh*kw*Cin*Cout per output pixel of the layer.  Only valid for the
    Simplemodel table; resolution factors follow the call graph in model_library.py:376-428.
    """
    res = {"layer0": 1, "down1": 1, "down2": 4, "down5": 16, "layer1_1": 64, "Coef_up1": 16,
           "Coef_up4": 4, "Coef_up5": 1, "layer2_1": 1, "coef": 1}
    per_px = 0.0
    per_img = 0.0
    for name, k, cin, cout, _ in layers:
        key = name.split(".")[0]
        f = 2.0 * k * k * cin * cout
        if key in res:
            per_px += f / res[key]
        elif key == "Basis_up1":
            per_img += f * 4
        elif key == "Basis_up4":
            per_img += f * 256
        elif key in ("layer3_1", "layer3_3"):
            per_img += f * 225
    return per_px, per_img


def save_npz(path, W):
    flat = {}
    for name, (w, b) in W.items():
        flat[name + "/kernel"] = w.numpy()
        flat[name + "/bias"] = b.numpy()
    np.savez(path, **flat)


def load_npz(path):
    z = np.load(path)
    names = sorted({k.rsplit("/", 1)[0] for k in z.files})
    return {n: (torch.from_numpy(z[n + "/kernel"]), torch.from_numpy(z[n + "/bias"])) for n in names}


def load(path, root="net", layers=None):
    """Weights from a file: the flat ``.npz`` of this package, or a TensorFlow object-graph checkpoint of the reference
    (eval.py:112-118) - a checkpoint directory (its ``checkpoint`` state file names the latest prefix, like
    ``tf.train.latest_checkpoint``) or a prefix ``.../ckpt-12`` (``.index`` + ``.data-*`` files).  ``root`` is the
    keyword the model was stored under in ``tf.train.Checkpoint`` (``net`` in the reference).  The TF reader is a
    restatement of the published file formats, not validated against TensorFlow here (tf_checkpoint.py).
    ``layers``: the model's layer list (``simplemodel_layers(params)``), only needed for checkpoints whose variables
    are named through Keras' ``layer_with_weights-N`` edges."""
    import os
    from . import tf_checkpoint as tfc
    if path.endswith(".npz"):
        return load_npz(path)
    order = None
    if layers is not None:
        order = []
        for entry in layers:
            top = entry[0].split(".")[0]
            if top not in order:
                order.append(top)
    def _tf(prefix):
        import sys
        print("imageenhancement_mp_b200.weights.load: TensorFlow checkpoint reader: UNPINNED against a TensorFlow-written "
              "file (restated from the published formats; no TensorFlow in the build image) - check the report numbers "
              f"against the reference before trusting them [{prefix}]", file=sys.stderr)
        return tfc.load_tf_checkpoint(prefix, root=root, layer_order=order)

    if os.path.isdir(path):
        prefix = tfc.latest_checkpoint(path)
        if prefix is None:
            raise FileNotFoundError(f"{path}: no 'checkpoint' state file (tf.train.CheckpointManager writes one)")
        return _tf(prefix)
    if path.endswith(".index"):
        path = path[:-len(".index")]
    if os.path.exists(path + ".index"):
        return _tf(path)
    raise FileNotFoundError(f"{path}: neither an .npz file nor a TensorFlow checkpoint prefix / directory")
