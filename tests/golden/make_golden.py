#!/usr/bin/env python
"""Regenerates tests/golden/*.npz from the CPU oracle (run from the repo root).

The reference ships no golden vectors and cannot run here (TensorFlow is absent), so these
fixtures pin the ORACLE against itself across rounds (regression), not against the reference:
parity with the reference stays "unpinned" (see oracle/__init__.py).  Weights are not stored
(50 M parameters); they are regenerated from the seed and summarised by a checksum.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from imageenhancement_mp_b200 import synth, weights  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def weight_checksum(W):
    return float(sum(float(w.double().abs().sum()) + float(b.double().abs().sum()) for w, b in W.values()))


def make(name, arch, params, n, h, w, scheme):
    torch.manual_seed(0)
    layers = weights.simplemodel_layers(params) if arch == "simple" else weights.basis_kpn_layers(params)
    W = weights.init_weights(layers, seed=1234, scheme=scheme)
    x, truth = synth.make_batch(n, h, w, params, seed=1234)
    xp, _ = synth.pad_to_multiple(x, 8 if arch == "simple" else 32)
    taps = {}
    fwd = oracle.simplemodel_forward if arch == "simple" else oracle.basis_kpn_forward
    res = fwd(W, params, xp, taps=taps)
    out = res[0][:, :h, :w]
    step = oracle.eval_step(out, x, truth, params["BURST_LENGTH"])
    np.savez_compressed(
        os.path.join(HERE, name + ".npz"),
        x=x.numpy(), truth=truth.numpy(), output=out.numpy(), Bas=res[1].numpy(),
        originbasis=(res[2] if arch == "simple" else taps["originbasis"]).numpy(),
        Coef=taps["Coef"].numpy().astype(np.float16), coef_logits_absmean=float(taps["coef_logits"].abs().mean()),
        weight_checksum=weight_checksum(W),
        report=np.array([step["loss1"], step["perlayer_loss"], step["psnr"], *step["psnr_perlayer"],
                         step["psnr_noise0"], step["psnr_average"]], dtype=np.float64))
    print(name, "written; psnr", step["psnr"])


if __name__ == "__main__":
    P = dict(synth.DEFAULT_PARAMS)
    only = sys.argv[1:]
    if only:                                    # regenerate just the named fixtures
        _make = make
        make = lambda name, *a: _make(name, *a) if name in only else None
    make("simple_glorot_32", "simple", P, 2, 32, 32, "glorot")          # eval.py defaults (32x32, T=4)
    make("simple_stress_32", "simple", P, 2, 32, 32, "stress")
    make("simple_stress_100", "simple", P, 1, 100, 100, "stress")       # BASELINE config shape (padded to 104)
    make("simple_stress_T2", "simple", dict(P, BURST_LENGTH=2), 1, 40, 48, "stress")   # 3-channel reading
    # the second entry point with the remote/ settings (running_train_remote.py:29,34: T = 8, Basis_num = 50, 64x64)
    make("basis_kpn_stress_64", "basis_kpn", dict(P, BURST_LENGTH=8, layer_type="dualparams", Basis_num=50), 1, 64, 64,
         "stress")
