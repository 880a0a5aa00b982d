"""The oracle against the committed fixtures of tests/golden/ (regression pin; see make_golden.py)."""
import glob
import os

import numpy as np
import pytest
import torch

import oracle
from imageenhancement_mp_b200 import synth, weights

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = {
    "simple_glorot_32": (dict(synth.DEFAULT_PARAMS), "glorot"),
    "simple_stress_32": (dict(synth.DEFAULT_PARAMS), "stress"),
    "simple_stress_100": (dict(synth.DEFAULT_PARAMS), "stress"),
    "simple_stress_T2": (dict(synth.DEFAULT_PARAMS, BURST_LENGTH=2), "stress"),
    "basis_kpn_stress_64": (dict(synth.DEFAULT_PARAMS, BURST_LENGTH=8, layer_type="dualparams", Basis_num=50), "stress"),
}


def load(name):
    return np.load(os.path.join(HERE, "golden", name + ".npz"))


def test_fixtures_present():
    names = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(HERE, "golden", "*.npz")))
    # ref_* / tf_* are the reference-run fixtures (tests/test_ref_golden.py); the rest are the oracle's regression pins
    assert [n for n in names if not n.startswith(("ref_", "tf_"))] == sorted(CASES)


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden(name):
    params, scheme = CASES[name]
    z = load(name)
    kpn = name.startswith("basis_kpn")
    layers = weights.basis_kpn_layers(params) if kpn else weights.simplemodel_layers(params)
    W = weights.init_weights(layers, seed=1234, scheme=scheme)
    chk = sum(float(w.double().abs().sum()) + float(b.double().abs().sum()) for w, b in W.values())
    assert chk == pytest.approx(float(z["weight_checksum"]), rel=1e-9)      # same weights from the same seed
    x, truth = torch.from_numpy(z["x"]), torch.from_numpy(z["truth"])
    n, h, w, _ = x.shape
    xs, ts = synth.make_batch(n, h, w, params, seed=1234)
    # synthetic inputs are reproducible (ulp slack: pow/interpolate may vectorise differently per host)
    assert torch.allclose(xs, x, rtol=1e-5, atol=1e-7) and torch.allclose(ts, truth, rtol=1e-5, atol=1e-7)
    xp, _ = synth.pad_to_multiple(x, 32 if kpn else 8)
    if kpn:
        taps = {}
        out, bas = oracle.basis_kpn_forward(W, params, xp, taps=taps)
        ob = taps["originbasis"]
    else:
        out, bas, ob = oracle.simplemodel_forward(W, params, xp)
    out = out[:, :h, :w]
    assert np.allclose(out.numpy(), z["output"], atol=2e-5)
    assert np.allclose(bas.numpy(), z["Bas"], atol=1e-7, rtol=1e-3)
    assert np.allclose(ob.numpy(), z["originbasis"], atol=1e-4, rtol=1e-3)
    step = oracle.eval_step(out, x, truth, params["BURST_LENGTH"])
    rep = [step["loss1"], step["perlayer_loss"], step["psnr"], *step["psnr_perlayer"], step["psnr_noise0"],
           step["psnr_average"]]
    assert np.allclose(rep, z["report"], atol=1e-3)
