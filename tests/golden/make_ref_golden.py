#!/usr/bin/env python
"""Golden vectors produced by RUNNING THE REFERENCE'S OWN CODE (run from the repo root, where /root/reference exists):

    python tests/golden/make_ref_golden.py [--reference /root/reference] [names...]

Imports the UNMODIFIED ``model_library.py`` and ``data_utils.py`` from the reference tree, builds ``Simplemodel`` /
``Basis_kpn`` with the repo's seeded weight dict assigned to the Keras layers by attribute name, runs the repo's
synthetic batches (``synth.make_batch``) through ``model(x)``, then the body of ``evaluate()``'s validation loop
(eval.py:141-181 - restated here call by call, because ``evaluate()`` itself needs PNG folders, argparse and a
checkpoint directory) through the reference's own metric functions, and ``DataLoader.preprocess_image`` on a synthetic
uint8 image.  Writes ``tests/golden/<prefix>_*.npz``.

Backend.  With a real TensorFlow 2.x importable, that is used and the files are named ``tf_*.npz`` - parity is then
pinned to the reference outright.  Without one (this build container: no TensorFlow, no network) the reference runs
over ``oracle/tf_standin.py``, a ~300-line stand-in for the ~60 TensorFlow / Keras primitives the reference calls; the
files are named ``ref_*.npz`` and say ``backend = "standin"``.  Then the WIRING of every function on the path is the
reference's own executing code; what remains assumed is the meaning of the primitives (Conv2D, UpSampling2D bilinear =
half-pixel, resize AREA = box mean ...; listed in the stand-in and in DESIGN.md section 6).

tests/test_golden.py and tests/test_gpu_model.py check the oracle and the CUDA path against whichever files exist,
preferring ``tf_*`` over ``ref_*``.
"""
import argparse
import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from imageenhancement_mp_b200 import synth, weights  # noqa: E402

# Keras attribute path (model_library.py:196-227, 323-368, 72-73, 89-91) -> key of the repo's weight dict
def layer_names(arch, params):
    layers = weights.simplemodel_layers(params) if arch == "simple" else weights.basis_kpn_layers(params)
    return [name for name, *_ in layers]


def resolve(model, dotted):
    obj = model
    for part in dotted.split("."):
        obj = getattr(obj, part)
    return obj


def to_np(t):
    return t.numpy() if hasattr(t, "numpy") else np.asarray(t)


def load_reference(ref_dir):
    """Returns (backend name, tf module, model_library, data_utils) with the reference files imported unmodified."""
    try:
        import tensorflow as tf  # noqa: F401
        backend = "tensorflow-" + tf.__version__
        if getattr(tf, "__standin__", False):
            backend = "standin"
    except Exception:
        from oracle import tf_standin
        tf = tf_standin.install()
        backend = "standin"
    sys.path.insert(0, ref_dir)
    du = importlib.import_module("data_utils")
    ml = importlib.import_module("model_library")
    for mod, fname in ((du, "data_utils.py"), (ml, "model_library.py")):
        assert os.path.samefile(mod.__file__, os.path.join(ref_dir, fname)), f"{fname} was not imported from {ref_dir}"
    return backend, tf, ml, du


def as_tf(tf, a, backend):
    return torch.from_numpy(np.ascontiguousarray(a)) if backend == "standin" else tf.convert_to_tensor(a)


def build_model(tf, ml, backend, arch, params, W, x):
    model = (ml.Simplemodel if arch == "simple" else ml.Basis_kpn)(params)
    model(as_tf(tf, x[:1], backend))                                     # builds the Keras variables
    for name in layer_names(arch, params):
        layer = resolve(model, name)
        k, b = W[name]
        layer.kernel.assign(k.numpy())
        layer.bias.assign(b.numpy())
    return model


def eval_loop_body(tf, du, params, reconstructed, x_batch_burst, x_batch_truth):
    """eval.py:141-181, one batch, through the reference's own functions.  Returns the seven numbers evaluate() would
    accumulate for this batch: loss1, perlayer_loss, psnr, psnr_perlayer[T], psnr_noise0, psnr_average."""
    burst_length = params["BURST_LENGTH"]
    x_batch_burst_images = x_batch_burst[..., 0:burst_length]                                        # :141
    white_noise = tf.expand_dims(x_batch_truth[..., 1], axis=-1)                                     # :144
    white_noise = tf.reduce_mean(tf.reduce_mean(white_noise, axis=1, keepdims=True), axis=2, keepdims=True)   # :145
    gt = x_batch_truth[..., 0]                                                                       # :146
    invert_gt = du.invert_preproc(gt, white_noise)                                                   # :147
    Deblur = reconstructed[..., 0]                                                                   # :148
    invert_deblur = du.invert_preproc(Deblur, white_noise)                                           # :149
    loss1 = du.deblur_loss(invert_deblur, invert_gt)                                                 # :151
    perlayer_loss = du.deblur_layer_loss(reconstructed, invert_gt, white_noise)                      # :156
    invert_perlayers = du.invert_deblur_layer(reconstructed, white_noise)                            # :158
    psnr = du.psnr_deblur(invert_deblur, invert_gt)                                                  # :170
    per = du.psnr_each_layer(invert_gt, white_noise, reconstructed)                                  # :174
    noise0 = du.psnr_burst0(invert_gt, white_noise, x_batch_burst_images)                            # :176
    avg = du.psnr_average_f(invert_gt, white_noise, x_batch_burst_images)                            # :179
    report = [float(loss1), float(perlayer_loss), float(psnr)] + \
             [float(per['da{}_noshow'.format(i)]) for i in range(burst_length)] + [float(noise0), float(avg)]
    return dict(report=np.array(report, dtype=np.float64), invert_gt=to_np(invert_gt), invert_deblur=to_np(invert_deblur),
                invert_perlayer=to_np(invert_perlayers), white_noise=to_np(white_noise))


def make_model_case(ctx, name, arch, params, n, h, w, scheme, out_dir, prefix):
    backend, tf, ml, du = ctx
    layers = weights.simplemodel_layers(params) if arch == "simple" else weights.basis_kpn_layers(params)
    W = weights.init_weights(layers, seed=1234, scheme=scheme)
    x, truth = synth.make_batch(n, h, w, params, seed=1234)
    taps = None
    if backend == "standin":
        from oracle import tf_standin
        tf_standin.LAYER_TAPS = taps = {}
    model = build_model(tf, ml, backend, arch, dict(params), W, x.numpy())
    if taps is not None:
        taps.clear()
    res = model(as_tf(tf, x.numpy(), backend))
    output, Bas = res[0], res[1]
    ev = eval_loop_body(tf, du, params, output, as_tf(tf, x.numpy(), backend), as_tf(tf, truth.numpy(), backend))
    cv = float(du.cost_volume(Bas))                                                                   # eval.py:160-162
    extra = {}
    if len(res) > 2:
        extra["originbasis"] = to_np(res[2])
    if taps is not None:
        # a few intermediates by Keras attribute path (fp16 keeps the files small; they are compared in relative L2)
        keep = ("layer1_1", "Coef_up1", "coef", "Basis_up1", "Basis_up4", "layer3_1") if h * w <= 32 * 32 else ("coef",)
        for k in keep:
            if k in taps:
                extra["tap." + k] = taps[k].numpy().astype(np.float16)
        from oracle import tf_standin
        tf_standin.LAYER_TAPS = None
    chk = float(sum(float(k.double().abs().sum()) + float(b.double().abs().sum()) for k, b in W.values()))
    path = os.path.join(out_dir, f"{prefix}_{name}.npz")
    np.savez_compressed(path, backend=backend, x=x.numpy(), truth=truth.numpy(), output=to_np(output), Bas=to_np(Bas),
                        report=ev["report"], cost_volume=cv, invert_gt=ev["invert_gt"].astype(np.float32),
                        invert_deblur=ev["invert_deblur"].astype(np.float32),
                        invert_perlayer_shape=np.array(ev["invert_perlayer"].shape), weight_checksum=chk, **extra)
    print(f"{path}: backend {backend}, psnr {ev['report'][2]:.4f} dB, {os.path.getsize(path) / 1e3:.0f} kB")


def make_preprocess_case(ctx, name, params, src_hw, seed, out_dir, prefix):
    """DataLoader.preprocess_image (data_utils.py:198-265) on a synthetic uint8 image.  Under the stand-in the random
    draws are logged in call order, so the oracle / the CUDA kernel can replay them (tests/test_golden.py)."""
    backend, tf, ml, du = ctx
    if backend != "standin":
        print(f"skip {name}: replaying TensorFlow's random streams needs the stand-in's draw log")
        return
    from oracle import tf_standin
    g = np.random.default_rng(seed)
    img = g.integers(0, 256, size=(src_hw[0], src_hw[1], 1), dtype=np.uint8)
    import tempfile
    with tempfile.TemporaryDirectory() as empty:                               # an empty folder: the constructor only globs it,
        loader = du.DataLoader(dict(params, batch_size=1, color=False, train_path=empty, percent=1.0))   # and sets .channels
    tf_standin.reseed(seed)
    x, truth = loader.preprocess_image(torch.from_numpy(img), params)
    log = list(tf_standin.DRAW_LOG)
    T = params["BURST_LENGTH"]
    # order of the draws: make_first_truth crop (:439); poisson (:453); per frame flip (:455) + crop (:457);
    # white level (:225); sig_read, sig_shot (:232-233); read normals, shot normals (:463-464)
    assert [k for k, _ in log] == ["random_crop", "random_poisson"] + ["random_uniform", "random_crop"] * (T - 1) + \
        ["random_uniform"] * 3 + ["random_normal"] * 2, [k for k, _ in log]
    crop0 = log[0][1][:2]
    prob = min(float(log[1][1]) / T, 1.0)
    flips = [float(log[2 + 2 * k][1]) for k in range(T - 1)]
    offs = [log[3 + 2 * k][1][:2] for k in range(T - 1)]
    base = 2 + 2 * (T - 1)
    wl = 10.0 ** float(log[base][1].reshape(-1)[0])
    sr = 10.0 ** float(log[base + 1][1].reshape(-1)[0])
    ss = 10.0 ** float(log[base + 2][1].reshape(-1)[0])
    path = os.path.join(out_dir, f"{prefix}_{name}.npz")
    np.savez_compressed(path, backend=backend, image=img, x=to_np(x), truth=to_np(truth), crop0=np.array(crop0),
                        use_big=np.array([f < prob for f in flips]), frame_off=np.array(offs), white_level=wl, sig_read=sr,
                        sig_shot=ss, n_read=log[base + 3][1].numpy(), n_shot=log[base + 4][1].numpy(),
                        params=np.array(sorted((k, str(v)) for k, v in params.items())))
    print(f"{path}: backend {backend}, x {tuple(x.shape)}, {os.path.getsize(path) / 1e3:.0f} kB")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=HERE)
    ap.add_argument("names", nargs="*")
    a = ap.parse_args()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    ctx = load_reference(a.reference)
    prefix = "ref" if ctx[0] == "standin" else "tf"
    P = dict(synth.DEFAULT_PARAMS)
    want = lambda n: not a.names or n in a.names
    cases = [
        ("simple_glorot_32", "simple", P, 2, 32, 32, "glorot"),              # eval.py defaults: 32x32, T = 4
        ("simple_stress_32", "simple", P, 2, 32, 32, "stress"),
        ("simple_stress_104", "simple", P, 1, 104, 104, "stress"),           # the BASELINE patch size, padded to the stride
        ("simple_stress_T2", "simple", dict(P, BURST_LENGTH=2), 1, 40, 48, "stress"),    # run_training_val.py:28
        # the second entry point with the remote/ settings (running_train_remote.py:29,34): T = 8, dualparams
        ("basis_kpn_stress_64", "basis_kpn", dict(P, BURST_LENGTH=8, layer_type="dualparams", Basis_num=10), 1, 64, 64,
         "stress"),
    ]
    for c in cases:
        if want(c[0]):
            make_model_case(ctx, *c, out_dir=a.out, prefix=prefix)
    pre = [
        ("preprocess_T4", dict(P, height=24, width=32), (300, 340), 7),
        ("preprocess_T8_dual", dict(P, height=16, width=24, BURST_LENGTH=8, layer_type="dualparams"), (260, 300), 8),
        ("preprocess_small_source", dict(P, height=24, width=32), (100, 120), 9),    # source smaller than the crop: tf.pad :438
    ]
    for c in pre:
        if want(c[0]):
            make_preprocess_case(ctx, *c, out_dir=a.out, prefix=prefix)


if __name__ == "__main__":
    main()
