"""Known-answer tests that pin the CPU oracle (no GPU).

The reference has no tests; these are the analytic identities derivable from
its own code, plus cross-checks between independent formulations inside the oracle.
"""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import oracle
from oracle import model as omodel
from oracle import preprocess as opre
from imageenhancement_mp_b200 import synth, weights

P = dict(synth.DEFAULT_PARAMS)


def test_param_and_flop_counts_match_survey():
    L = weights.simplemodel_layers(P)
    W = weights.init_weights(L, scheme="zeros")
    assert weights.count_params(W) == 50_447_218
    per_px, per_img = weights.conv_flops(L, P, 0, 0)
    assert per_px == 4_256_640 and per_img == 691_955_712
    Lk = weights.basis_kpn_layers(dict(P, BURST_LENGTH=8, layer_type="dualparams"))
    assert weights.count_params(weights.init_weights(Lk, scheme="zeros")) == 77_891_610    # 77.9 M


def test_zero_weights_give_box_mean():
    """Coef = 1/B and Bas = 1/(K*K*T) => frame t out = zero-padded 15x15 box mean (x T cancels 1/T)."""
    W = weights.init_weights(weights.simplemodel_layers(P), scheme="zeros")
    x, _ = synth.make_batch(2, 32, 40, P)
    out, bas, ob = oracle.simplemodel_forward(W, P, x)
    b = x[..., :4].permute(0, 3, 1, 2)
    box = F.avg_pool2d(F.pad(b, (7, 7, 7, 7)), 15, 1).permute(0, 2, 3, 1)
    assert torch.allclose(out[..., 1:], box, atol=1e-6)
    assert torch.allclose(out[..., 0], box.mean(-1), atol=1e-6)
    assert torch.allclose(bas, torch.full_like(bas, 1 / 900))


def test_filter_is_a_partition_of_unity():
    g = torch.Generator().manual_seed(3)
    coef = torch.softmax(torch.randn(1, 24, 24, 10, generator=g), -1).double()
    bas = omodel.basis_softmax(torch.randn(1, 15, 15, 40, generator=g).double(), 15, 4, 10)
    ones = torch.ones(1, 24, 24, 4, dtype=torch.float64)
    out = oracle.kpn_apply_literal(ones, coef, bas)
    assert torch.allclose(out[:, 7:-7, 7:-7, 0], torch.ones(1, 10, 10, dtype=torch.float64), atol=1e-12)


def test_literal_and_algebraic_filter_agree():
    g = torch.Generator().manual_seed(4)
    burst = torch.rand(2, 20, 28, 3, generator=g).double()
    coef = torch.softmax(torch.randn(2, 20, 28, 7, generator=g), -1).double()
    bas = omodel.basis_softmax(torch.randn(2, 15, 15, 21, generator=g).double(), 15, 3, 7)
    a = oracle.kpn_apply_literal(burst, coef, bas)
    b = oracle.kpn_apply_algebraic(burst, coef, bas)
    assert float((a - b).abs().max()) < 1e-13


def test_srgb_known_values():
    v = oracle.sRGBforward(torch.tensor([0., .0031308, 1., 2.], dtype=torch.float64))
    assert torch.allclose(v, torch.tensor([0., .040450, 1., 1 + 1.055 / 2.4], dtype=torch.float64), atol=2e-5)


def test_psnr_constant_error_is_20db():
    a = torch.rand(3, 30, 30, dtype=torch.float64)
    assert abs(float(oracle.psnr_tf_batch(a + 0.1, a)) - 20.0) < 1e-9


def test_invert_preproc_shapes_and_scaling():
    img = torch.full((2, 40, 48), 0.25, dtype=torch.float64)
    wl = torch.tensor([0.5, 1.0], dtype=torch.float64).view(2, 1, 1, 1)
    out = oracle.invert_preproc(img, wl)
    assert out.shape == (2, 24, 32)
    assert torch.allclose(out[0], oracle.sRGBforward(torch.tensor(0.5, dtype=torch.float64)).expand(24, 32))
    assert torch.allclose(out[1], oracle.sRGBforward(torch.tensor(0.25, dtype=torch.float64)).expand(24, 32))


def test_gradient_loss_linear_ramp():
    y, x = torch.meshgrid(torch.arange(12.), torch.arange(16.), indexing="ij")
    a = (2 * x + 3 * y)[None]
    z = torch.zeros_like(a)
    # |d/dy| = 3/2, |d/dx| = 2/2 -> mean over both components = 1.25
    assert abs(float(oracle.gradient_loss(a, z)) - 1.25) < 1e-6


def test_fp32_and_fp64_forward_agree():
    W = weights.init_weights(weights.simplemodel_layers(P), scheme="stress")
    x, _ = synth.make_batch(1, 32, 32, P)
    o32 = oracle.simplemodel_forward(W, P, x)[0]
    W64 = {k: (w.double(), b.double()) for k, (w, b) in W.items()}
    o64 = oracle.simplemodel_forward(W64, P, x.double())[0]
    assert float((o32.double() - o64).abs().max()) < 2e-5


def test_upsample_matches_manual_half_pixel():
    x = torch.arange(6.).view(1, 2, 3, 1)
    y = omodel.upsample_bilinear(x, 2)[0, :, :, 0]
    # half-pixel centres: out[0] = in[0], out[1] = .75 in[0] + .25 in[1], ...
    assert torch.allclose(y[0, :3], torch.tensor([0., 0.25, 0.75]))
    assert torch.allclose(y[1, 0], torch.tensor(0.75))            # .75*0 + .25*3
    legacy = omodel.upsample_bilinear(x, 2, legacy=True)[0, :, :, 0]
    assert not torch.allclose(y, legacy)                          # the TF1 kernel differs (kept as a switch)


def test_ssim_identities():
    a = torch.rand(2, 40, 44)
    assert torch.allclose(oracle.ssim(a, a), torch.ones(2, dtype=torch.float64), atol=1e-12)
    b = (a + 0.2 * torch.randn(2, 40, 44)).clamp(0, 1)
    s = oracle.ssim(a, b)
    assert torch.all(s < 1) and torch.allclose(s, oracle.ssim(b, a))


def test_ssim_against_scipy_separable_gaussian():
    """Independent restatement of tf.image.ssim's published definition (Wang et al. 2004 with an 11-tap sigma-1.5
    Gaussian window, VALID region, K1 = .01, K2 = .03, L = 1) on scipy.ndimage: separable correlate1d, the
    variance / covariance form of the formula - against oracle.ssim's single 2-D convolution + moment form."""
    import numpy as np
    from scipy import ndimage
    rng = np.random.default_rng(5)
    a = rng.random((3, 37, 52))
    b = np.clip(a + 0.15 * rng.standard_normal(a.shape), 0, 1)
    x = np.arange(11) - 5.0
    g = np.exp(-0.5 * (x / 1.5) ** 2)
    g /= g.sum()

    def blur(img):                                          # VALID part of the separable Gaussian correlation
        t = ndimage.correlate1d(img, g, axis=1, mode="constant")
        t = ndimage.correlate1d(t, g, axis=2, mode="constant")
        return t[:, 5:-5, 5:-5]
    mu_a, mu_b = blur(a), blur(b)
    var_a, var_b, cov = blur(a * a) - mu_a ** 2, blur(b * b) - mu_b ** 2, blur(a * b) - mu_a * mu_b
    c1, c2 = 0.01 ** 2, 0.03 ** 2
    ref = ((2 * mu_a * mu_b + c1) * (2 * cov + c2) / ((mu_a ** 2 + mu_b ** 2 + c1) * (var_a + var_b + c2))).mean(axis=(1, 2))
    got = oracle.ssim(torch.from_numpy(a), torch.from_numpy(b)).numpy()
    assert np.allclose(got, ref, rtol=0, atol=1e-12)


def test_preprocess_area_resize_matches_cv2():
    cv2 = pytest.importorskip("cv2")
    g = torch.Generator().manual_seed(5)
    img = torch.rand(1, 64, 96, 1, generator=g)
    ours = opre.area_down(img, 4)[0, :, :, 0].numpy()
    ref = cv2.resize(img[0, :, :, 0].numpy(), (24, 16), interpolation=cv2.INTER_AREA)
    assert np.allclose(ours, ref, atol=1e-6)


def test_preprocess_statistics_and_layout():
    params = dict(P, height=16, width=16, BURST_LENGTH=3)
    up, jit = 4, 16
    hs = 16 * up + 2 * jit * up
    src = torch.full((hs, hs, 1), 128, dtype=torch.uint8)
    draws = {"crop0": (0, 0), "use_big": [True, False], "frame_off": [(3, 5), (1, 2)], "white_level": 0.5,
             "sig_read": 0.01, "sig_shot": 0.05, "n_read": torch.zeros(16, 16, 3), "n_shot": torch.zeros(16, 16, 3)}
    x, t = opre.preprocess_image(src, params, draws)
    val = 0.5 * (128 / 255) ** 2.2
    assert x.shape == (16, 16, 4) and t.shape == (16, 16, 2)
    assert torch.allclose(x[..., :3], torch.full((16, 16, 3), val), atol=1e-6)
    assert torch.allclose(x[..., 3], torch.full((16, 16), math.sqrt(0.01 ** 2 + val * 0.05 ** 2)), atol=1e-6)
    assert torch.allclose(t[..., 1], torch.full((16, 16), 0.5))
    assert opre.frame_origins(params, draws) == [(64, 64), (3, 5), (57, 58)]


def test_eval_report_reproduces_leading_zero_bias():
    steps = [{"loss1": 1.0, "perlayer_loss": 2.0, "psnr": 20.0, "psnr_perlayer": [30.0, 1, 1, 1],
              "psnr_noise0": 10.0, "psnr_average": 12.0}] * 3
    r = oracle.eval_report(steps, 4)
    assert r["val_psnrnoshow0"] == pytest.approx(30.0 * 3 / 4)       # eval.py:136,193 start the list with [0]
    assert r["val_psnrnoshow0_unbiased"] == pytest.approx(30.0)
    assert r["val_total_loss"] == pytest.approx(3.0)


def test_philox4x32_10_known_answers():
    """The device noise generator is Philox4x32-10; its numpy restatement against the Random123 known-answer vectors
    (kat_vectors: philox4x32 10 rounds)."""
    from oracle import preprocess as opre
    got = [int(v[0]) for v in opre.philox4x32_10([0], [0], [0], [0], 0, 0)]
    assert got == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    f = 0xffffffff
    got = [int(v[0]) for v in opre.philox4x32_10([f], [f], [f], [f], f, f)]
    assert got == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    got = [int(v[0]) for v in opre.philox4x32_10([0x243f6a88], [0x85a308d3], [0x13198a2e], [0x03707344], 0xa4093822, 0x299f31d0)]
    assert got == [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]
    zs, zr = opre.philox_normals(1234, 200000)
    assert abs(float(zs.mean())) < 0.01 and abs(float(zs.std()) - 1) < 0.01
    assert abs(float(zr.mean())) < 0.01 and abs(float(zr.std()) - 1) < 0.01
    assert abs(float((zs * zr).mean())) < 0.01


def test_conv_pool_resize_against_independent_implementations():
    """The TensorFlow semantics the oracle ASSUMES, cross-checked against independent
    implementations of the same published definitions that exist in this image: scipy's correlate2d for
    Conv2D(..., 'same') / 'valid' (cross-correlation, symmetric zero padding), a numpy block-max for
    MaxPooling2D(2,2), OpenCV's half-pixel-centre INTER_LINEAR for UpSampling2D('bilinear') and scipy's softmax."""
    import cv2
    import numpy as np
    from scipy.signal import correlate2d
    from scipy.special import softmax
    from oracle import model as omodel
    g = torch.Generator().manual_seed(3)
    x = torch.randn(1, 9, 11, 3, generator=g, dtype=torch.float64)
    w = torch.randn(3, 3, 3, 4, generator=g, dtype=torch.float64)
    b = torch.randn(4, generator=g, dtype=torch.float64)
    for padding, mode in (("same", "same"), ("valid", "valid")):
        got = omodel.conv2d_relu(x, (w, b), padding)[0].numpy()
        ref = np.stack([sum(correlate2d(x[0, :, :, ci].numpy(), w[:, :, ci, co].numpy(), mode=mode, boundary="fill")
                            for ci in range(3)) + float(b[co]) for co in range(4)], axis=-1)
        assert np.allclose(got, np.maximum(ref, 0), atol=1e-12)
    xp = torch.randn(2, 6, 8, 5, generator=g, dtype=torch.float64)
    blocks = xp.numpy().reshape(2, 3, 2, 4, 2, 5).max(axis=(2, 4))
    assert np.array_equal(omodel.maxpool2(xp).numpy(), blocks)
    xu = torch.rand(1, 5, 7, 3, generator=g, dtype=torch.float32)
    for s in (2, 8):
        ref = cv2.resize(xu[0].numpy(), (7 * s, 5 * s), interpolation=cv2.INTER_LINEAR)
        assert np.allclose(omodel.upsample_bilinear(xu, s)[0].numpy(), ref, atol=1e-5)
    ob = torch.randn(2, 15, 15, 40, generator=g, dtype=torch.float64)
    bas = omodel.basis_softmax(ob, 15, 4, 10)                                   # softmax over the 900 taps per basis
    ref = softmax(ob.numpy().reshape(2, 900, 10), axis=1).reshape(2, 15, 15, 4, 10)
    assert np.allclose(bas.numpy(), ref, atol=1e-14)


def test_cost_volume_known_answers():
    """cost_volume (data_utils.py:97-113): analytic cases + an independent numpy restatement.
    Uniform taps: every basis is the same (variance 0) and every per-frame tap sum is 1/T <= 0.75 (divergent 0).
    One-hot at (tap 0, frame 0) for every basis: variance 0 again, T*B columns of which B reach 1.0 -> 0.25^2 excess:
    cost = 0.1 * 0.0625 / T."""
    N, K, T, B = 2, 15, 4, 10
    uni = torch.full((N, K, K, T, B), 1.0 / (K * K * T), dtype=torch.float64)
    assert float(oracle.cost_volume(uni)) == pytest.approx(0.0, abs=1e-15)
    hot = torch.zeros(N, K, K, T, B, dtype=torch.float64)
    hot[:, 0, 0, 0, :] = 1.0
    assert float(oracle.cost_volume(hot)) == pytest.approx(0.1 * 0.0625 / T, rel=1e-12)
    g = torch.Generator().manual_seed(11)
    bas = omodel.basis_softmax(torch.relu(torch.randn(N, K, K, T * B, generator=g, dtype=torch.float64) * 6), K, T, B)
    a = bas.numpy()
    var = a.var(axis=-1).mean()                                       # population variance across the bases
    sums = a.reshape(N, K * K, T * B).sum(axis=1)
    ref = -var + 0.1 * np.mean((np.maximum(sums, 0.75) - 0.75) ** 2)
    assert float(oracle.cost_volume(bas)) == pytest.approx(ref, rel=1e-10)
    assert ref < 0 and (sums > 0.75).any()                            # both terms exercised
