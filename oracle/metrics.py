"""Torch-CPU restatement of the metric / loss functions (test infrastructure, parity unpinned).

Follows /root/reference/data_utils.py:24-164 and the per-batch body and report of
/root/reference/eval.py:139-195.  Function names and argument order are the reference's.
"""
from __future__ import annotations

import math

import numpy as np
import torch


def sRGBforward(x):
    """data_utils.py:24-36."""
    b = .0031308
    gamma = 1. / 2.4
    a = .055
    k0 = 12.92
    gammafn = (1 + a) * torch.pow(torch.clamp(x, min=b), gamma) - a
    srgb = torch.where(x < b, k0 * x, gammafn)
    k1 = (1 + a) * gamma
    srgb = torch.where(x > 1, k1 * x - k1 + 1, srgb)
    return srgb


def gradient(imgs):
    """data_utils.py:37-38."""
    return torch.stack([.5 * (imgs[..., 1:, :-1] - imgs[..., :-1, :-1]),
                        .5 * (imgs[..., :-1, 1:] - imgs[..., :-1, :-1])], dim=-1)


def gradient_loss(guess, truth):
    """data_utils.py:40-41."""
    return (gradient(guess) - gradient(truth)).abs().mean()


def invert_preproc(imgs, white_level):
    """data_utils.py:42-45: divide by the per-image white level, sRGB curve, crop 8 px."""
    lbuff = 8
    wl = white_level.reshape(-1)
    # tf.transpose with no perm reverses all axes
    rev = tuple(range(imgs.dim() - 1, -1, -1))
    y = (imgs.permute(rev) / wl).permute(rev)
    return sRGBforward(y)[:, lbuff:-lbuff, lbuff:-lbuff, ...]


def basic_img_loss(img, truth):
    """data_utils.py:46-51."""
    l2_pixel = ((img - truth) ** 2).mean()
    l1_grad = gradient_loss(img, truth)
    return l2_pixel + l1_grad


def deblur_layer_loss(y_pred, invert_gt, white_noise):
    """data_utils.py:52-73: sum over the T per-frame outputs of basic_img_loss."""
    burst_size = y_pred.shape[-1] - 1
    loss = basic_img_loss(invert_preproc(y_pred[..., 1], white_noise), invert_gt)
    for i in range(burst_size - 1):
        loss = loss + basic_img_loss(invert_preproc(y_pred[..., i + 2], white_noise), invert_gt)
    return loss


def invert_deblur_layer(y_pred, white_noise):
    """data_utils.py:74-80: T inverted frames concatenated along the last (width) axis."""
    burst_size = y_pred.shape[-1] - 1
    outs = [invert_preproc(y_pred[..., i + 1], white_noise) for i in range(burst_size)]
    return torch.cat(outs, dim=-1)


def deblur_loss(invert_deblur, invert_gt):
    """data_utils.py:81-96."""
    return basic_img_loss(invert_deblur, invert_gt)


def cost_volume(Basis):
    """data_utils.py:97-113: ``-variance + 0.1 * divergent`` of Basis [N,K,K,T,B] (the similarity regulariser that
    eval.py:159-162 reports when params["ps"])."""
    ish = Basis.shape
    average = Basis.mean(dim=-1)                                               # :100
    average_2 = Basis.square().mean(dim=-1)                                    # :102
    cost = average_2 - average.square()                                        # :105
    variance = cost.mean()                                                     # :107
    sum_value = Basis.reshape(ish[0], ish[1] ** 2, -1).sum(dim=1).clamp_min(0.75)   # :110
    divergent = (sum_value - 0.75).square().mean()                             # :111
    return -variance + 0.1 * divergent                                         # :113


def psnr_tf_batch(estimate, truth):
    """data_utils.py:118-119: per-image PSNR (peak 1), then batch mean."""
    n = estimate.shape[0]
    mse = ((estimate - truth) ** 2).reshape(n, -1).mean(dim=1)
    return (-10. * torch.log(mse) / math.log(10.)).mean()


def psnr_deblur(invert_deblur, invert_gt):
    """data_utils.py:121-130."""
    return psnr_tf_batch(invert_deblur, invert_gt)


def psnr_each_layer(invert_gt, white_noise, y_pred):
    """data_utils.py:131-144."""
    burst_size = y_pred.shape[-1] - 1
    psnr = {}
    for i in range(burst_size):
        layer = invert_preproc(y_pred[..., i + 1], white_noise)
        psnr['da{}_noshow'.format(i)] = psnr_tf_batch(layer, invert_gt)
    return psnr


def psnr_burst0(invert_gt, white_noise, x_batch_burst):
    """data_utils.py:145-154."""
    return psnr_tf_batch(invert_preproc(x_batch_burst[..., 0], white_noise), invert_gt)


def psnr_average_f(invert_gt, white_noise, x_batch_burst):
    """data_utils.py:155-164."""
    return psnr_tf_batch(invert_preproc(x_batch_burst.mean(dim=-1), white_noise), invert_gt)


def eval_step(reconstructed, x_batch_burst, x_batch_truth, burst_length):
    """Body of the validation loop, eval.py:141-182, for one batch.

    Returns a dict of python floats: loss1, perlayer_loss, psnr, psnr_perlayer[T],
    psnr_noise0, psnr_average.
    """
    burst_images = x_batch_burst[..., 0:burst_length]                          # :141
    white_noise = x_batch_truth[..., 1].unsqueeze(-1)                          # :144
    white_noise = white_noise.mean(dim=1, keepdim=True).mean(dim=2, keepdim=True)  # :145
    gt = x_batch_truth[..., 0]                                                 # :146
    invert_gt = invert_preproc(gt, white_noise)                                # :147
    invert_deblur = invert_preproc(reconstructed[..., 0], white_noise)         # :148-149
    loss1 = deblur_loss(invert_deblur, invert_gt)                              # :151
    perlayer_loss = deblur_layer_loss(reconstructed, invert_gt, white_noise)   # :156
    per = psnr_each_layer(invert_gt, white_noise, reconstructed)               # :174
    return {
        "loss1": float(loss1),
        "perlayer_loss": float(perlayer_loss),
        "psnr": float(psnr_deblur(invert_deblur, invert_gt)),                  # :170
        "psnr_perlayer": [float(per['da{}_noshow'.format(i)]) for i in range(burst_length)],
        "psnr_noise0": float(psnr_burst0(invert_gt, white_noise, burst_images)),     # :176
        "psnr_average": float(psnr_average_f(invert_gt, white_noise, burst_images)),  # :179
    }


def eval_report(steps, burst_length):
    """Aggregation of eval.py:183-195 over a list of eval_step dicts.

    Keras ``metrics.Mean`` of the per-batch losses; ``np.mean`` of the per-batch PSNRs.  The
    per-layer list starts as ``[0]`` (eval.py:136), so ``val_psnrnoshow0`` carries a leading
    zero - reproduced here, with the unbiased value alongside.
    """
    loss1 = float(np.mean([s["loss1"] for s in steps]))
    perl = float(np.mean([s["perlayer_loss"] for s in steps]))
    per0 = [0] + [s["psnr_perlayer"][0] for s in steps]
    return {
        "val_deblur_loss": loss1,
        "val_perlayer_loss": perl,
        "val_total_loss": loss1 + perl,
        "val_psnr": float(np.mean([s["psnr"] for s in steps])),
        "val_psnrnoshow0": float(np.mean(per0)),
        "val_psnrnoshow0_unbiased": float(np.mean(per0[1:])),
        "val_psnrburst0": float(np.mean([s["psnr_noise0"] for s in steps])),
        "val_psnraverage": float(np.mean([s["psnr_average"] for s in steps])),
    }
