"""Torch-CPU restatement of DataLoader.preprocess_image (test infrastructure, parity unpinned).

Follows /root/reference/data_utils.py:198-265 and helpers make_first_truth :432-439,
make_truth_hqjitter :442-461, add_read_shot_tf :462-466.  Every random draw the reference
takes from TF's RNG is an explicit argument here (``draws``), so the arithmetic is
deterministic and comparable:

  crop0        (y, x)  offset of tf.image.random_crop in make_first_truth (:439)
  use_big      [T-1]   bool, flip < prob (:455-456)
  frame_off    [T-1]   (y, x) offsets of the per-frame random_crop inside p2use (:457)
  white_level, sig_read, sig_shot   scalars (:225, :232-233)
  n_read, n_shot  [h, w, T] standard normals (:463-464)
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def area_down(x, up):
    """tf.image.resize(..., AREA) with an integer factor == exact up x up box mean (:459)."""
    f, hh, ww, c = x.shape
    return x.reshape(f, hh // up, up, ww // up, up, c).mean(dim=(2, 4))


def preprocess_image(image_u8, params, draws, dtype=torch.float32):
    """image_u8 [Hs,Ws,C] uint8 -> (x [h,w,T+add], truth [h,w,2])."""
    height, width = params["height"], params["width"]
    T = params["BURST_LENGTH"]
    degamma, up = params["degamma"], params["upscale"]
    jitter, smalljitter = params["jitter"], params["smalljitter"]
    img = (image_u8.to(dtype) / 255.) ** degamma                                   # :213
    # make_first_truth :432-439
    j_up = jitter * up
    h_up = height * up + 2 * j_up
    w_up = width * up + 2 * j_up
    v_err = max((h_up - img.shape[0] + 1) // 2, 0)
    h_err = max((w_up - img.shape[1] + 1) // 2, 0)
    img = F.pad(img, (0, 0, h_err, h_err, v_err, v_err))
    cy, cx = draws["crop0"]
    patches = img[cy:cy + h_up, cx:cx + w_up, :]
    assert patches.shape[0] == h_up and patches.shape[1] == w_up
    # make_truth_hqjitter :442-461
    hh, ww = height * up, width * up
    delta_up = (jitter - smalljitter) * up
    small = patches[delta_up:-delta_up, delta_up:-delta_up, :]
    curr = [patches[j_up:-j_up, j_up:-j_up, :]]
    for k in range(T - 1):
        p2use = patches if draws["use_big"][k] else small
        oy, ox = draws["frame_off"][k]
        crop = p2use[oy:oy + hh, ox:ox + ww, :]
        assert crop.shape[0] == hh and crop.shape[1] == ww
        curr.append(crop)
    curr = torch.stack(curr, dim=0)                   # [T, hh, ww, C]
    curr = area_down(curr, up)                        # [T, h, w, C]
    curr = curr.permute(1, 2, 3, 0)                   # [h, w, C, T]
    truth = curr.mean(dim=-2)                         # :220  [h, w, T]
    wl = torch.as_tensor(draws["white_level"], dtype=dtype)
    truth = wl * truth                                # :230 (degamma reset to 1 at :224)
    sr = torch.as_tensor(draws["sig_read"], dtype=dtype)
    ss = torch.as_tensor(draws["sig_shot"], dtype=dtype)
    read = sr * draws["n_read"].to(dtype)             # :463
    shot = torch.sqrt(truth) * ss * draws["n_shot"].to(dtype)   # :464
    noisy = truth + shot + read                       # :465
    lt = params["layer_type"]
    if lt == "singlestd":                             # :256
        sig = torch.sqrt(sr ** 2 + torch.clamp(noisy[..., 0:1], min=0.) * ss ** 2)
    elif lt == "dualparams":                          # :257
        sig = torch.stack([sr, ss]).to(dtype).expand(height, width, 2)
    else:                                             # :258
        sig = noisy[..., 0:0]
    x = torch.cat([noisy, sig], dim=-1)               # :265
    t = torch.cat([truth[..., 0:1], wl.expand(height, width, 1)], dim=-1)
    return x, t


def frame_origins(params, draws):
    """Absolute (y, x) origin of every frame's crop in the (zero-padded) de-gamma'd source.

    Derived from :436-457; this is what a device kernel needs instead of nested crops.
    """
    up, jitter, smalljitter = params["upscale"], params["jitter"], params["smalljitter"]
    j_up = jitter * up
    delta_up = (jitter - smalljitter) * up
    cy, cx = draws["crop0"]
    out = [(cy + j_up, cx + j_up)]
    for k in range(params["BURST_LENGTH"] - 1):
        oy, ox = draws["frame_off"][k]
        base = 0 if draws["use_big"][k] else delta_up
        out.append((cy + base + oy, cx + base + ox))
    return out
