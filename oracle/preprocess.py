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


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw - SC'11, Random123) on numpy uint64 arrays holding 32-bit values."""
    import numpy as np
    c = [np.asarray(v, dtype=np.uint64) for v in (c0, c1, c2, c3)]
    k0, k1 = np.uint64(k0), np.uint64(k1)
    M0, M1, MASK = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c = [hi1 ^ c[1] ^ k0, lo1, hi0 ^ c[3] ^ k1, lo0]
        k0 = (k0 + np.uint64(0x9E3779B9)) & MASK
        k1 = (k1 + np.uint64(0xBB67AE85)) & MASK
    return c


def philox_normals(seed, count):
    """numpy restatement of the device noise generator (csrc/preprocess.cu): Philox4x32-10 keyed by ``seed``,
    counter = element index, Box-Muller on (u0,u1) -> z_shot and (u2,u3) -> z_read.  Returns two float32 arrays."""
    import numpy as np
    idx = np.arange(count, dtype=np.uint64)
    zero = np.zeros(count, np.uint64)
    c = philox4x32_10(idx & np.uint64(0xFFFFFFFF), idx >> np.uint64(32), zero, zero,
                      seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)
    u = [((v >> np.uint64(8)).astype(np.float64) * 2.0 ** -24 + 2.0 ** -25) for v in c]
    z_shot = np.sqrt(-2.0 * np.log(u[0])) * np.cos(2.0 * np.pi * u[1])
    z_read = np.sqrt(-2.0 * np.log(u[2])) * np.cos(2.0 * np.pi * u[3])
    return z_shot.astype(np.float32), z_read.astype(np.float32)
