"""Synthetic inputs of the shapes BASELINE.json names (there is no dataset in this image).

The statistics follow the reference's synthetic-burst generator
(/root/reference/data_utils.py:213-265): a de-gamma'd scene scaled by a white level
10^U(-1,0), read + shot noise with sigma_r = 10^U(-3,-1.5), sigma_s = 10^U(-2,-1), and a
``singlestd`` noise-level channel.  All draws come from one seeded CPU generator
(seed 1234 = the reference's, eval.py:63-64), so the CPU oracle and every GPU rank see
the same data.
"""
from __future__ import annotations

import torch

from .weights import ADD_LENGTHS

DEFAULT_PARAMS = {                    # eval.py:28-51 defaults
    "BURST_LENGTH": 4, "Kernel_size": 15, "Basis_num": 10, "regu": 0.0, "ps": False,
    "layer_type": "singlestd", "height": 32, "width": 32, "batch_size": 1,
    "degamma": 2.2, "to_shift": 1.0, "upscale": 4, "jitter": 16, "smalljitter": 2, "color": False,
}


def smooth_scene(N, H, W, g):
    """A band-limited random scene in [0,1] (bilinear-upsampled coarse noise + fine texture)."""
    coarse = torch.rand(N, 1, max(H // 8, 2), max(W // 8, 2), generator=g)
    base = torch.nn.functional.interpolate(coarse, size=(H, W), mode="bilinear", align_corners=False)
    fine = torch.rand(N, 1, H, W, generator=g)
    return (0.8 * base + 0.2 * fine)[:, 0].clamp(0, 1)


def make_batch(N, H, W, params=None, seed=1234, shifts=True):
    """Returns (x [N,H,W,T+add] f32, truth [N,H,W,2] f32) on CPU.

    x[...,0:T] is the noisy burst (frame t is the scene shifted by a small integer jitter when
    ``shifts``), x[...,T:] the noise-level channels; truth[...,0] is the clean reference
    frame, truth[...,1] the white-level map (data_utils.py:265).
    """
    p = dict(DEFAULT_PARAMS)
    if params:
        p.update(params)
    T = p["BURST_LENGTH"]
    add = ADD_LENGTHS[p["layer_type"]]
    g = torch.Generator().manual_seed(seed)
    scene = smooth_scene(N, H + 8, W + 8, g) ** 2.2
    wl = 10 ** (torch.rand(N, generator=g) * 1.0 - 1.0)
    sr = 10 ** (torch.rand(N, generator=g) * 1.5 - 3.0)
    ss = 10 ** (torch.rand(N, generator=g) * 1.0 - 2.0)
    frames = []
    for t in range(T):
        if t == 0 or not shifts:
            dy = dx = 4
        else:
            dy = int(torch.randint(2, 7, (1,), generator=g))
            dx = int(torch.randint(2, 7, (1,), generator=g))
        frames.append(scene[:, dy:dy + H, dx:dx + W])
    clean = torch.stack(frames, dim=-1) * wl.view(N, 1, 1, 1)
    n1 = torch.randn(N, H, W, T, generator=g)
    n2 = torch.randn(N, H, W, T, generator=g)
    noisy = clean + clean.sqrt() * ss.view(N, 1, 1, 1) * n1 + sr.view(N, 1, 1, 1) * n2
    if p["layer_type"] == "singlestd":
        sig = torch.sqrt(sr.view(N, 1, 1, 1) ** 2 + noisy[..., 0:1].clamp(min=0) * ss.view(N, 1, 1, 1) ** 2)
    elif p["layer_type"] == "dualparams":
        sig = torch.stack([sr, ss], dim=-1).view(N, 1, 1, 2).expand(N, H, W, 2)
    else:
        sig = noisy[..., 0:0]
    assert sig.shape[-1] == add
    x = torch.cat([noisy, sig], dim=-1).contiguous().float()
    truth = torch.stack([clean[..., 0], wl.view(N, 1, 1).expand(N, H, W)], dim=-1).contiguous().float()
    return x, truth


def pad_to_multiple(x, m):
    """Zero-pad H and W (dims 1, 2 of NHWC) up to a multiple of m; returns (padded, (H, W))."""
    H, W = x.shape[1], x.shape[2]
    Hp, Wp = -(-H // m) * m, -(-W // m) * m
    if (Hp, Wp) == (H, W):
        return x, (H, W)
    return torch.nn.functional.pad(x, (0, 0, 0, Wp - W, 0, Hp - H)), (H, W)
