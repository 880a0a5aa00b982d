"""SSIM oracle (test infrastructure).  EXTENSION - not in the reference, parity unpinned.

The reference has no SSIM (grep ssim|gauss in /root/reference: 0 hits); BASELINE.json's
north_star asks for one.  It is defined here with ``tf.image.ssim`` semantics: 11x11
Gaussian window (sigma 1.5, normalised), VALID windows, K1=0.01, K2=0.03, max_val=1,
per-image mean of the SSIM map.  fp64 by default.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def gaussian_window(size=11, sigma=1.5, dtype=torch.float64):
    coords = torch.arange(size, dtype=dtype) - (size - 1) / 2.0
    g = torch.exp(-0.5 * coords ** 2 / sigma ** 2)
    return g / g.sum()


def ssim(a, b, max_val=1.0, size=11, sigma=1.5, k1=0.01, k2=0.03, dtype=torch.float64):
    """a, b: [N,H,W] -> per-image SSIM [N]."""
    a = a.to(dtype).unsqueeze(1)
    b = b.to(dtype).unsqueeze(1)
    g = gaussian_window(size, sigma, dtype)
    w = (g[:, None] * g[None, :]).reshape(1, 1, size, size)
    c1 = (k1 * max_val) ** 2
    c2 = (k2 * max_val) ** 2
    mu_a = F.conv2d(a, w)
    mu_b = F.conv2d(b, w)
    e_aa = F.conv2d(a * a, w)
    e_bb = F.conv2d(b * b, w)
    e_ab = F.conv2d(a * b, w)
    num0 = 2 * mu_a * mu_b
    den0 = mu_a * mu_a + mu_b * mu_b
    lum = (num0 + c1) / (den0 + c1)
    cs = (2 * e_ab - num0 + c2) / (e_aa + e_bb - den0 + c2)
    return (lum * cs).mean(dim=(1, 2, 3))
