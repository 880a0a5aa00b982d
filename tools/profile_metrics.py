#!/usr/bin/env python
"""One warm-up call, then ONE profiled call of each per-pixel kernel of BASELINE configs[4] on 4K pairs.

    ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/metrics \
        python tools/profile_metrics.py [--n 4]
Without ncu it just runs (the plain run that has to exit 0 first).
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imageenhancement_mp_b200 import _lib, data_utils as du, synth
from imageenhancement_mp_b200._lib import call, ptr, stream

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=4)
ap.add_argument("--h", type=int, default=2160); ap.add_argument("--w", type=int, default=3840)
ap.add_argument("--only", nargs="*", default=None)
a = ap.parse_args()
dev = torch.device("cuda", 0)
n, h, w, T = a.n, a.h, a.w, 4
g = torch.Generator(device=dev).manual_seed(1)
truth = torch.rand(n, h, w, device=dev, generator=g)
pred = (truth + 0.03 * torch.randn(n, h, w, device=dev, generator=g)).clamp_(0, 1)
s1 = torch.zeros(n, dtype=torch.float64, device=dev); s2 = torch.zeros(2, dtype=torch.float64, device=dev)
wl = torch.full((n,), 0.5, device=dev)
inv = torch.empty(n, h - 16, w - 16, device=dev)
recon = torch.rand(n, h, w, T + 1, device=dev, generator=g)
burst = torch.rand(n, h, w, T + 1, device=dev, generator=g)
tr2 = torch.rand(n, h, w, 2, device=dev, generator=g)
sums = torch.zeros(n, 2 * T + 4, dtype=torch.float64, device=dev)
np_ = min(n, 2)
src = torch.randint(0, 256, (np_, 4 * h + 8, 4 * w + 16, 1), dtype=torch.uint8, device=dev, generator=g)
params = dict(synth.DEFAULT_PARAMS, height=h, width=w)
org = torch.zeros(np_, T, 2, dtype=torch.int32, device=dev)
org[:, 1:] = torch.randint(0, 9, (np_, T - 1, 2), dtype=torch.int32, device=dev, generator=g)
one = torch.full((np_,), 0.5, device=dev)
nr = torch.randn(np_, h, w, T, device=dev, generator=g); ns = torch.randn(np_, h, w, T, device=dev, generator=g)

kernels = {
    "psnr": lambda: call("ie_sqdiff_sum_f32", ptr(pred), ptr(truth), n, h * w, ptr(s1), stream()),
    "img_loss": lambda: call("ie_img_loss_sums_f32", ptr(pred), ptr(truth), n, h, w, ptr(s2), stream()),
    "ssim": lambda: call("ie_ssim_f32", ptr(pred), ptr(truth), n, h, w, ptr(s1), stream()),
    "invert": lambda: call("ie_invert_preproc_f32", ptr(pred), 1, 0, 1, ptr(wl), n, h, w, 8, ptr(inv), stream()),
    "eval_metrics": lambda: call("ie_eval_metrics_f32", ptr(recon), ptr(burst), T + 1, ptr(tr2), ptr(wl), n, h, w, T, 8,
                                 ptr(sums), stream()),
    "preprocess": lambda: du.preprocess_image(src, org, params, one, one * 0.01, one * 0.05, nr, ns),
}
names = a.only or list(kernels)
for k in names:
    kernels[k]()
torch.cuda.synchronize()
torch.cuda.profiler.start()
for k in names:
    kernels[k]()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", names)
