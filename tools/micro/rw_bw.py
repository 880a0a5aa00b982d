#!/usr/bin/env python
"""Read-only / write-only / copy bandwidth of this GPU with the library's own kernels and torch copies (GB/s):
what a write-dominated kernel (bilinear up-sampling writes 4x what it reads) can reach."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from imageenhancement_mp_b200 import ops
dev = torch.device("cuda", 0)
def timed(fn, reps=5):
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best
n, h, w, c = 256, 104, 104, 256
r = ops.new_raster(n, h, w, c, dev); r.data.normal_()
nbytes = r.data.numel() * 2
vec = torch.randn(n, c, device=dev)
print(f"tensor {nbytes / 1e9:.2f} GB")
ms = timed(lambda: ops.broadcast_hw(vec, r.slice())); print(f"write-only  broadcast_kernel (16 B stores): {nbytes / ms / 1e6:8.0f} GB/s")
ms = timed(lambda: ops.channel_mean(r.slice()));      print(f"read-only   channel_mean_kernel        : {nbytes / ms / 1e6:8.0f} GB/s")
ms = timed(lambda: r.data.zero_());                   print(f"write-only  torch zero_                : {nbytes / ms / 1e6:8.0f} GB/s")
d2 = torch.empty_like(r.data)
ms = timed(lambda: d2.copy_(r.data));                 print(f"copy        torch copy_ (read + write) : {2 * nbytes / ms / 1e6:8.0f} GB/s")
ms = timed(lambda: r.data.sum());                     print(f"read-only   torch sum                  : {nbytes / ms / 1e6:8.0f} GB/s")
# the up-sample kernels of the cfg2 step
for (hh, cc) in ((13, 1024), (26, 512), (52, 64)):
    src = ops.new_raster(n, hh, hh, cc, dev, dense=hh <= 26); src.data.normal_()
    dst = ops.new_raster(n, 2 * hh, 2 * hh, cc, dev, dense=hh < 26)
    ms = timed(lambda: ops.upsample_bilinear(src.slice(), dst.slice(), 2))
    tot = src.data.numel() * 2 + dst.data.numel() * 2
    print(f"upsample2 {hh:3d}^2 x {cc:4d} ch: {ms * 1e3:7.1f} us  {tot / ms / 1e6:8.0f} GB/s (read + write, {dst.data.numel() * 2 / 1e6:.0f} MB written)")
x = ops.new_raster(n, 104, 104, 64, dev); x.data.normal_()
y = ops.new_raster(n, 52, 52, 64, dev)
ms = timed(lambda: ops.maxpool2(x.slice(), y.slice(), want_mean=True))
print(f"maxpool2 104^2 x 64: {ms * 1e3:7.1f} us  {(x.data.numel() + y.data.numel()) * 2 / ms / 1e6:8.0f} GB/s")
