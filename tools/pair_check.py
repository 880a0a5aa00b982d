#!/usr/bin/env python
"""A-B of the CTA-pair (cta_group::2) 64->64 wide-N convolution against the single-CTA kernel: same output, time."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imageenhancement_mp_b200 import ops, _lib
dev = torch.device("cuda")
lib = _lib.load()
ok = True
for n, h, w in [(1, 8, 8), (2, 16, 16), (3, 40, 56), (5, 24, 104), (256, 104, 104)]:
    g = torch.Generator(device=dev).manual_seed(n * h + w)
    src = ops.new_raster(n, h, w, 64, dev)
    x = torch.randn(n, h, w, 64, device=dev, generator=g)
    ops.nhwc_to_raster(x, src) if hasattr(ops, "nhwc_to_raster") else src.data.normal_()
    wt = torch.randn(3, 3, 64, 64, device=dev, generator=g) * 0.05
    wp = ops.pack_conv_weights(wt)
    b = torch.randn(64, device=dev, generator=g) * 0.1
    outs = []
    for flags in (0, 256):
        lib.ie_conv_set_mode(-1, flags)
        dst = ops.new_raster(n, h, w, 64, dev)
        dst.data.fill_(7.0)
        ops.conv2d(src.slice(), wp, b, dst.slice(), k=3)
        torch.cuda.synchronize()
        outs.append(dst.data.clone())
        best = 1e9
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                ops.conv2d(src.slice(), wp, b, dst.slice(), k=3)
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 4)
        print(f"n={n} {h}x{w} flags={flags}: {best*1e3:.1f} us", flush=True)
    lib.ie_conv_set_mode(-1, 0)
    d = (outs[0].float() - outs[1].float()).abs().max().item()
    same = torch.equal(outs[0], outs[1])
    print(f"n={n} {h}x{w}: max abs diff {d:.3e} bit-identical={same}", flush=True)
    ok &= d <= 1e-2
print("PAIR_OK" if ok else "PAIR_MISMATCH")
