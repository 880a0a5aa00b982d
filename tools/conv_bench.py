#!/usr/bin/env python
"""Per-layer timing of the conv kernel on the cfg2 (or cfg3) layer shapes: CUDA events, best-of-N."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imageenhancement_mp_b200 import ops, _lib

# name, res divisor, cin, cout, k, epilogue
LAYERS = [
    ("layer0(1x1,K64)", 1, 64, 64, 1, 0), ("down1.c", 1, 64, 64, 3, 0), ("down2.c1", 2, 64, 128, 3, 0),
    ("down2.c2", 2, 128, 128, 3, 0), ("down5.c1", 4, 128, 1024, 3, 0), ("down5.c2", 4, 1024, 1024, 3, 0),
    ("layer1_1", 8, 1024, 1024, 3, 0), ("Cup1.c1", 4, 2048, 512, 3, 0), ("Cup1.c2", 4, 512, 512, 3, 0),
    ("Cup4.c1", 2, 640, 64, 3, 0), ("Cup4.c2", 2, 64, 64, 3, 0), ("Cup5.c1", 1, 128, 64, 3, 0),
    ("coef", 1, 64, 10, 3, 2),
]

def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--h", type=int, default=104)
    ap.add_argument("--w", type=int, default=104)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="")
    ap.add_argument("--mode", type=int, default=-1, help="-1 auto, 0 stream, 1 resident, 2 wide-N (forced)")
    ap.add_argument("--flags", type=int, default=0, help="ie_conv_set_mode flags (4: wide streams weights, 8: one epilogue set, "
                    "2048: exchange epilogue for one-block wide layers, 8192: no tail split-K, 16384: narrow N tiles instead of split-K on small grids)")
    ap.add_argument("--raster", action="store_true", help="shared-border rasters at every resolution (default: dense NHWC "
                    "from 1/4 resolution down, like the engine)")
    a = ap.parse_args()
    dev = torch.device("cuda")
    _lib.load().ie_conv_set_mode(a.mode, a.flags)
    out = []
    for name, div, cin, cout, k, epi in LAYERS:
        if a.only and a.only not in name:
            continue
        h, w = a.h // div, a.w // div
        dense = (not a.raster) and div >= 4
        src = ops.new_raster(a.n, h, w, cin, dev, dense=dense); src.data.normal_()
        wt = torch.randn(k, k, cin, cout, device=dev) * 0.05
        wp = ops.pack_conv_weights(wt, epi)
        b = torch.zeros(cout, device=dev)
        dst = ops.new_raster(a.n, h, w, max(cout, 64), dev, dense=dense) if epi == 0 else None
        def run():
            if epi == 0:
                ops.conv2d(src.slice(), wp, b, dst.slice(), k=k, workspace=ops.conv_workspace(dev))
            else:
                ops.conv2d_f32(src.slice(), wp, b, cout, k=k, softmax=True)
        try:
            run(); torch.cuda.synchronize()
        except Exception as ex:
            print(f"{name:16s} skipped: {ex}", flush=True)
            continue
        best = 1e9
        for _ in range(a.reps):
            # 4 back-to-back launches per measurement: the GPU is busy when the later ones are enqueued, so the
            # host-side launch latency does not leak into the device time of small kernels
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(4):
                run()
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) / 4)
        flops = 2.0 * k * k * cin * cout * a.n * h * w
        out.append((name, h, cin, cout, best, flops / best / 1e9))
        print(f"{name:16s} {h:4d}x{w:<4d} {cin:5d}->{cout:<5d} {best*1e3:9.1f} us  {flops/best/1e9:8.1f} TFLOP/s (real work)", flush=True)
    print("total ms (unweighted sum of listed):", sum(o[4] for o in out))

if __name__ == "__main__":
    main()
