import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imageenhancement_mp_b200 import ops, _lib
from tests.test_gpu_conv import make_case, to_raster, ref_conv, bf16_round
dev = torch.device("cuda")
lib = _lib.load()
for (n, h, w, cin, cout, k) in [(2, 16, 16, 64, 64, 3), (2, 24, 40, 128, 64, 3), (2, 24, 40, 64, 64, 1), (1, 16, 16, 64, 64, 2)]:
    x, wt, b = make_case(n, h, w, cin, cout, k)
    ref = bf16_round(ref_conv(x, wt, b, k))
    src = to_raster(x.to(dev)); wp = ops.pack_conv_weights(wt.to(dev))
    for mode, bo in [(0, 0), (1, 1), (1, 0)]:
        lib.ie_conv_set_mode(mode, bo)
        dst = ops.new_raster(n, h, w, cout, dev); dst.data.zero_()
        try:
            ops.conv2d(src.slice(), wp, b.to(dev), dst.slice(), k=k, valid=(h - 1, w - 1) if k == 2 else None)
            torch.cuda.synchronize()
            got = ops.raster_to_nhwc(dst.slice()).cpu()
            if k == 2: got = got[:, :h-1, :w-1]
            print((n,h,w,cin,cout,k), "mode", mode, "base_off", bo, "max err", float((got - ref).abs().max()), "ref max", float(ref.abs().max()), flush=True)
        except Exception as e:
            print("FAILED", mode, bo, e, flush=True)
            raise
lib.ie_conv_set_mode(-1, 0)
