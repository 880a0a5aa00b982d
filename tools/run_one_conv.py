#!/usr/bin/env python
"""Launch one conv layer a few times (for ncu captures of a single kernel flavour)."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imageenhancement_mp_b200 import ops, _lib
ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=256); ap.add_argument("--hw", type=int, default=104)
ap.add_argument("--cin", type=int, default=64); ap.add_argument("--cout", type=int, default=64)
ap.add_argument("--k", type=int, default=3); ap.add_argument("--mode", type=int, default=-1)
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda")
_lib.load().ie_conv_set_mode(a.mode, 0)
src = ops.new_raster(a.n, a.hw, a.hw, a.cin, dev); src.data.normal_()
wp = ops.pack_conv_weights(torch.randn(a.k, a.k, a.cin, a.cout, device=dev) * 0.05)
b = torch.zeros(a.cout, device=dev)
dst = ops.new_raster(a.n, a.hw, a.hw, a.cout, dev)
for _ in range(a.reps):
    ops.conv2d(src.slice(), wp, b, dst.slice(), k=a.k)
torch.cuda.synchronize()
print("ok")
