#!/usr/bin/env python
"""Launch the coef head (64 -> 10, fp32 softmax epilogue) a few times (for ncu captures)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imageenhancement_mp_b200 import ops, _lib
dev = torch.device("cuda")
_lib.load()
src = ops.new_raster(256, 104, 104, 64, dev); src.data.normal_()
wp = ops.pack_conv_weights(torch.randn(3, 3, 64, 10, device=dev) * 0.05, _lib.IE_EPI_F32_SOFTMAX)
b = torch.zeros(10, device=dev)
for _ in range(3):
    y, _ = ops.conv2d_f32(src.slice(), wp, b, 10, softmax=True)
torch.cuda.synchronize()
print("ok")
