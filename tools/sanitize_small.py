#!/usr/bin/env python
"""A small pass over every kernel of the hot path, meant to run under `compute-sanitizer --tool memcheck`:
Simplemodel (2 x 40 x 48: raster + dense layers, slab and exchange wide-N kernels, split-K, tcgen05 filter),
Basis_kpn (1 x 64 x 64, T = 8, B = 50), the eval metrics + SSIM, and the uint8 preprocessing."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imageenhancement_mp_b200 import synth, weights, model_library as ml, data_utils as du
dev = torch.device("cuda", 0)
P = dict(synth.DEFAULT_PARAMS, graph_max_pixels=0)
m = ml.Simplemodel(P, weights=weights.init_weights(weights.simplemodel_layers(P), scheme="stress"), device=dev)
x, t = synth.make_batch(2, 40, 48, P)
out = m(x.to(dev))[0]
print("simple", float(out.abs().mean()), du.eval_metrics(out, x.to(dev), t.to(dev), 4)["psnr"])
print("ssim", float(du.ssim_deblur_sums(out, t.to(dev)).sum()))
P2 = dict(P, BURST_LENGTH=8, layer_type="dualparams", Basis_num=50)
m2 = ml.Basis_kpn(P2, weights=weights.init_weights(weights.basis_kpn_layers(P2), scheme="stress"), device=dev)
x2, t2 = synth.make_batch(1, 64, 64, P2)
print("basis_kpn", float(m2(x2.to(dev))[0].abs().mean()))
imgs = torch.randint(0, 256, (3, 300, 340, 1), dtype=torch.uint8)
for xb, tb in du.val_batches_from_u8(imgs, dict(P, height=24, width=32), batch_size=3, device=dev):
    print("preprocess", tuple(xb.shape), float(xb.mean()))
torch.cuda.synchronize()
print("done")
