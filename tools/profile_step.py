#!/usr/bin/env python
"""One warm-up step, then ONE profiled step (forward + fused metrics) of a bench config.

Run under ncu with `--profile-from-start off`: only the launches between
cudaProfilerStart/Stop are captured.  Without ncu it just runs (the plain run ncu requires first).
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imageenhancement_mp_b200 import synth, weights, model_library as ml, data_utils as du

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=256)
ap.add_argument("--h", type=int, default=100)
ap.add_argument("--w", type=int, default=100)
a = ap.parse_args()
dev = torch.device("cuda", 0)
params = dict(synth.DEFAULT_PARAMS)
T = params["BURST_LENGTH"]
model = ml.Simplemodel(params, weights=weights.init_weights(weights.simplemodel_layers(params)), device=dev)
batches = [tuple(t.to(dev) for t in synth.make_batch(a.n, a.h, a.w, params, seed=1234 + i)) for i in range(2)]

def step(xb, tb):
    out = model(xb)[0]                                  # the bench step (bench.py::run_ours.step)
    wl = du.white_level_of(tb)
    sums, ssim_sums = du.eval_metric_sums_with_ssim(out, xb, tb, T, white_noise=wl)
    return du.reduce_metric_sums(sums, a.h, a.w, T, ssim_sums=ssim_sums)

step(*batches[0]); torch.cuda.synchronize()
torch.cuda.profiler.start()
tot = step(*batches[1]); torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok", du.totals_to_report(tot.cpu(), T)["psnr"])
