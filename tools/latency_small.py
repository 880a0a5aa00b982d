#!/usr/bin/env python
"""Latency of small batches (eval.py's default: one 32x32 patch per step; BASELINE configs[0]: 32 patches of 100x100):
eager launches vs CUDA-graph replay, to place Engine.graph_max_pixels at the crossover."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imageenhancement_mp_b200 import synth, weights, model_library as ml
dev = torch.device("cuda", 0)
params = dict(synth.DEFAULT_PARAMS)
W = weights.init_weights(weights.simplemodel_layers(params))
for name, p in (("eager", dict(params, graph_max_pixels=0)), ("graph", dict(params, graph_max_pixels=1 << 30))):
    model = ml.Simplemodel(p, weights=W, device=dev)
    for n, h, w in ((1, 32, 32), (4, 64, 64), (8, 100, 100), (16, 100, 100), (32, 100, 100), (64, 100, 100), (128, 100, 100)):
        x = synth.make_batch(n, h, w, params)[0].to(dev)
        for _ in range(5):
            model(x)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(50):
            model(x)
        torch.cuda.synchronize()
        print(f"{name}: {n}x{h}x{w}: {(time.perf_counter() - t0) / 50 * 1e3:.3f} ms per forward", flush=True)
