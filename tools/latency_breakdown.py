#!/usr/bin/env python
"""Per-launch CUDA-event times of ONE small forward (default: eval.py's 1 x 32 x 32 patch), eager launches."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imageenhancement_mp_b200 import _lib, synth, weights, model_library as ml
n, h, w = (int(v) for v in (sys.argv[1:4] or (1, 32, 32)))
arch = sys.argv[4] if len(sys.argv) > 4 else "simple"          # or "basis_kpn" (remote/ settings: T = 8, dualparams)
dev = torch.device("cuda", 0)
if arch == "simple":
    params = dict(synth.DEFAULT_PARAMS, graph_max_pixels=0)
    W = weights.init_weights(weights.simplemodel_layers(params))
    model = ml.Simplemodel(params, weights=W, device=dev)
else:
    params = dict(synth.DEFAULT_PARAMS, graph_max_pixels=0, BURST_LENGTH=8, layer_type="dualparams",
                  Basis_num=int(sys.argv[5]) if len(sys.argv) > 5 else 10)
    W = weights.init_weights(weights.basis_kpn_layers(params))
    model = ml.Basis_kpn(params, weights=W, device=dev)
x = synth.make_batch(n, h, w, params)[0].to(dev)
for _ in range(5):
    model(x)
torch.cuda.synchronize()
agg = None
for rep in range(5):
    _lib.TRACE = []
    model(x)
    torch.cuda.synchronize()
    t = [(name, a.elapsed_time(b) * 1e3) for name, a, b in _lib.TRACE]
    agg = t if agg is None else [(nm, min(u, v)) for (nm, u), (_, v) in zip(agg, t)]
_lib.TRACE = None
tot = 0.0
for i, (name, us) in enumerate(agg):
    print(f"{i:2d} {name:34s} {us:8.1f} us")
    tot += us
print(f"sum {tot:.1f} us over {len(agg)} launches ({n}x{h}x{w})")
