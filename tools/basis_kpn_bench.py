"""Throughput of the second entry point, Basis_kpn (model_library.py:180-295), with the remote/ settings
(running_train_remote.py:29,34 and record.txt: T = 8, dualparams, Basis_num 50 / 90, 64x64 patches): forward + fused
eval metrics per step, CUDA-event timed, inputs rotated (larger than L2), with the per-entry-point breakdown.

    python tools/basis_kpn_bench.py [--batch 256 --steps 5 --bases 10 50 90]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imageenhancement_mp_b200 import _lib, data_utils as du, model_library as ml, synth, weights  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--bases", type=int, nargs="+", default=[10, 50, 90])
    args = ap.parse_args()
    dev = torch.device("cuda")
    for B in args.bases:
        params = dict(synth.DEFAULT_PARAMS, BURST_LENGTH=8, layer_type="dualparams", Basis_num=B)
        T = params["BURST_LENGTH"]
        layers = weights.basis_kpn_layers(params)
        model = ml.Basis_kpn(params, weights=weights.init_weights(layers), device=dev)
        n, h = args.batch, args.size
        batches = [tuple(t.to(dev) for t in synth.make_batch(n, h, h, params, seed=7 + i)) for i in range(4)]

        def step(xb, tb):
            out = model(xb)[0]
            wl = du.white_level_of(tb)
            return du.reduce_metric_sums(du.eval_metric_sums(out, xb, tb, T, white_noise=wl), h, h, T)

        for i in range(3):
            step(*batches[i % 4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            tot = step(*batches[i % 4])
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / args.steps
        _lib.TRACE = []
        step(*batches[0])
        torch.cuda.synchronize()
        agg = {}
        for name, a, b in _lib.TRACE:
            agg[name] = agg.get(name, 0.0) + a.elapsed_time(b)
        _lib.TRACE = None
        rep = du.totals_to_report(tot.cpu(), T)
        print(json.dumps({"model": "Basis_kpn", "T": T, "B": B, "batch": n, "image": [h, h], "ms_per_step": ms,
                          "mp_per_s": n * h * h / 1e6 / (ms / 1e3), 
                          "breakdown_ms": {k: round(v, 3) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:6]},
                          "psnr": rep["psnr"]}))
        del model
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
