"""Per-layer device time of one forward: every ie_conv2d_nhwc_bf16 call with its descriptor, CUDA-event timed
(warm, eager launches behind a queue of dummy work so that host call overhead is not in the brackets; with programmatic dependent launch the prologue of a layer overlaps the tail of the previous
one, so single rows are +-2 us), useful TFLOP/s per layer.

    python tools/layer_times.py [--model Simplemodel|Basis_kpn] [--batch 256] [--size 100] [--bases 10] [--burst 4]
"""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imageenhancement_mp_b200 import _lib, model_library as ml, ops, synth, weights  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="Simplemodel")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--size", type=int, default=100)
    ap.add_argument("--bases", type=int, default=10)
    ap.add_argument("--burst", type=int, default=4)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--flags", type=int, default=0, help="ie_conv_set_mode flags (include/imgenh_b200.h)")
    args = ap.parse_args()
    dev = torch.device("cuda")
    _lib.load().ie_conv_set_mode(-1, args.flags)
    if args.model == "Basis_kpn":
        params = dict(synth.DEFAULT_PARAMS, BURST_LENGTH=args.burst, layer_type="dualparams", Basis_num=args.bases)
        model = ml.Basis_kpn(params, weights=weights.init_weights(weights.basis_kpn_layers(params)), device=dev)
    else:
        params = dict(synth.DEFAULT_PARAMS, BURST_LENGTH=args.burst, Basis_num=args.bases)
        model = ml.Simplemodel(params, device=dev)
    model._engine.overlap_branches = False          # serial schedule: a bracket then holds one launch only
    x = synth.make_batch(args.batch, args.size, args.size, params, seed=3)[0].to(dev)
    for _ in range(3):
        model(x)
    rows = []
    real_call = ops.call

    def traced(name, *a):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        real_call(name, *a)
        e1.record()
        d = a[0]._obj if name == "ie_conv2d_nhwc_bf16" else None
        desc = None
        if d is not None:
            desc = (d.n_img, d.h, d.w, d.hv, d.wv, d.kh, d.cin, d.cout, d.dense, d.epilogue)
        rows.append((name, desc, e0, e1))

    acc = {}
    busy = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
    for rep in range(args.reps):
        rows.clear()
        for _ in range(4):          # ~4 ms of queued device work: the host runs ahead, so the events below
            busy @ busy             # bracket device time only, not the Python / ctypes call overhead
        ops.call = traced
        model(x)
        ops.call = real_call
        torch.cuda.synchronize()
        for i, (name, desc, e0, e1) in enumerate(rows):
            acc.setdefault(i, [name, desc, 0.0])[2] += e0.elapsed_time(e1) / args.reps
    tot = sum(v[2] for v in acc.values())
    print(f"# {args.model} batch {args.batch} of {args.size}x{args.size}, T={args.burst}, B={args.bases}: "
          f"{tot:.3f} ms over {len(acc)} calls")
    for i in sorted(acc):
        name, desc, ms = acc[i]
        if desc is None:
            print(f"{i:3d} {name:34s} {ms * 1e3:8.1f} us")
            continue
        n, h, w, hv, wv, k, cin, cout, dense, epi = desc
        fl = 2.0 * n * hv * wv * k * k * cin * cout
        print(f"{i:3d} conv {k}x{k} {cin:5d}->{cout:5d} {n}x{h}x{w}{' dense' if dense else ''} epi{epi:d}".ljust(52)
              + f"{ms * 1e3:8.1f} us {fl / ms / 1e9:8.1f} TFLOP/s")


if __name__ == "__main__":
    main()
