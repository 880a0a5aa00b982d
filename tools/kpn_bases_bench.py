"""Per-pixel filter at Basis_kpn sizes (remote/record.txt: T = 8, B up to 90): CUDA-core fp32 kernel vs the TF32
tensor-core kernel run in chunks of 16 bases.  CUDA-event times, L2 flushed between repetitions.

    python tools/kpn_bases_bench.py [--n 16 --h 256 --w 256]
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from imageenhancement_mp_b200 import ops  # noqa: E402


def timed(fn, flush, reps=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=16)
    ap.add_argument("--h", type=int, default=256)
    ap.add_argument("--w", type=int, default=256)
    args = ap.parse_args()
    dev = torch.device("cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    g = torch.Generator(device=dev).manual_seed(3)
    for T, B in [(4, 10), (8, 10), (8, 32), (8, 90)]:
        x = torch.rand(args.n, args.h, args.w, T + 2, device=dev, generator=g)
        coef = torch.softmax(torch.randn(args.n, args.h, args.w, B, device=dev, generator=g), -1)
        bas = torch.softmax(torch.randn(args.n, 225 * T, B, device=dev, generator=g) * 3, 1).view(args.n, 15, 15, T, B)
        out32 = ops.kpn_apply(x, T, coef, bas)
        outtf = ops.kpn_apply(x, T, coef, bas, precision="tf32")
        err = float((out32 - outtf).abs().max())
        ms32 = timed(lambda: ops.kpn_apply(x, T, coef, bas), flush)
        mstf = timed(lambda: ops.kpn_apply(x, T, coef, bas, precision="tf32"), flush)
        mp = args.n * args.h * args.w / 1e6
        print(json.dumps({"T": T, "B": B, "megapixels": mp, "fp32_ms": ms32, "tf32_ms": mstf, "speedup": ms32 / mstf,
                          "tf32_mp_per_s": mp / mstf * 1e3, "max_abs_diff": err}))


if __name__ == "__main__":
    main()
