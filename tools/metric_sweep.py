#!/usr/bin/env python
"""BASELINE.json configs[4]: metric-only sweep on 4K image pairs - achieved HBM GB/s of the per-pixel kernels
vs batch size.  Algorithmic bytes per pixel: PSNR / loss / SSIM pair 8 B, invert_preproc
4 B in + 4 B out, fused eval metrics 4*(2T+2) = 40 B (T=4), preprocess 16*C B in (4x AREA, C=1) + 4*(T+add+2) out.

    python tools/metric_sweep.py [--h 2160 --w 3840] [--batches 1 2 4 8] > profiles/rNN_metric_sweep.jsonl
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imageenhancement_mp_b200 import _lib, data_utils as du, synth
from imageenhancement_mp_b200._lib import call, ptr, stream

ap = argparse.ArgumentParser()
ap.add_argument("--h", type=int, default=2160); ap.add_argument("--w", type=int, default=3840)
ap.add_argument("--batches", type=int, nargs="+", default=[1, 2, 4, 8, 16])
ap.add_argument("--reps", type=int, default=5)
ap.add_argument("--variants", action="store_true", help="also time the A-B variants (legacy kernels, rows per batch)")
a = ap.parse_args()
dev = torch.device("cuda", 0)
_lib.load()
peak = 6553.0
pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
if os.path.exists(pk):
    peak = json.load(open(pk)).get("hbm_gbs", peak)
T = 4
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > L2 (126 MB)


def timed(fn):
    best = 1e9
    for _ in range(a.reps):
        flush.zero_()                                               # evict the operands from L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


for n in a.batches:
    h, w = a.h, a.w
    px = n * h * w
    g = torch.Generator(device=dev).manual_seed(n)
    truth = torch.rand(n, h, w, device=dev, generator=g)
    pred = (truth + 0.03 * torch.randn(n, h, w, device=dev, generator=g)).clamp_(0, 1)
    sums1 = torch.zeros(n, dtype=torch.float64, device=dev)
    sums2 = torch.zeros(2, dtype=torch.float64, device=dev)
    wl = torch.full((n,), 0.5, device=dev)
    inv = torch.empty(n, h - 16, w - 16, device=dev)
    rows = []
    rows.append(("psnr_pair(ie_sqdiff_sum_f32)", 8 * px,
                 timed(lambda: call("ie_sqdiff_sum_f32", ptr(pred), ptr(truth), n, h * w, ptr(sums1), stream()))))
    rows.append(("img_loss(ie_img_loss_sums_f32)", 8 * px,
                 timed(lambda: call("ie_img_loss_sums_f32", ptr(pred), ptr(truth), n, h, w, ptr(sums2), stream()))))
    rows.append(("ssim(ie_ssim_f32)", 8 * px,
                 timed(lambda: call("ie_ssim_f32", ptr(pred), ptr(truth), n, h, w, ptr(sums1), stream()))))
    if a.variants:
        for tag, knob in (("legacy-1col", 1),):
            _lib.load().ie_ssim_tune(knob)
            rows.append((f"ssim[{tag}]", 8 * px,
                         timed(lambda: call("ie_ssim_f32", ptr(pred), ptr(truth), n, h, w, ptr(sums1), stream()))))
        _lib.load().ie_ssim_tune(0)
    rows.append(("invert_preproc(ie_invert_preproc_f32)", 4 * px + 4 * n * (h - 16) * (w - 16),
                 timed(lambda: call("ie_invert_preproc_f32", ptr(pred), 1, 0, 1, ptr(wl), n, h, w, 8, ptr(inv), stream()))))
    if n <= 8:
        recon = torch.rand(n, h, w, T + 1, device=dev, generator=g)
        burst = torch.rand(n, h, w, T + 1, device=dev, generator=g)
        tr2 = torch.rand(n, h, w, 2, device=dev, generator=g)
        sums = torch.zeros(n, 2 * T + 4, dtype=torch.float64, device=dev)
        em = lambda: call("ie_eval_metrics_f32", ptr(recon), ptr(burst), T + 1, ptr(tr2), ptr(wl), n, h, w, T, 8,
                          ptr(sums), stream())
        rows.append(("eval_metrics_fused(ie_eval_metrics_f32,T=4)", 4 * (2 * T + 2) * px, timed(em)))
        if a.variants:
            lib = _lib.load()
            for tag, knobs in (("legacy-tile", (0, 0, 1)), ("rb1", (1, 4, 0)), ("rb4", (4, 4, 0)), ("rb8", (8, 4, 0)),
                               ("rb2-3warps", (2, 3, 0))):
                lib.ie_eval_metrics_tune(*knobs)
                rows.append((f"eval_metrics_fused[{tag}]", 4 * (2 * T + 2) * px, timed(em)))
            lib.ie_eval_metrics_tune(0, 0, 0)
        del recon, burst, tr2
    if n <= 4:
        # preprocess: u8 [n, 4h, 4w, 1] -> x [n,h,w,5], truth [n,h,w,2]   (4x AREA down-sample, T=4 frames)
        src = torch.randint(0, 256, (n, 4 * h + 8, 4 * w + 16, 1), dtype=torch.uint8, device=dev, generator=g)
        params = dict(synth.DEFAULT_PARAMS, height=h, width=w)
        org = torch.zeros(n, T, 2, dtype=torch.int32, device=dev)
        one = torch.full((n,), 0.5, device=dev)
        nr = torch.randn(n, h, w, T, device=dev, generator=g)
        ns = torch.randn(n, h, w, T, device=dev, generator=g)
        nbytes = 16 * px + 4 * (T + 1 + 2) * px + 2 * 4 * T * px      # u8 in (read once) + outputs + noise inputs
        org[:, 1:] = torch.randint(0, 9, (n, T - 1, 2), dtype=torch.int32, device=dev, generator=g)   # frame jitter
        pre = lambda: du.preprocess_image(src, org, params, one, one * 0.01, one * 0.05, nr, ns)
        rows.append(("preprocess(ie_preprocess_u8,up=4,T=4)", nbytes, timed(pre)))
        if a.variants:
            lib = _lib.load()
            for tag, knob in (("legacy-1px", 1),):
                lib.ie_preprocess_tune(knob)
                rows.append((f"preprocess[{tag}]", nbytes, timed(pre)))
            lib.ie_preprocess_tune(0)
        del src, nr, ns
    for name, nbytes, ms in rows:
        gbs = nbytes / ms / 1e6
        print(json.dumps({"kernel": name, "batch": n, "image": [h, w], "ms": round(ms, 4), "algorithmic_GB": round(nbytes / 1e9, 4),
                          "GBps": round(gbs, 1), "frac_of_measured_hbm": round(gbs / peak, 3), "peak_GBps": peak}), flush=True)
    del truth, pred, inv
    torch.cuda.empty_cache()
