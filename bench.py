#!/usr/bin/env python
"""Headline benchmark: megapixels/s enhanced (Simplemodel forward + PSNR/SSIM-style eval metrics).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config cfg1|cfg2|cfg3|cfg4]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one pass of the hot path over one batch of synthetic input: model forward
(model_library.Simplemodel) + the fused eval metrics of eval.py:144-182 + SSIM + the all-reduce of the
metric totals.  Workload at N=1 is BASELINE.json configs[1]: batch 256 of 100x100 patches
(T=4, singlestd -> 5 channels; computed at 104x104 because the network needs multiples of 8,
pixels counted at 100x100).  cfg1 / cfg2 are weak scaling (every rank its own batch); cfg3
(configs[2]: ONE batch of 64 1280x720 images) and cfg4 (configs[3]: 8 photos of 3264x2448) are STRONG
scaling - the batch is split over the ranks - and start from uint8 source frames through the device
preprocessing (`ie_preprocess_u8_rng`: crops, 4x AREA down-sample, white level, Philox noise), as
data_utils.py:387-394 / 198-265 do on the CPU.

`value`   : inputs already resident in HBM, CUDA-event timed over exactly K steps, max over ranks; the
            timed loop is clean (no per-kernel events).
`sustained`: the same step looped for >= 3 s, with the median SM clock sampled in that region.
`e2e`     : the same step through the public API from pinned HOST buffers (H2D of inputs + D2H of the
            metric totals inside the timed region).
`roofline`: the convolution kernels (every conv launch of a step): algorithmic conv FLOPs (SURVEY.md
            section 8d) / summed CUDA-event kernel time - measured in a SEPARATE instrumented pass
            after the timed loop - against the measured sustained bf16 peak (burst fraction alongside).
`extra`   : (N=1, cfg2 only) short runs of cfg3 / cfg4 from uint8 and the metric-kernel HBM fractions on
            4K pairs (BASELINE configs[2..4]) so the driver's one line carries them.
`cpu_baseline` / `--impl reference`: the torch-CPU oracle port of the reference (TensorFlow is not
            installable here, so the reference itself cannot run) on the box's host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

from imageenhancement_mp_b200 import synth, weights  # noqa: E402

METRIC = "megapixels/sec enhanced (fwd+PSNR/SSIM)"
UNIT = "MP/s"

# name: images per step, H, W, scaling, from_u8, micro-batch per forward, description
#   weak:   every rank processes `images` per step;  strong: `images` per step are split over the ranks
CONFIGS = {
    "cfg1": dict(images=32, h=100, w=100, scaling="weak", u8=False, micro=32,
                 desc="configs[0]: 32x100x100x5 patches (reference CPU case)"),
    "cfg2": dict(images=256, h=100, w=100, scaling="weak", u8=False, micro=256,
                 desc="configs[1]: batch 256 of 100x100 patches, bf16 trunk"),
    "cfg3": dict(images=64, h=720, w=1280, scaling="strong", u8=True, micro=8,
                 desc="configs[2]: ONE batch of 64 1280x720 images from uint8 frames, split over the ranks, micro-batches of 8"),
    "cfg4": dict(images=8, h=2448, w=3264, scaling="strong", u8=True, micro=1,
                 desc="configs[3]: 8 photos of 3264x2448 from uint8 frames, image-sharded over the ranks"),
}
NETWORK = "Simplemodel T=4 K=15 B=10 singlestd, glorot init"
PRECISION = ("bf16 operands / fp32 accumulation in the convolutions, fp32 softmaxes and metrics, "
             "fp16 coefficient / basis operands, fp32 burst and accumulation in the per-pixel filter (tcgen05)")


def build_config(cfg_name, world):
    """The `config` object of the JSON line - identical for both arms (`--impl ours` / `--impl reference`)."""
    c = CONFIGS[cfg_name]
    h, w = c["h"], c["w"]
    per_step = c["images"] * world if c["scaling"] == "weak" else c["images"]
    per_gpu = c["images"] if c["scaling"] == "weak" else -(-c["images"] // world)
    T = synth.DEFAULT_PARAMS["BURST_LENGTH"]
    return {"workload": c["desc"], "images_per_gpu_per_step": per_gpu, "images_per_step": per_step, "image": [h, w],
            "computed_at": [-(-h // 8) * 8, -(-w // 8) * 8], "channels": T + 1, "network": NETWORK, "precision": PRECISION,
            "parallelism": f"image-sharded x{world}", "input": "uint8 source frames -> device preprocessing" if c["u8"]
            else "fp32 NHWC bursts",
            "l2": "inputs of successive steps rotate through distinct buffers and every layer's activations exceed the "
                  "126 MB L2: inputs larger than L2"}


def read_conv_traffic(cfg_name):
    """DRAM bytes of the conv launches of one step, from the committed ncu capture (profiles/r*_conv_traffic.json)."""
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_conv_traffic.json")))
    if cfg_name != "cfg2" or not files:
        return None, None
    d = json.load(open(files[-1]))
    return d.get("conv_dram_bytes_per_step"), os.path.relpath(files[-1], ROOT)


def read_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"bf16_sustained": p.get("bf16_tflops_sustained", 1393.7), "bf16_burst": p.get("bf16_tflops", 1667.0),
                "hbm": p.get("hbm_gbs", 6553.0), "source": "measured"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms.

    nvidia-smi needs up to a second to enumerate an 8-GPU box before its first line: `wait_first` blocks until a
    sample has arrived.  `mark()` returns the current sample index; `summary(a, b)` summarises the samples taken
    between two marks (all samples if the region was too short to catch one)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        import threading
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.idx}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
        except Exception:
            self.proc = None
            return
        def pump():
            for line in self.proc.stdout:
                self.lines.append(line)
        threading.Thread(target=pump, daemon=True).start()

    def wait_first(self, timeout=8.0):
        t0 = time.perf_counter()
        while self.proc is not None and not self.lines and time.perf_counter() - t0 < timeout:
            time.sleep(0.02)

    def mark(self):
        return len(self.lines)

    def settle(self):
        time.sleep(0.12)                       # one more sampling period: the last line covers the end of the region

    def stop(self):
        if self.proc is not None:
            self.proc.terminate()

    def summary(self, a=0, b=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        lines = self.lines[a:b] or self.lines
        sm, mx, pw, reasons = [], [], [], set()
        for line in lines:
            f = [s.strip() for s in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w": statistics.median(pw) if pw else None}


# ---------------------------------------------------------------------------------- CPU arm
def cpu_oracle_rate(n, h, w, params, W, repeats=1, min_seconds=0.0, chunk=32):
    """Forward + eval metrics of the oracle port on all host cores, the n-image sample in chunks of ``chunk`` images.

    Runs the sample ``repeats`` times, then keeps repeating until ``min_seconds`` of CPU work were timed;
    returns (MP/s over everything timed, seconds timed, cores, passes)."""
    import oracle
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    x, truth = synth.make_batch(n, h, w, params)
    xp, _ = synth.pad_to_multiple(x, 8)
    total, passes = 0.0, 0
    while passes < repeats or total < min_seconds:
        t0 = time.perf_counter()
        with torch.no_grad():
            for c0 in range(0, n, chunk):
                out = oracle.simplemodel_forward(W, params, xp[c0:c0 + chunk])[0][:, :h, :w]
                oracle.eval_step(out, x[c0:c0 + chunk], truth[c0:c0 + chunk], params["BURST_LENGTH"])
        total += time.perf_counter() - t0
        passes += 1
    return passes * n * h * w / 1e6 / total, total, cores, passes


def run_reference(args, cfg_name):
    """The reference's CPU implementation of the path (the oracle port: TensorFlow is not installable here) on all
    host cores.  A step is the WHOLE per-GPU batch of the configuration when that is a patch batch (cfg1 / cfg2: 256
    images = ~6 s on 16 cores); for the full-resolution configurations it is ONE image (a 1280x720 forward is ~2 s of
    CPU), stated in cpu_baseline.sample.  MP/s is size-normalised either way."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    c = CONFIGS[cfg_name]
    h, w = c["h"], c["w"]
    params = dict(synth.DEFAULT_PARAMS)
    W = weights.init_weights(weights.simplemodel_layers(params))
    sample_n = args.ref_sample or (c["images"] if h * w <= 128 * 128 else 1)
    chunk = 32 if h * w <= 128 * 128 else 1
    for _ in range(args.warmup):
        cpu_oracle_rate(min(sample_n, 4), min(h, 104), min(w, 104), params, W)
    times = []
    for _ in range(args.steps):
        _, dt, cores, _ = cpu_oracle_rate(sample_n, h, w, params, W, chunk=chunk)
        times.append(dt)
    total = sum(times)
    value = args.steps * sample_n * h * w / 1e6 / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True,
        "scaling": c["scaling"], "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": build_config(cfg_name, world),
        "reference_note": "torch-CPU oracle port of the reference (TensorFlow is not installable here)",
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample_n} image(s) of {h}x{w} per step (forward + eval metrics) on rank 0's host "
                                   f"cores, {args.steps} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---------------------------------------------------------------------------------- GPU arm
def u8_sources(n, h, w, params, seed):
    """Synthetic decoded frames for n output images of h x w: uint8 [n, h*up + 2*jitter*up, w*up + 2*jitter*up, 1] on
    the host (pinned) - the size of the reference's first random crop (data_utils.py:432-439) - plus the per-image
    random draws of preprocess_image (host, tiny)."""
    from imageenhancement_mp_b200 import data_utils as du
    up, jit = params["upscale"], params["jitter"]
    hs, ws = h * up + 2 * jit * up, w * up + 2 * jit * up
    g = torch.Generator().manual_seed(seed)
    src = torch.empty(n, hs, ws, 1, dtype=torch.uint8).pin_memory()
    # a cheap band-limited scene: a random 1/16-resolution image, bilinearly up-sampled, plus byte noise
    for i in range(n):
        coarse = torch.rand(1, 1, hs // 16 + 2, ws // 16 + 2, generator=g)
        img = torch.nn.functional.interpolate(coarse, size=(hs, ws), mode="bilinear", align_corners=False)[0, 0]
        src[i, :, :, 0] = (img * 200 + torch.randint(0, 56, (hs, ws), generator=g)).to(torch.uint8)
    d = du.draw_burst_params(n, (hs, ws), dict(params, height=h, width=w), generator=g)
    return src, d


class Workload:
    """One rank's share of a configuration: device-resident inputs for `value`, pinned host inputs for `e2e`."""

    def __init__(self, cfg_name, rank, world, dev, params, nrot=4):
        from imageenhancement_mp_b200 import data_utils as du
        c = CONFIGS[cfg_name]
        self.c, self.dev, self.params, self.du = c, dev, params, du
        self.h, self.w, self.T = c["h"], c["w"], params["BURST_LENGTH"]
        if c["scaling"] == "weak":
            self.n_local = c["images"]
        else:
            lo = rank * c["images"] // world
            self.n_local = (rank + 1) * c["images"] // world - lo
        self.micro = min(c["micro"], max(self.n_local, 1))
        self.pp = dict(params, height=self.h, width=self.w)
        self.host, self.devb = [], []
        if c["u8"]:
            # one set of source frames = this rank's whole share of the batch (16 bytes per output pixel: far larger
            # than L2 from one micro-batch to the next); two sets of random draws (crops, levels, noise seeds) rotate
            if self.n_local:
                src, d0 = u8_sources(self.n_local, self.h, self.w, params, seed=1234 + 17 * rank)
                src_dev = src.to(dev)
                for i in range(2):
                    d = d0 if i == 0 else du.draw_burst_params(self.n_local, tuple(src.shape[1:3]), self.pp,
                                                               generator=torch.Generator().manual_seed(99 + rank))
                    dd = {k: d[k].to(dev) for k in ("org", "white_level", "sig_read", "sig_shot")}
                    self.host.append((src, dd))
                    self.devb.append((src_dev, dd))
            self.h2d_bytes = self.host[0][0].numel() if self.host else 0
        else:
            for i in range(nrot):
                x, truth = synth.make_batch(self.n_local, self.h, self.w, params, seed=1234 + 17 * rank + i)
                self.host.append((x.pin_memory(), truth.pin_memory()))
            self.devb = [(x.to(dev), t.to(dev)) for x, t in self.host]
            self.h2d_bytes = self.host[0][0].numel() * 4 + self.host[0][1].numel() * 4

    def batches(self, buf, seed=0):
        """The (x, truth) device micro-batches of one step from buffer set ``buf`` (device or pinned host tensors)."""
        if not self.c["u8"]:
            yield buf
            return
        src, dd = buf
        for m0 in range(0, self.n_local, self.micro):
            m1 = min(m0 + self.micro, self.n_local)
            s = src[m0:m1]
            if not s.is_cuda:
                s = s.to(self.dev, non_blocking=True)
            yield self.du.preprocess_image(s, dd["org"][m0:m1], self.pp, dd["white_level"][m0:m1], dd["sig_read"][m0:m1],
                                           dd["sig_shot"][m0:m1], seed=seed * 1000003 + m0)


def run_ours(args, cfg_name):
    from imageenhancement_mp_b200 import _lib, data_utils as du, dist as idist, model_library as ml, ops
    from imageenhancement_mp_b200 import eval as ieval
    import torch.distributed as dist

    rank, world, local = idist.init_from_env()
    assert torch.cuda.is_available(), "bench.py needs a GPU (the hot path has no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    _lib.load()
    c = CONFIGS[cfg_name]
    h, w = c["h"], c["w"]
    params = dict(synth.DEFAULT_PARAMS)
    T = params["BURST_LENGTH"]
    layers = weights.simplemodel_layers(params)
    W = weights.init_weights(layers)                       # Keras default init, seed 1234
    model = ml.Simplemodel(params, weights=W, device=dev)
    wl_ = Workload(cfg_name, rank, world, dev, params)
    nbuf = len(wl_.devb)

    def step_on(work, buf, seed=0):
        """One step of the hot path on this rank: forward + fused metrics + SSIM per micro-batch, one all-reduce."""
        tot = None
        for xb, tb in work.batches(buf, seed):
            out = model(xb)[0]
            wl = du.white_level_of(tb)
            # SSIM (BASELINE metric "fwd+PSNR/SSIM"; an extension - the reference's eval.py reports PSNR and losses only)
            sums, ssim_sums = du.eval_metric_sums_with_ssim(out, xb, tb, T, white_noise=wl)
            t = du.reduce_metric_sums(sums, work.h, work.w, T, ssim_sums=ssim_sums)
            tot = t if tot is None else tot.add_(t)
        if tot is None:                                    # strong scaling with fewer images than ranks
            tot = torch.zeros(T + 7, dtype=torch.float64, device=dev)
        idist.all_reduce_totals(tot)                       # the one collective of the step
        return tot

    step = lambda i: step_on(wl_, wl_.devb[i % nbuf] if nbuf else None, seed=i)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(fn, k):
        """Exactly k calls between two events, a barrier + device sync on both sides; returns ms (this rank)."""
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            r = fn(i)
        e1.record()
        sync_all()
        return e0.elapsed_time(e1), r

    # ---- resident-input timing: W warm-up steps, then exactly K clean steps
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step(i)
    if rank == 0:
        sampler.wait_first()
    sync_all()
    m0 = sampler.mark()
    _lib.LAUNCHES.clear()
    ms, tot = timed(step, args.steps)
    launches = sum(_lib.LAUNCHES.values())
    report = du.totals_to_report(tot.cpu(), T)

    # ---- roofline pass (separate, instrumented): CUDA events around every convolution launch
    # (with the serial schedule: when the basis branch runs on its second stream, a bracket around one of its launches
    #  also holds whatever the decoder runs beside it - the sum of brackets would count that time twice)
    n_rf = max(1, min(args.steps, 5))
    # and with eager launches: a replayed CUDA graph (small configurations) has no per-launch brackets at all
    overlap, graph_px = model._engine.overlap_branches, model._engine.graph_max_pixels
    model._engine.overlap_branches = False
    model._engine.graph_max_pixels = 0
    ops.CONV_EVENTS = []
    for i in range(n_rf):
        step(i)
    torch.cuda.synchronize()
    conv_ms = sum(a.elapsed_time(b) for a, b in ops.CONV_EVENTS) / n_rf
    n_conv = len(ops.CONV_EVENTS) // n_rf
    ops.CONV_EVENTS = None

    if args.breakdown and rank == 0:
        _lib.TRACE = []
        for i in range(3):
            step(i)
        torch.cuda.synchronize()
        agg = {}
        for name, a, b in _lib.TRACE:
            t = agg.setdefault(name, [0, 0.0])
            t[0] += 1
            t[1] += a.elapsed_time(b)
        _lib.TRACE = None
        tsum = sum(v[1] for v in agg.values()) / 3
        for name, (cnt, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            print(f"breakdown: {name:34s} {cnt // 3:3d} launches/step {t / 3:8.3f} ms/step {100 * t / 3 / tsum:5.1f}%", file=sys.stderr)
        print(f"breakdown: sum of kernel times {tsum:.3f} ms/step (events around every launch; includes launch gaps "
              f"inside each bracket; serial schedule)", file=sys.stderr)
    model._engine.overlap_branches, model._engine.graph_max_pixels = overlap, graph_px

    # ---- end-to-end timing from pinned host buffers through the public API.  fp32 configurations: eval.evaluate()
    # stages every batch host->device on a side stream (overlapping the previous step), runs forward + fused metrics,
    # and reads every step's metric totals back to pinned host memory.  uint8 configurations: the source frames are
    # copied host->device per micro-batch, preprocessed on the device, and the totals read back per step.
    if c["u8"]:
        # the public API for decoded uint8 frames: data_utils.val_batches_from_u8 (get_val_ds, data_utils.py:387-394)
        # feeding eval.evaluate - source frames cross PCIe on a side stream one micro-batch ahead, are preprocessed
        # on the device, and every micro-batch's metric totals are read back to pinned host memory
        def u8_batches(k):
            for i in range(k):
                if wl_.n_local:
                    yield from du.val_batches_from_u8(wl_.host[0][0], wl_.pp, batch_size=wl_.micro, seed=1000 + i,
                                                      device=dev, shuffle=False)
        ieval.evaluate(model, u8_batches(1), wl_.pp, out=None, step_results=[], pre_sharded=True, ssim=True)
        res = []
        def e2e_all(_):
            return ieval.evaluate(model, u8_batches(args.steps), wl_.pp, out=None, step_results=res, pre_sharded=True,
                                  ssim=True)
        e2e_ms, e2e_report = timed(e2e_all, 1)
        assert abs(e2e_report["count"] - c["images"] * args.steps) < 0.5, (e2e_report["count"], c["images"], args.steps)
        d2h = int(res[0].numel() * 8) * max(1, len(res) // args.steps) if res else 0
    else:
        def host_batches(k):
            for i in range(k):
                yield wl_.host[i % nbuf]
        ieval.evaluate(model, host_batches(max(2, args.warmup // 2)), params, out=None, step_results=[], pre_sharded=True,
                       ssim=True)
        res = []
        def e2e_all(_):
            return ieval.evaluate(model, host_batches(args.steps), params, out=None, step_results=res, pre_sharded=True,
                                  ssim=True)
        e2e_ms, e2e_report = timed(e2e_all, 1)
        assert len(res) == args.steps and abs(e2e_report["count"] - world * c["images"] * args.steps) < 0.5
        d2h = int(res[0].numel() * 8)
    m1 = sampler.mark()

    # ---- sustained leg: the same step for >= 3 s (the K-step region above is a fraction of a second: burst clocks)
    sustained = None
    if args.sustain > 0:
        k_s = max(args.steps, int(args.sustain * 1e3 / max(ms / args.steps, 1e-3)) + 1)
        sus_ms, _ = timed(step, k_s)
        sampler.settle()
        m2 = sampler.mark()
        sustained = (k_s, sus_ms, m1, m2)
    clocks = sampler.summary(m0, m1) if rank == 0 else None
    sus_clocks = sampler.summary(sustained[2], sustained[3]) if (rank == 0 and sustained) else None

    tv = torch.tensor([ms, e2e_ms, conv_ms, sustained[1] if sustained else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tv, op=dist.ReduceOp.MAX)
    ms, e2e_ms, conv_ms, sus_ms = [float(v) for v in tv.cpu()]

    extra = None
    if world == 1 and cfg_name == "cfg2" and not args.no_extra:
        extra = run_extras(model, params, dev, step_on)
    if rank == 0:
        sampler.stop()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    cfg = build_config(cfg_name, world)
    mp_per_step = cfg["images_per_step"] * h * w / 1e6
    value = mp_per_step * args.steps / (ms / 1e3)
    e2e = mp_per_step * args.steps / (e2e_ms / 1e3)
    peaks = read_peaks()
    hp, wp = -(-h // 8) * 8, -(-w // 8) * 8
    per_px, per_img = weights.conv_flops(layers, params, hp, wp)
    conv_flops_step = wl_.n_local * (per_px * hp * wp + per_img)     # rank 0's share, at the computed (padded) size
    achieved = conv_flops_step / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    px_user, _ = weights.conv_flops(layers, params, h, w)
    achieved_user = wl_.n_local * (px_user * h * w + per_img) / (conv_ms / 1e3) / 1e12 if conv_ms > 0 else 0.0
    traffic, traffic_src = read_conv_traffic(cfg_name)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": c["scaling"], "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": cfg, "clocks": clocks,
        "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": wl_.h2d_bytes, "d2h_bytes_per_step": d2h,
                "ms_per_step": e2e_ms / args.steps},
        "gpu_launches": launches,
        "roofline": {"kernel": "tcgen05 conv kernels: conv_stream / conv_wide / conv_resident / conv_first "
                               "(all %d launches of a step; achieved and traffic are per step)" % n_conv,
                     "bound": "tensor", "achieved": achieved, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                     "frac": achieved / peaks["bf16_sustained"], "traffic": traffic,
                     "traffic_note": ("DRAM bytes (read+write) of the same launches of one step, ncu --set full: " + traffic_src)
                     if traffic_src else "no ncu capture for this config",
                     "peak_source": peaks["source"] + " sustained bf16", "frac_of_burst": achieved / peaks["bf16_burst"],
                     "conv_ms_per_step": conv_ms, "conv_tflop_per_step": conv_flops_step / 1e12,
                     "timing": f"CUDA events around each conv launch in a separate pass of {n_rf} steps after the timed loop",
                     "step_tflops": conv_flops_step / (ms / args.steps / 1e3) / 1e12,
                     "flops_counted_at": [hp, wp],
                     "frac_at_user_pixels": achieved_user / peaks["bf16_sustained"]},
        "quality": {"psnr": report["psnr"], "psnr_noise0": report["psnr_noise0"], "psnr_average": report["psnr_average"],
                    "ssim": report["ssim"]},
    }
    if sustained:
        k_s = sustained[0]
        line["sustained"] = {"seconds": sus_ms / 1e3, "steps": k_s, "ms_per_step": sus_ms / k_s,
                             "value": mp_per_step * k_s / (sus_ms / 1e3), "unit": UNIT, "clocks": sus_clocks,
                             "step_tflops": conv_flops_step / (sus_ms / k_s / 1e3) / 1e12}
    if extra is not None:
        line["extra"] = extra
    if world == 1 and not args.no_cpu_baseline:
        sample_n = 16 if h * w <= 128 * 128 else 1
        cpu_oracle_rate(2 if sample_n > 1 else 1, min(h, 104), min(w, 104), params, W)          # warm the CPU path up
        v, dt, cores, passes = cpu_oracle_rate(sample_n, h, w, params, W, min_seconds=12.0)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"{passes} passes over {sample_n} image(s) of {h}x{w}, forward + eval metrics, "
                                          f"torch-CPU oracle port on all host cores ({dt:.1f} s of CPU work)"}
    emit(line)


def run_extras(model, params, dev, step_on):
    """Short single-GPU runs of the other BASELINE configurations, for the driver's one line: cfg3 (one micro-batch of
    8 1280x720 images) and cfg4 (one 3264x2448 photo), both from uint8 source frames through the device
    preprocessing, and the HBM fraction of the metric / preprocessing kernels on 4K pairs (configs[4])."""
    from imageenhancement_mp_b200 import data_utils as du
    from imageenhancement_mp_b200._lib import call, ptr, stream
    peaks = read_peaks()
    out = {}
    for name, n_img in (("cfg3", 8), ("cfg4", 1)):
        c = dict(CONFIGS[name], images=n_img, scaling="weak")
        saved = CONFIGS[name]
        CONFIGS[name] = c
        try:
            work = Workload(name, 0, 1, dev, params)
        finally:
            CONFIGS[name] = saved
        fn = lambda i: step_on(work, work.devb[i % 2], seed=i)
        for i in range(2):
            fn(i)
        torch.cuda.synchronize()
        k = 4
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(k):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / k
        out[name + "_short"] = {"workload": f"{n_img} image(s) of {c['h']}x{c['w']} per step from uint8 frames "
                                            f"(preprocess + forward + metrics), {k} steps after 2 warm-up",
                                "value": n_img * c["h"] * c["w"] / 1e6 / (ms / 1e3), "unit": UNIT, "ms_per_step": ms,
                                "u8_bytes_per_step": int(work.h2d_bytes)}
        del work
        model._engine._plans.clear()
        torch.cuda.empty_cache()
    # ---- metric kernels on 4K pairs, batch 8 (preprocess: batch 4); L2 flushed between repetitions; algorithmic bytes
    hh, ww, n, T = 2160, 3840, 8, 4
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def best_ms(fn, reps=4):
        best = 1e9
        for _ in range(reps):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        return best
    g = torch.Generator(device=dev).manual_seed(5)
    px = n * hh * ww
    truth = torch.rand(n, hh, ww, device=dev, generator=g)
    pred = (truth + 0.03 * torch.randn(n, hh, ww, device=dev, generator=g)).clamp_(0, 1)
    s1 = torch.zeros(n, dtype=torch.float64, device=dev)
    s2 = torch.zeros(2, dtype=torch.float64, device=dev)
    wl = torch.full((n,), 0.5, device=dev)
    inv = torch.empty(n, hh - 16, ww - 16, device=dev)
    rows = [("psnr_pair", 8 * px, lambda: call("ie_sqdiff_sum_f32", ptr(pred), ptr(truth), n, hh * ww, ptr(s1), stream())),
            ("img_loss", 8 * px, lambda: call("ie_img_loss_sums_f32", ptr(pred), ptr(truth), n, hh, ww, ptr(s2), stream())),
            ("ssim", 8 * px, lambda: call("ie_ssim_f32", ptr(pred), ptr(truth), n, hh, ww, ptr(s1), stream())),
            ("invert_preproc", 4 * px + 4 * n * (hh - 16) * (ww - 16),
             lambda: call("ie_invert_preproc_f32", ptr(pred), 1, 0, 1, ptr(wl), n, hh, ww, 8, ptr(inv), stream()))]
    sweep = {}
    for name, nbytes, fn in rows:
        ms = best_ms(fn)
        sweep[name] = {"GBps": nbytes / ms / 1e6, "frac_hbm": nbytes / ms / 1e6 / peaks["hbm"], "ms": ms}
    del pred, inv
    recon = torch.rand(n, hh, ww, T + 1, device=dev, generator=g)
    burst = torch.rand(n, hh, ww, T + 1, device=dev, generator=g)
    tr2 = torch.rand(n, hh, ww, 2, device=dev, generator=g)
    sums = torch.zeros(n, 2 * T + 4, dtype=torch.float64, device=dev)
    ms = best_ms(lambda: call("ie_eval_metrics_f32", ptr(recon), ptr(burst), T + 1, ptr(tr2), ptr(wl), n, hh, ww, T, 8,
                              ptr(sums), stream()))
    nbytes = 4 * (2 * T + 2) * px
    sweep["eval_metrics_fused"] = {"GBps": nbytes / ms / 1e6, "frac_hbm": nbytes / ms / 1e6 / peaks["hbm"], "ms": ms}
    del recon, burst, tr2, truth
    n4 = 4
    px4 = n4 * hh * ww
    src = torch.randint(0, 256, (n4, 4 * hh + 8, 4 * ww + 16, 1), dtype=torch.uint8, device=dev, generator=g)
    pp = dict(params, height=hh, width=ww)
    org = torch.zeros(n4, T, 2, dtype=torch.int32, device=dev)
    org[:, 1:] = torch.randint(0, 9, (n4, T - 1, 2), dtype=torch.int32, device=dev, generator=g)
    one = torch.full((n4,), 0.5, device=dev)
    ms = best_ms(lambda: du.preprocess_image(src, org, pp, one, one * 0.01, one * 0.05, seed=11))
    nbytes = 16 * px4 + 4 * (T + 1 + 2) * px4          # u8 in (each source byte once) + x and truth out; noise drawn on chip
    sweep["preprocess_u8_rng"] = {"GBps": nbytes / ms / 1e6, "frac_hbm": nbytes / ms / 1e6 / peaks["hbm"], "ms": ms}
    del src
    torch.cuda.empty_cache()
    out["metric_sweep_4k"] = {"image": [hh, ww], "batch": n, "preprocess_batch": n4, "peak_GBps": peaks["hbm"],
                              "bytes": "algorithmic (SURVEY.md section 8d)", "kernels": sweep}
    return out


def emit(line):
    """The ONE JSON line of the contract, written to the real stdout (see main: fd 1 is parked on stderr while
    the benchmark runs, because NCCL prints its version banner to stdout from C)."""
    os.write(_REAL_STDOUT, (json.dumps(line) + "\n").encode())


_REAL_STDOUT = 1


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="cfg2", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the short cfg3 / cfg4 / metric-sweep sub-records")
    ap.add_argument("--sustain", type=float, default=3.0, help="seconds of the sustained leg (0 = skip)")
    ap.add_argument("--ref-sample", type=int, default=0,
                    help="--impl reference: images per step (default: the configuration's per-GPU batch for patch "
                         "configurations, one image for full-resolution ones)")
    ap.add_argument("--breakdown", action="store_true",
                    help="extra untimed pass: per-kernel CUDA-event times of one step, printed to stderr")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args, args.config)
    else:
        run_ours(args, args.config)


if __name__ == "__main__":
    main()
