"""Drop-in for the metric / loss half of the reference's ``data_utils.py`` on B200.

Same names, argument order and meaning as /root/reference/data_utils.py:24-164, but on
``torch`` CUDA tensors and computed by the kernels in csrc/metrics.cu.  Two layers:

* the reference's own fine-grained functions (``invert_preproc``, ``psnr_deblur`` ...), one or two
  kernel launches each, for code written against the reference API;
* ``eval_metrics`` - everything eval.py:144-182 computes for a batch from ONE fused pass.

``preprocess_image`` is the arithmetic of ``DataLoader.preprocess_image`` (:198-265) with the
random draws passed in (file I/O, decoding and tf.data plumbing are out of scope).
"""
from __future__ import annotations

import math

import psutil
import torch

from . import _lib
from ._lib import call, ptr, stream

LBUFF = 8      # border crop of invert_preproc, data_utils.py:43
LAYER_TYPES = {"empty": 0, "singlestd": 1, "dualparams": 2}


def getMemCpu():
    """data_utils.py:17-23."""
    data = psutil.virtual_memory()
    return int(round(data.percent)), psutil.cpu_percent(interval=1)


# ------------------------------------------------------------------ helpers
def _nhw_view(img):
    """Describe an [N,H,W] fp32 CUDA tensor as (base tensor, pitch, coff) without copying when it
    is a channel slice ``y[..., k]`` of a contiguous NHWC tensor."""
    _lib.require_cuda(img)
    if img.dtype != torch.float32:
        img = img.float()
    n, h, w = img.shape
    sn, sh, sw = img.stride()
    if sw >= 1 and sh == w * sw and sn == h * sh:
        return img, sw, 0              # data_ptr() already includes the channel offset; pitch = stride
    img = img.contiguous()
    return img, 1, 0


def _wl_vec(white_level, n):
    wl = white_level.reshape(-1).float().contiguous()
    assert wl.numel() == n, "white_level must have one entry per image"
    return wl


def white_level_of(x_batch_truth):
    """eval.py:144-145: per-image mean of truth[...,1] -> [N,1,1,1]."""
    _lib.require_cuda(x_batch_truth)
    t = x_batch_truth.contiguous().float()
    n, h, w, c = t.shape
    out = torch.empty(n, dtype=torch.float32, device=t.device)
    call("ie_mean_hw_f32", ptr(t), n, h, w, c, 1, ptr(out), stream())
    return out.view(n, 1, 1, 1)


# ------------------------------------------------------------------ reference API
def sRGBforward(x):
    """data_utils.py:24-36 (elementwise)."""
    _lib.require_cuda(x)
    xf = x.contiguous().float()
    one = torch.ones(1, dtype=torch.float32, device=x.device)
    out = torch.empty_like(xf)
    call("ie_invert_preproc_f32", ptr(xf), 1, 0, 1, ptr(one), 1, 1, xf.numel(), 0, ptr(out), stream())
    return out


def invert_preproc(imgs, white_level, _nch=1):
    """data_utils.py:42-45: sRGBforward(imgs / white_level)[:, 8:-8, 8:-8]."""
    base, pitch, coff = _nhw_view(imgs)
    n, h, w = imgs.shape
    wl = _wl_vec(white_level, n)
    if pitch < _nch:
        raise _lib.ImgEnhError(f"invert_preproc: averaging {_nch} channels needs an NHWC view with pitch >= {_nch} (got {pitch})")
    out = torch.empty(n, h - 2 * LBUFF, w - 2 * LBUFF, dtype=torch.float32, device=imgs.device)
    call("ie_invert_preproc_f32", ptr(base), pitch, coff, _nch, ptr(wl), n, h, w, LBUFF, ptr(out), stream())
    return out


def _loss_sums(a, b):
    _lib.require_cuda(a, b)
    a, b = a.contiguous().float(), b.contiguous().float()
    n, h, w = a.shape
    sums = torch.zeros(2, dtype=torch.float64, device=a.device)
    call("ie_img_loss_sums_f32", ptr(a), ptr(b), n, h, w, ptr(sums), stream())
    return sums, n, h, w


def gradient_loss(guess, truth):
    """data_utils.py:40-41."""
    sums, n, h, w = _loss_sums(guess, truth)
    return (sums[1] / (n * (h - 1) * (w - 1) * 2)).float()


def basic_img_loss(img, truth):
    """data_utils.py:46-51: mean squared error + mean |grad img - grad truth|."""
    sums, n, h, w = _loss_sums(img, truth)
    return (sums[0] / (n * h * w) + sums[1] / (n * (h - 1) * (w - 1) * 2)).float()


def deblur_loss(invert_deblur, invert_gt):
    """data_utils.py:81-96."""
    return basic_img_loss(invert_deblur, invert_gt)


def deblur_layer_loss(y_pred, invert_gt, white_noise):
    """data_utils.py:52-73."""
    burst_size = y_pred.shape[-1] - 1
    loss = basic_img_loss(invert_preproc(y_pred[..., 1], white_noise), invert_gt)
    for i in range(burst_size - 1):
        loss = loss + basic_img_loss(invert_preproc(y_pred[..., i + 2], white_noise), invert_gt)
    return loss


def cost_volume_sums(Basis):
    """fp64 [2] on the device: {cost_volume(Basis) (data_utils.py:97-113), the same summed over the images} - the
    second value is additive over equal-size batches and ranks."""
    _lib.require_cuda(Basis)
    bas = Basis.contiguous().float()
    assert bas.dim() == 5 and bas.shape[1] == bas.shape[2], "Basis must be [N,K,K,T,B]"
    n, K, _, T, B = bas.shape
    per_image = torch.empty(n, dtype=torch.float64, device=bas.device)
    out = torch.empty(2, dtype=torch.float64, device=bas.device)
    call("ie_cost_volume_f32", ptr(bas), n, K * K, T * B, B, ptr(per_image), ptr(out), stream())
    return out


def cost_volume(Basis):
    """data_utils.py:97-113: -(mean variance across the bases) + 0.1 * (mean squared excess of the per-basis tap sums
    over 0.75), a 0-dim fp32 tensor like the reference's."""
    return cost_volume_sums(Basis)[0].float()


def invert_deblur_layer(y_pred, white_noise):
    """data_utils.py:74-80: inverted frames concatenated along the last (width) axis."""
    burst_size = y_pred.shape[-1] - 1
    return torch.cat([invert_preproc(y_pred[..., i + 1], white_noise) for i in range(burst_size)], dim=-1)


def psnr_tf_batch(estimate, truth):
    """data_utils.py:118-119."""
    _lib.require_cuda(estimate, truth)
    a, b = estimate.contiguous().float(), truth.contiguous().float()
    n = a.shape[0]
    count = a.numel() // n
    sums = torch.zeros(n, dtype=torch.float64, device=a.device)
    call("ie_sqdiff_sum_f32", ptr(a), ptr(b), n, count, ptr(sums), stream())
    return (-10.0 * torch.log10(sums / count)).mean().float()


def psnr_deblur(invert_deblur, invert_gt):
    """data_utils.py:121-130."""
    return psnr_tf_batch(invert_deblur, invert_gt)


def psnr_each_layer(invert_gt, white_noise, y_pred):
    """data_utils.py:131-144."""
    burst_size = y_pred.shape[-1] - 1
    psnr = {}
    for i in range(burst_size):
        psnr['da{}_noshow'.format(i)] = psnr_tf_batch(invert_preproc(y_pred[..., i + 1], white_noise), invert_gt)
    return psnr


def psnr_burst0(invert_gt, white_noise, x_batch_burst):
    """data_utils.py:145-154."""
    return psnr_tf_batch(invert_preproc(x_batch_burst[..., 0], white_noise), invert_gt)


def psnr_average_f(invert_gt, white_noise, x_batch_burst):
    """data_utils.py:155-164: the burst mean is taken inside the invert kernel."""
    xb = x_batch_burst.contiguous().float()      # dense NHWC fp32: the kernel reads the T channels of a pixel at pitch T
    T = xb.shape[-1]
    return psnr_tf_batch(invert_preproc(xb[..., 0], white_noise, _nch=T), invert_gt)


def ssim_map_sums(a, b):
    """EXTENSION: per-image SUM of the SSIM map of [N,H,W] pairs ((H-10)*(W-10) values each), fp64 on the device."""
    _lib.require_cuda(a, b)
    a, b = a.contiguous().float(), b.contiguous().float()
    n, h, w = a.shape
    sums = torch.zeros(n, dtype=torch.float64, device=a.device)
    call("ie_ssim_f32", ptr(a), ptr(b), n, h, w, ptr(sums), stream())
    return sums


def ssim(a, b):
    """EXTENSION (not in the reference): per-image SSIM of [N,H,W] pairs, tf.image.ssim semantics."""
    n, h, w = a.shape
    return (ssim_map_sums(a, b) / ((h - 10) * (w - 10))).float()


def ssim_deblur_sums(reconstructed, x_batch_truth, white_noise=None):
    """EXTENSION: SSIM-map sums of invert_preproc(deblurred) vs invert_preproc(ground truth) - the pair eval.py:146-149
    feeds to psnr_deblur - per image (3 launches: two invert_preproc, one SSIM)."""
    wl = white_level_of(x_batch_truth) if white_noise is None else white_noise
    return ssim_map_sums(invert_preproc(reconstructed[..., 0], wl), invert_preproc(x_batch_truth[..., 0], wl))


# ------------------------------------------------------------------ fused path
def eval_metric_sums(reconstructed, x_batch_burst, x_batch_truth, burst_length, white_noise=None, want_crops=False):
    """ONE pass over (recon, burst, truth): per-image fp64 sums [N, (T+3)+(T+1)] (see imgenh_b200.h).

    ``want_crops``: also return ``(invert_preproc(deblurred), invert_preproc(gt))`` [N,h-16,w-16] - the two images
    eval.py:146-149 forms and the SSIM extension consumes - as a by-product of the same pass."""
    _lib.require_cuda(reconstructed, x_batch_burst, x_batch_truth)
    T = burst_length
    rec = reconstructed.contiguous().float()
    xb = x_batch_burst.contiguous().float()
    tr = x_batch_truth.contiguous().float()
    n, h, w, _ = rec.shape
    wl = _wl_vec(white_level_of(tr) if white_noise is None else white_noise, n)
    sums = torch.zeros(n, 2 * T + 4, dtype=torch.float64, device=rec.device)
    if want_crops:
        db = torch.empty(n, h - 2 * LBUFF, w - 2 * LBUFF, dtype=torch.float32, device=rec.device)
        gt = torch.empty_like(db)
        call("ie_eval_metrics_crops_f32", ptr(rec), ptr(xb), xb.shape[-1], ptr(tr), ptr(wl), n, h, w, T, LBUFF, ptr(sums),
             ptr(db), ptr(gt), stream())
        return sums, db, gt
    call("ie_eval_metrics_f32", ptr(rec), ptr(xb), xb.shape[-1], ptr(tr), ptr(wl), n, h, w, T, LBUFF, ptr(sums),
         stream())
    return sums


def eval_metric_sums_with_ssim(reconstructed, x_batch_burst, x_batch_truth, burst_length, white_noise=None):
    """``eval_metric_sums`` plus the per-image SSIM-map sums of the deblurred image (EXTENSION): the sRGB'd crops come out
    of the fused metrics pass itself (no separate invert_preproc launches); two launches in total."""
    sums, db, gt = eval_metric_sums(reconstructed, x_batch_burst, x_batch_truth, burst_length, white_noise=white_noise,
                                    want_crops=True)
    return sums, ssim_map_sums(db, gt)


def reduce_metric_sums(sums, h, w, T, ssim_sums=None):
    """Per-image sums -> additive totals (fp64, on device, no sync):

    [ sum_n psnr_deblur, sum_n psnr_frame_0..T-1, sum_n psnr_burst0, sum_n psnr_average,
      sum_n loss_deblur_n, sum_n loss_perlayer_n, n ]
    where loss_n = mse_n + gradl1_n; batch losses of the reference (global means over equal-size
    images) are these sums divided by n.
    """
    _lib.require_cuda(sums)
    assert sums.dtype == torch.float64 and sums.shape[1] == 2 * T + 4 and sums.is_contiguous()
    if ssim_sums is not None:                       # extension: [..., sum_n mean-SSIM_n, n]  (T + 7 values)
        assert ssim_sums.dtype == torch.float64 and ssim_sums.shape == (sums.shape[0],) and ssim_sums.is_contiguous()
        tot = torch.empty(T + 7, dtype=torch.float64, device=sums.device)
        call("ie_metric_totals_ssim_f64", ptr(sums), ptr(ssim_sums), sums.shape[0], h, w, T, LBUFF, ptr(tot), stream())
        return tot
    tot = torch.empty(T + 6, dtype=torch.float64, device=sums.device)
    call("ie_metric_totals_f64", ptr(sums), sums.shape[0], h, w, T, LBUFF, ptr(tot), stream())
    return tot


def totals_to_report(tot, T):
    """Host-side view of the additive totals (a list / 1-D tensor on CPU)."""
    tot = [float(v) for v in tot]
    n = tot[-1]
    assert len(tot) in (T + 6, T + 7)
    extra = {"ssim": tot[T + 5] / n} if len(tot) == T + 7 else {}
    return {
        **extra,
        "psnr": tot[0] / n,
        "psnr_perlayer": [tot[1 + t] / n for t in range(T)],
        "psnr_noise0": tot[T + 1] / n,
        "psnr_average": tot[T + 2] / n,
        "loss1": tot[T + 3] / n,
        "perlayer_loss": tot[T + 4] / n,
        "count": n,
    }


def eval_metrics(reconstructed, x_batch_burst, x_batch_truth, burst_length):
    """Everything eval.py:144-182 computes for one batch, as a dict of python floats (one D2H copy)."""
    n, h, w, _ = reconstructed.shape
    sums = eval_metric_sums(reconstructed, x_batch_burst, x_batch_truth, burst_length)
    tot = reduce_metric_sums(sums, h, w, burst_length).cpu()
    return totals_to_report(tot, burst_length)


# ------------------------------------------------------------------ preprocessing
def preprocess_image(src_u8, org, params, white_level, sig_read, sig_shot, n_read=None, n_shot=None, seed=None):
    """Batched arithmetic of DataLoader.preprocess_image (data_utils.py:198-265).

    src_u8 [N,Hs,Ws,C] uint8 CUDA; org [N,T,2] int32 crop origins (y,x) per frame in source pixels
    (see oracle.preprocess.frame_origins for how the reference's nested crops map to them);
    white_level / sig_read / sig_shot [N] fp32; n_read / n_shot [N,h,w,T] standard normals or None;
    ``seed`` (int, instead of the noise tensors): draw the normals on the device (Philox4x32-10 + Box-Muller).
    Returns (x [N,h,w,T+add], truth [N,h,w,2]) like the reference's (noisy++sig, truth++white_level).
    """
    _lib.require_cuda(src_u8, org, white_level, sig_read, sig_shot)
    assert src_u8.dtype == torch.uint8 and org.dtype == torch.int32
    n, hs, ws, c = src_u8.shape
    T = params["BURST_LENGTH"]
    h, w, up = params["height"], params["width"], params["upscale"]
    lt = LAYER_TYPES[params["layer_type"]]
    add = {0: 0, 1: 1, 2: 2}[lt]
    dev = src_u8.device
    x = torch.empty(n, h, w, T + add, dtype=torch.float32, device=dev)
    truth = torch.empty(n, h, w, 2, dtype=torch.float32, device=dev)
    f = lambda t: t.contiguous().float()
    nr = f(n_read) if n_read is not None else None
    ns = f(n_shot) if n_shot is not None else None
    if seed is not None:
        assert nr is None and ns is None, "give either the noise tensors or a seed"
        call("ie_preprocess_u8_rng", ptr(src_u8.contiguous()), n, hs, ws, c, ptr(org.contiguous()), up,
             float(params["degamma"]), ptr(f(white_level)), ptr(f(sig_read)), ptr(f(sig_shot)), int(seed) & (2 ** 64 - 1),
             lt, h, w, T, ptr(x), ptr(truth), stream())
        return x, truth
    call("ie_preprocess_u8", ptr(src_u8.contiguous()), n, hs, ws, c, ptr(org.contiguous()), up,
         float(params["degamma"]), ptr(f(white_level)), ptr(f(sig_read)), ptr(f(sig_shot)), ptr(nr), ptr(ns), lt,
         h, w, T, ptr(x), ptr(truth), stream())
    return x, truth


# ------------------------------------------------------------------ validation batches from decoded uint8 images
def draw_burst_params(n, src_hw, params, generator=None):
    """The per-image random draws of DataLoader.preprocess_image for ``n`` source images of size ``src_hw``, on the host
    (data_utils.py:222-236 and the crops of make_first_truth :432-439 / make_truth_hqjitter :442-457):

    * ``crop0``  [n,2]: origin of the (h*up + 2*jitter*up)-sized random crop of the source, in SOURCE pixels - negative
      when the source is smaller than the crop and the reference zero-pads it symmetrically (:435-438);
    * ``use_big`` [n,T-1] / ``frame_off`` [n,T-1,2]: per later frame, whether it is cropped from the big-jitter patch
      (probability min(Poisson(1.5)/T, 1), :451-454) and its random_crop offset inside that patch;
    * ``org`` [n,T,2] int32: the resulting absolute crop origin of every frame (what the kernel consumes);
    * ``white_level`` = 10^U(-1,0), ``sig_read`` = 10^U(-3,-1.5), ``sig_shot`` = 10^U(-2,-1)  [n] fp32.

    The distributions are the reference's; the random streams are torch's (``generator``), not TensorFlow's."""
    g = generator
    T, up = params["BURST_LENGTH"], params["upscale"]
    h, w, jitter, sj = params["height"], params["width"], params["jitter"], params["smalljitter"]
    hs, ws = src_hw
    j_up, delta_up = jitter * up, (jitter - sj) * up
    h_up, w_up = h * up + 2 * j_up, w * up + 2 * j_up
    v_err, h_err = max((h_up - hs + 1) // 2, 0), max((w_up - ws + 1) // 2, 0)            # :435-436
    ri = lambda hi, shape: torch.randint(0, hi + 1, shape, generator=g)                  # inclusive upper bound
    crop0 = torch.stack([ri(hs + 2 * v_err - h_up, (n,)) - v_err, ri(ws + 2 * h_err - w_up, (n,)) - h_err], dim=-1)
    prob = torch.clamp(torch.poisson(torch.full((n,), 1.5), generator=g) / T, max=1.0)   # :451
    use_big = torch.rand(n, T - 1, generator=g) < prob[:, None]                           # :453-454
    off_big = torch.stack([ri(2 * j_up, (n, T - 1)), ri(2 * j_up, (n, T - 1))], dim=-1)
    off_small = torch.stack([ri(2 * sj * up, (n, T - 1)), ri(2 * sj * up, (n, T - 1))], dim=-1)
    frame_off = torch.where(use_big[..., None], off_big, off_small)
    base = torch.where(use_big, 0, delta_up)[..., None]
    org = torch.cat([(crop0 + j_up)[:, None, :], crop0[:, None, :] + base + frame_off], dim=1).to(torch.int32)
    u = lambda lo, hi: 10.0 ** (torch.rand(n, generator=g) * (hi - lo) + lo)
    return {"crop0": crop0, "use_big": use_big, "frame_off": frame_off, "org": org,
            "white_level": u(-1.0, 0.0).float(), "sig_read": u(-3.0, -1.5).float(), "sig_shot": u(-2.0, -1.0).float()}


def val_batches_from_u8(images_u8, params, batch_size=None, seed=1234, device=None, shuffle=True):
    """``DataLoader.get_val_ds`` (data_utils.py:387-394) for images that are already decoded: ``images_u8`` is a uint8
    tensor [N,Hs,Ws,C] (host - pinned or not - or device).  Shuffles the images, turns every ``batch_size`` of them
    into a synthetic burst on the device (``preprocess_image``: jittered crops, 4x AREA down-sample, white level,
    read/shot noise drawn on the device) and yields ``(x [B,h,w,T+add], truth [B,h,w,2])`` CUDA tensors -
    ``drop_remainder=True`` like the reference - ready for ``eval.evaluate(..., pre_sharded=...)``.
    Host images are staged like tf.data's ``prefetch`` (:393): the source frames of batch i+1 cross PCIe on a side
    stream (through a pinned gather buffer when the batch is not a contiguous slice of pinned memory) while batch i
    is preprocessed and consumed; two device slots rotate.
    File listing, decoding and the tf.data cache are out of scope."""
    if images_u8.dtype != torch.uint8 or images_u8.dim() != 4:
        raise _lib.ImgEnhError("images_u8 must be a uint8 tensor [N,Hs,Ws,C]")
    if device is None:
        device = images_u8.device if images_u8.is_cuda else torch.device("cuda", torch.cuda.current_device())
    device = torch.device(device)
    bs = int(params.get("batch_size", 1) if batch_size is None else batch_size)
    n = images_u8.shape[0]
    g = torch.Generator().manual_seed(int(seed))
    order = torch.randperm(n, generator=g) if shuffle else torch.arange(n)
    starts = list(range(0, n - bs + 1, bs))                                # drop_remainder=True (:392)
    on_host = not images_u8.is_cuda
    if on_host:
        main = torch.cuda.current_stream(device)
        copy = torch.cuda.Stream(device)
        shape = (bs,) + tuple(images_u8.shape[1:])
        dev_slots = [torch.empty(shape, dtype=torch.uint8, device=device) for _ in range(2)]
        pin_slots = [None, None]
        ready = [None, None]
        released = [None, None]
        copied = [None, None]                                              # the pinned gather buffer may be refilled

    def stage(i):
        """Start the host->device copy of batch i into slot i % 2 on the side stream."""
        k = i % 2
        idx = order[starts[i]:starts[i] + bs]
        contiguous = bool((idx[1:] - idx[:-1] == 1).all()) if bs > 1 else True
        if contiguous and images_u8.is_pinned():
            src = images_u8[int(idx[0]):int(idx[0]) + bs]
        else:
            if pin_slots[k] is None:
                pin_slots[k] = torch.empty(shape, dtype=torch.uint8).pin_memory()
            if copied[k] is not None:
                copied[k].synchronize()                                    # the previous DMA out of this buffer is done
            torch.index_select(images_u8, 0, idx, out=pin_slots[k])
            src = pin_slots[k]
        if released[k] is not None:
            copy.wait_event(released[k])                                   # the consumer is done with the device slot
        with torch.cuda.stream(copy):
            dev_slots[k].copy_(src, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(copy)
        ready[k] = copied[k] = ev

    if on_host and starts:
        stage(0)
    for i, b0 in enumerate(starts):
        d = draw_burst_params(bs, images_u8.shape[1:3], params, generator=g)
        noise_seed = int(torch.randint(0, 2 ** 62, (1,), generator=g))
        if on_host:
            k = i % 2
            if i + 1 < len(starts):
                stage(i + 1)                                               # overlaps this batch's kernels
            main.wait_event(ready[k])
            src = dev_slots[k]
        else:
            idx = order[b0:b0 + bs]
            src = images_u8[idx.to(images_u8.device)].to(device, non_blocking=True)
        out = preprocess_image(src, d["org"].to(device), params, d["white_level"].to(device), d["sig_read"].to(device),
                               d["sig_shot"].to(device), seed=noise_seed)
        if on_host:
            released[k] = torch.cuda.Event()
            released[k].record(main)
        yield out
