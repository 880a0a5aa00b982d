"""Drop-in for the validation loop and report of the reference's ``eval.py``.

``evaluate`` mirrors /root/reference/eval.py:139-195: for every validation batch run the model,
invert the preprocessing, accumulate two losses and four PSNR flavours, then print the seven
report lines in the reference's format.  What changes is how: one fused metrics kernel per
batch instead of 3T+4 ``invert_preproc`` passes, running totals kept on the device in fp64
(the reference does a ``.numpy()`` sync for every number of every batch, :165-182), batches
sharded over the ranks of a torchrun job and combined by ONE all-reduce at the end.

Batches may live in (pinned) host memory: they are then staged to the device on a side stream, two
slots deep, so the host->device copy of batch i+1 overlaps the forward of batch i - the reference's
tf.data ``prefetch`` (data_utils.py:393) moved to where it matters on a GPU.

The ``--visualization`` ``.npz`` dump (eval.py:201-207) and the ``ps`` lines (``cost_volume`` of the basis,
eval.py:159-162,189-191) are covered; out of scope: argparse CLI, TensorBoard writer.
"""
from __future__ import annotations

import torch

from . import data_utils as du
from . import dist as _dist

REPORT_KEYS = ("val_deblur_loss", "val_perlayer_loss", "val_total_loss", "val_psnr", "val_psnrnoshow0",
               "val_psnrburst0", "val_psnraverage")


def gpu_step_totals(model, x_batch_burst, x_batch_truth, burst_length, vis=None, ssim=False, ps=False):
    """Forward + fused metrics for one (local) batch -> additive fp64 totals on the device.

    ``ps`` (params["ps"], eval.py:159-162): returns ``(totals, cv)`` with cv = fp64 [1], ``cost_volume(Bas)`` summed
    over the images of the batch.

    ``ssim`` (EXTENSION, not in the reference): also the SSIM of the deblurred vs ground-truth sRGB crops; the totals
    then carry one more value (``ie_metric_totals_ssim_f64``).

    ``vis``: optional dict of lists receiving the reference's visualisation arrays for this batch (eval.py:164-169):
    invert_gt, invert_deblur, invert_perlayer, Basis, originbasis (numpy, one device->host copy each)."""
    res = model(x_batch_burst)                                                    # eval.py:143
    reconstructed = res[0]
    n, h, w, _ = reconstructed.shape
    wl = du.white_level_of(x_batch_truth)                                         # :144-145
    if ssim:        # the SSIM extension takes its two sRGB'd crops from the same pass
        sums, ssim_sums = du.eval_metric_sums_with_ssim(reconstructed, x_batch_burst, x_batch_truth, burst_length,
                                                        white_noise=wl)
    else:
        sums, ssim_sums = du.eval_metric_sums(reconstructed, x_batch_burst, x_batch_truth, burst_length,
                                              white_noise=wl), None                                  # :146-182
    if vis is not None:
        vis["invert_gt"].append(du.invert_preproc(x_batch_truth[..., 0], wl).cpu().numpy())         # :146-147
        vis["invert_deblur"].append(du.invert_preproc(reconstructed[..., 0], wl).cpu().numpy())     # :148-149
        vis["invert_perlayer"].append(du.invert_deblur_layer(reconstructed, wl).cpu().numpy())      # :158
        vis["Basis"].append(res[1].cpu().numpy())
        if len(res) > 2:
            vis["originbasis"].append(res[2].cpu().numpy())
    totals = du.reduce_metric_sums(sums, h, w, burst_length, ssim_sums=ssim_sums)
    if ps:
        return totals, du.cost_volume_sums(res[1])[1:2]                            # :160-162
    return totals


def staged_batches(val_batches, device, depth=2, pre_sharded=False):
    """Yields each (sharded) batch as device tensors; host batches are copied on a side stream into
    ``depth`` rotating device slots, one batch ahead of the consumer.  Device batches pass through."""
    device = torch.device(device)
    main = torch.cuda.current_stream(device)
    copy = torch.cuda.Stream(device)
    slots = [None] * depth          # per slot: list of device buffers
    released = [None] * depth       # event on `main`: the consumer is done with the slot

    def stage(k, tensors):
        if all(t.is_cuda for t in tensors):
            return tensors, None
        if slots[k] is None or any(b.shape != t.shape or b.dtype != t.dtype for b, t in zip(slots[k], tensors)):
            slots[k] = [torch.empty(t.shape, dtype=t.dtype, device=device) for t in tensors]
            released[k] = torch.cuda.Event()
            released[k].record(main)
        copy.wait_event(released[k])
        with torch.cuda.stream(copy):
            for b, t in zip(slots[k], tensors):
                b.copy_(t, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(copy)
        return slots[k], ready

    shard = (lambda b: list(b)) if pre_sharded else (lambda b: _dist.shard_batch(list(b)))
    it = iter(val_batches)
    k = 0
    nxt = next(it, None)
    staged = stage(k, shard(nxt)) if nxt is not None else None
    while staged is not None:
        cur, cur_k = staged, k
        nxt = next(it, None)
        k = (k + 1) % depth
        staged = stage(k, shard(nxt)) if nxt is not None else None   # overlaps the step below
        tensors, ready = cur
        if ready is not None:
            main.wait_event(ready)
        yield tensors
        if ready is not None:
            released[cur_k] = torch.cuda.Event()
            released[cur_k].record(main)


def make_report(totals, num_batches, burst_length, cost_volume_sum=None, beta_coef=100.0):
    """Totals (after the all-reduce) -> the numbers eval.py:186-195 prints.

    ``cost_volume_sum`` (params["ps"]): adds ``variance`` (the mean of cost_volume(Bas), eval.py:162,191) and
    ``variance loss`` = beta_coef * variance (:160,190).

    ``val_psnrnoshow0`` reproduces the reference's leading-zero bias (its per-layer lists start as
    ``[0]``, eval.py:136,193): mean over num_batches+1 entries.  ``val_psnrnoshow0_unbiased`` is the
    plain mean.
    """
    r = du.totals_to_report(totals, burst_length)
    extra = {"val_ssim": r["ssim"]} if "ssim" in r else {}      # extension
    if cost_volume_sum is not None:
        variance = float(cost_volume_sum) / r["count"]
        extra.update({"variance loss": beta_coef * variance, "variance": variance})
    return {
        **extra,
        "val_deblur_loss": r["loss1"],
        "val_perlayer_loss": r["perlayer_loss"],
        "val_total_loss": r["loss1"] + r["perlayer_loss"],
        "val_psnr": r["psnr"],
        "val_psnrnoshow0": r["psnr_perlayer"][0] * num_batches / (num_batches + 1),
        "val_psnrnoshow0_unbiased": r["psnr_perlayer"][0],
        "val_psnr_perlayer": r["psnr_perlayer"],
        "val_psnrburst0": r["psnr_noise0"],
        "val_psnraverage": r["psnr_average"],
        "count": r["count"],
    }


def format_report(report, step=1):
    """The seven lines of eval.py:186-195 (nine with the two ``ps`` lines of :189-191), verbatim format and order."""
    keys = list(REPORT_KEYS)
    if "variance" in report:
        keys[3:3] = ["variance loss", "variance"]
    return ['epoch %s: %s = %s' % (int(step), k, report[k]) for k in keys]


def evaluate(model, val_batches, params, step=1, out=print, step_totals=None, step_results=None, device=None,
             pre_sharded=False, visualization=False, dump_path=None, ssim=False, beta_coef=100.0):
    """Validation loop.  val_batches yields (x_batch_burst [N,H,W,T+add], x_batch_truth [N,H,W,2]), on the
    device or in (pinned) host memory.

    Every rank of a torchrun job must iterate the same batches; each takes its contiguous slice
    (``pre_sharded=True``: every rank is handed its own slice already, e.g. one loader per rank).
    ``step_totals(model, xb, xt, T)`` may be injected (CPU tests); default is the GPU path.
    ``step_results``: optional list; every step's additive totals are appended to it as pinned host
    tensors (asynchronous device->host copies, complete when evaluate returns) - the per-batch numbers
    the reference reads with ``.numpy()`` at eval.py:165-182, without its per-batch stalls.
    ``visualization`` (eval.py:41 ``--visualization``): also collect invert_gt / invert_deblur / invert_perlayer /
    Basis / originbasis of every local batch and write them with ``np.savez`` like eval.py:201-207 (to ``dump_path``,
    default ``data<time>.npz``; with several ranks every rank writes ``<stem>.rank<r>.npz`` for its shard).
    ``params["ps"]`` (eval.py:103,159-162,189-191): also report ``variance`` = mean ``cost_volume(Bas)`` and
    ``variance loss`` = ``beta_coef`` * variance.  The reference's eval.py never defines ``beta_coef`` (its ``ps`` branch
    raises NameError at :160); the default here is run_training.py:178 at iterate 0 (anneal^0 * 10^2).  The step
    function then returns ``(totals, cost_volume_sum[1])``; the sum rides in the same all-reduce as the totals.
    ``ssim`` (EXTENSION): also report ``val_ssim`` (deblurred vs ground truth, tf.image.ssim semantics); it travels
    in the same all-reduced totals vector.
    Returns the report dict (identical on all ranks); rank 0 prints it through ``out``.
    """
    T = params["BURST_LENGTH"]
    ps = bool(params.get("ps", False))
    vis = {k: [] for k in ("invert_gt", "invert_deblur", "invert_perlayer", "Basis", "originbasis")} \
        if (visualization and step_totals is None) else None
    if step_totals is not None:
        fn = step_totals
    else:
        fn = lambda m, xb, xt, T_: gpu_step_totals(m, xb, xt, T_, vis=vis, ssim=ssim, ps=ps)
    totals = cv_total = None
    pinned, pinned_used = None, 0
    nb = 0
    if step_totals is None:
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        batches = staged_batches(val_batches, device, pre_sharded=pre_sharded)
    else:
        batches = (list(b) if pre_sharded else _dist.shard_batch(list(b)) for b in val_batches)
    for xb, xt in batches:                                                         # eval.py:139
        nb += 1
        if xb.shape[0] == 0:
            continue
        t = fn(model, xb, xt, T)
        if ps:
            t, cv = t
            cv_total = cv.clone() if cv_total is None else cv_total.add_(cv)
        if step_results is not None:
            # pinned rows are handed out from blocks of 64 steps: one cudaHostAlloc per step (every result stays alive
            # in step_results, so the caching host allocator cannot recycle) cost more than the step's metric kernels
            if pinned is None or pinned_used == pinned.shape[0]:
                pinned = torch.empty((64,) + tuple(t.shape), dtype=t.dtype, pin_memory=t.is_cuda)
                pinned_used = 0
            host = pinned[pinned_used]
            pinned_used += 1
            host.copy_(t, non_blocking=True)
            step_results.append(host)
        totals = t.clone() if totals is None else totals.add_(t)
    if totals is None:
        # this rank's shard of every batch was empty (batch_size < world size, e.g. the reference's default batch of
        # one under torchrun): it still has to enter the collective - with zeros - or the other ranks block in the
        # all-reduce until the NCCL timeout.  "No data anywhere" is decided AFTER the exchange, on every rank.
        n_tot = T + (7 if (ssim and step_totals is None) else 6)
        zdev = device if step_totals is None else "cpu"
        totals = torch.zeros(n_tot, dtype=torch.float64, device=zdev)
        cv_total = torch.zeros(1, dtype=torch.float64, device=zdev)
    if ps:
        totals = torch.cat([totals, cv_total.to(totals.dtype)])
    _dist.all_reduce_totals(totals)                                                # the one exchange step
    host = totals.cpu()                                                            # the one D2H copy
    if float(host[-2 if ps else -1]) == 0.0:
        raise ValueError("evaluate: no validation data on any rank")
    if ps:
        report = make_report(host[:-1], nb, T, cost_volume_sum=host[-1], beta_coef=beta_coef)
    else:
        report = make_report(host, nb, T)
    if vis is not None:
        import datetime
        import numpy as np
        path = dump_path or ("data" + datetime.datetime.now().strftime("%Y%m%d-%H%M%S") + ".npz")
        if _dist.world() > 1:
            path = (path[:-4] if path.endswith(".npz") else path) + ".rank%d.npz" % _dist.rank()
        np.savez(path, **{k: np.stack(v) for k, v in vis.items() if v})            # eval.py:201-207
        report["dump_path"] = path
    if _dist.rank() == 0 and out is not None:
        out('-----------------------------validation resule for %d------------------------------' % int(report["count"]))
        for line in format_report(report, step):
            out(line)
    return report
