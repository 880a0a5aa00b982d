"""Drop-in for the validation loop and report of the reference's ``eval.py``.

``evaluate`` mirrors /root/reference/eval.py:139-195: for every validation batch run the model,
invert the preprocessing, accumulate two losses and four PSNR flavours, then print the seven
report lines in the reference's format.  What changes is how: one fused metrics kernel per
batch instead of 3T+4 ``invert_preproc`` passes, running totals kept on the device in fp64
(the reference does a ``.numpy()`` sync for every number of every batch, :165-182), batches
sharded over the ranks of a torchrun job and combined by ONE all-reduce at the end.

Out of scope here: argparse CLI, checkpoint restore, TensorBoard writer,
the ``.npz`` visualisation dump.
"""
from __future__ import annotations

import torch

from . import data_utils as du
from . import dist as _dist

REPORT_KEYS = ("val_deblur_loss", "val_perlayer_loss", "val_total_loss", "val_psnr", "val_psnrnoshow0",
               "val_psnrburst0", "val_psnraverage")


def gpu_step_totals(model, x_batch_burst, x_batch_truth, burst_length):
    """Forward + fused metrics for one (local) batch -> additive fp64 totals on the device."""
    reconstructed = model(x_batch_burst)[0]                                       # eval.py:143
    n, h, w, _ = reconstructed.shape
    sums = du.eval_metric_sums(reconstructed, x_batch_burst, x_batch_truth, burst_length)   # :144-182
    return du.reduce_metric_sums(sums, h, w, burst_length)


def make_report(totals, num_batches, burst_length):
    """Totals (after the all-reduce) -> the numbers eval.py:186-195 prints.

    ``val_psnrnoshow0`` reproduces the reference's leading-zero bias (its per-layer lists start as
    ``[0]``, eval.py:136,193): mean over num_batches+1 entries.  ``val_psnrnoshow0_unbiased`` is the
    plain mean.
    """
    r = du.totals_to_report(totals, burst_length)
    return {
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
    """The seven lines of eval.py:186-195, verbatim format."""
    return ['epoch %s: %s = %s' % (int(step), k, report[k]) for k in REPORT_KEYS]


def evaluate(model, val_batches, params, step=1, out=print, step_totals=None):
    """Validation loop.  val_batches yields (x_batch_burst [N,H,W,T+add], x_batch_truth [N,H,W,2]).

    Every rank of a torchrun job must iterate the same batches; each takes its contiguous slice.
    ``step_totals(model, xb, xt, T)`` may be injected (CPU tests); default is the GPU path.
    Returns the report dict (identical on all ranks); rank 0 prints it through ``out``.
    """
    T = params["BURST_LENGTH"]
    fn = gpu_step_totals if step_totals is None else step_totals
    totals = None
    nb = 0
    for x_batch_burst, x_batch_truth in val_batches:                               # eval.py:139
        xb, xt = _dist.shard_batch([x_batch_burst, x_batch_truth])
        nb += 1
        if xb.shape[0] == 0:
            continue
        t = fn(model, xb, xt, T)
        totals = t.clone() if totals is None else totals.add_(t)
    if totals is None:
        raise ValueError("evaluate: no validation data on this rank")
    _dist.all_reduce_totals(totals)                                                # the one exchange step
    report = make_report(totals.cpu(), nb, T)                                      # the one D2H copy
    if _dist.rank() == 0 and out is not None:
        out('-----------------------------validation resule for %d------------------------------' % int(report["count"]))
        for line in format_report(report, step):
            out(line)
    return report
