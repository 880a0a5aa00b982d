"""CPU oracle for the hot path of hanxuel/ImageEnhancement_MP.

TEST INFRASTRUCTURE ONLY.  Nothing in ``imageenhancement_mp_b200`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and only as the checker or the reported baseline.

PARITY UNPINNED: the reference (``/root/reference``) is pure Python on TensorFlow 2.x /
Keras, neither of which is installed in this image (no network), and the reference
ships no tests, fixtures or golden vectors.  This package is therefore a torch-CPU
restatement of the reference's arithmetic, op for op, each function citing the
reference ``file:line`` it follows.  It is pinned only by (a) the analytic
known-answer tests derivable from the reference's own code (tests/test_oracle_kat.py),
(b) two independent formulations of the per-pixel filter agreeing to 1e-6, and
(c) fp32-vs-fp64 agreement, (d) cross-checks of the assumed op semantics against
independent implementations present in this image (scipy correlate2d / ndimage / softmax,
OpenCV half-pixel INTER_LINEAR and INTER_AREA, the Random123 Philox known answers) and
(e) the self-generated regression fixtures of tests/golden/.  TensorFlow semantics that
are assumed rather than observed are listed in DESIGN.md ("Oracle").
"""
from .model import (  # noqa: F401
    simplemodel_forward,
    basis_kpn_forward,
    kpn_apply_literal,
    kpn_apply_algebraic,
)
from .metrics import (  # noqa: F401
    sRGBforward,
    invert_preproc,
    gradient,
    gradient_loss,
    basic_img_loss,
    deblur_loss,
    deblur_layer_loss,
    invert_deblur_layer,
    psnr_tf_batch,
    psnr_deblur,
    psnr_each_layer,
    psnr_burst0,
    psnr_average_f,
    cost_volume,
    eval_step,
    eval_report,
)
from .ssim import ssim  # noqa: F401
from .preprocess import preprocess_image  # noqa: F401
