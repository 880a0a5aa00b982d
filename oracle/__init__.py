"""CPU oracle for the hot path of hanxuel/ImageEnhancement_MP.

TEST INFRASTRUCTURE ONLY.  Nothing in ``imageenhancement_mp_b200`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs use it, and only as the checker or the reported baseline.

PARITY STATUS: pinned to the reference's own code, modulo TensorFlow's primitive-op semantics.  The reference
(``/root/reference``) is pure Python on TensorFlow 2.x / Keras, neither of which is installed in this image (no network),
and it ships no tests, fixtures or golden vectors.  This package is a torch-CPU restatement of the reference's
arithmetic, op for op, each function citing the reference ``file:line`` it follows.  It is pinned by
(a) fixtures produced by EXECUTING the unmodified reference ``model_library.py`` / ``data_utils.py`` over
``oracle/tf_standin.py`` - a stand-in for the ~60 TensorFlow / Keras primitives the reference calls - committed as
``tests/golden/ref_*.npz`` (``tests/golden/make_ref_golden.py`` wrote them; the same script runs on a real TensorFlow and
then writes ``tf_*.npz``, which the tests prefer): every wiring decision is the reference's executing code, what stays
assumed is the meaning of the primitives (listed in DESIGN.md section 6);
(b) the analytic known-answer tests derivable from the reference's own code (tests/test_oracle_kat.py),
(c) two independent formulations of the per-pixel filter agreeing to 1e-6, (d) fp32-vs-fp64 agreement,
(e) cross-checks of the assumed op semantics against independent implementations present in this image (scipy
correlate2d / ndimage / softmax, OpenCV half-pixel INTER_LINEAR and INTER_AREA, the Random123 Philox known answers) and
(f) the oracle's own regression fixtures of tests/golden/.
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
