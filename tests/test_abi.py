"""The C-ABI shared library loads and exports exactly what include/imgenh_b200.h declares (no GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "imgenh_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ie_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_hot_path():
    names = declared_functions()
    for must in ["ie_conv2d_nhwc_bf16", "ie_kpn_apply_f32", "ie_eval_metrics_f32", "ie_ssim_f32", "ie_preprocess_u8",
                 "ie_maxpool2_nhwc_bf16", "ie_upsample_bilinear_nhwc_bf16", "ie_softmax_taps_f32"]:
        assert must in names


def test_library_exports_every_declared_symbol():
    from imageenhancement_mp_b200 import _lib, build
    build.build()                                   # no-op when up to date; nvcc cross-compiles without a GPU
    assert os.path.exists(_lib.LIB_PATH)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(lib, name), f"{name} is declared in the header but not exported"
    lib.ie_version.restype = ctypes.c_int
    assert lib.ie_version() == 100


def test_python_binding_covers_the_header():
    from imageenhancement_mp_b200 import _lib
    bound = set(_lib.SIGNATURES) | {"ie_last_error"}
    assert set(declared_functions()) <= bound, set(declared_functions()) - bound


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from imageenhancement_mp_b200 import _lib
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    monkeypatch.setattr(_lib, "_lib", None)
    with pytest.raises(_lib.ImgEnhError):
        _lib.load()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "imageenhancement_mp_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert not re.search(r"^\s*(import|from)\s+oracle\b", src, flags=re.M), fn


def test_cpu_tensors_are_rejected_not_computed():
    import torch
    from imageenhancement_mp_b200 import data_utils as du, ImgEnhError
    with pytest.raises(ImgEnhError):
        du.psnr_tf_batch(torch.zeros(1, 4, 4), torch.zeros(1, 4, 4))
