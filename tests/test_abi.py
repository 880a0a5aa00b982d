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


def header_prototypes():
    """{name: [ctypes kind per argument]} parsed from the header: 'P' pointer, 'I' int, 'LL' long long,
    'ULL' unsigned long long, 'F' float."""
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    src = re.sub(r"//[^\n]*", "", src)
    protos = {}
    for ret, name, args in re.findall(r"\b(long long|int|const char\s*\*)\s+(ie_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", src):
        kinds = []
        for a in [a.strip() for a in args.split(",")]:
            if a in ("", "void"):
                continue
            if "*" in a:
                kinds.append("P")
            elif re.search(r"\bunsigned long long\b", a):
                kinds.append("ULL")
            elif re.search(r"\blong long\b", a):
                kinds.append("LL")
            elif re.search(r"\bfloat\b", a):
                kinds.append("F")
            else:
                assert re.search(r"\b(int|int32_t)\b", a), (name, a)
                kinds.append("I")
        protos[name] = kinds
    return protos


def test_ctypes_signatures_match_the_header_argument_by_argument():
    """A wrong argtypes list would not fail at load time - it would shift every later argument.  Compare the binding's
    tables with the prototypes of include/imgenh_b200.h, kind by kind."""
    from imageenhancement_mp_b200 import _lib
    kind = {ctypes.c_void_p: "P", ctypes.c_int: "I", ctypes.c_longlong: "LL", ctypes.c_ulonglong: "ULL",
            ctypes.c_float: "F"}
    protos = header_prototypes()
    assert set(declared_functions()) == set(protos)
    for name, kinds in protos.items():
        if name == "ie_last_error":
            continue
        got = [kind.get(a) or ("P" if issubclass(a, ctypes._Pointer) else None) for a in _lib.SIGNATURES[name]]
        assert got == kinds, name


def test_integration_stub_matches_the_header(monkeypatch):
    """The ctypes stub printed in INTEGRATION.md section 2 is executed against the built library (load + argtypes only,
    no compute) and its argument lists are compared with the header."""
    from imageenhancement_mp_b200 import _lib, build
    build.build()
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    code = re.search(r"```python\n(# imgenh_ffi\.py.*?)```", text, flags=re.S).group(1)
    code = code.replace('C.CDLL("libimgenh_b200.so")', "C.CDLL(%r)" % _lib.LIB_PATH)
    ns = {}
    exec(compile(code, "INTEGRATION.md", "exec"), ns)
    calls = []

    class Probe:
        def __init__(self, fn):
            self.fn, self.argtypes = fn, None

        def __call__(self, *args):
            calls.append((self.argtypes, args))
            return 0

    lib = ns["_lib"]
    protos = header_prototypes()
    kind = {ctypes.c_void_p: "P", ctypes.c_int: "I", ctypes.c_longlong: "LL", ctypes.c_float: "F"}

    class FakeLib:
        ie_last_error = lib.ie_last_error
        ie_kpn_apply_f32 = Probe(lib.ie_kpn_apply_f32)
        ie_eval_metrics_f32 = Probe(lib.ie_eval_metrics_f32)

    ns["_lib"] = FakeLib
    ns["kpn_apply"](1, 5, 2, 8, 8, 3, 4, 1, 8, 8, 4, 15, 10)
    ns["eval_metrics"](1, 2, 5, 3, 4, 1, 32, 32, 4, 5)
    for (argtypes, args), name in zip(calls, ["ie_kpn_apply_f32", "ie_eval_metrics_f32"]):
        assert [kind[a] for a in argtypes] == protos[name], name
        assert len(args) == len(protos[name]), name
    fields = [f for f, _ in ns["ie_conv_desc"]._fields_]
    hdr = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    body = re.search(r"typedef struct(?:\s+\w+)?\s*\{(.*?)\}\s*ie_conv_desc\s*;", hdr, flags=re.S).group(1)
    names = []
    for decl in re.findall(r"int32_t\s+([^;]+);", body):
        names += [n.strip() for n in decl.split(",")]
    assert fields == names
    assert ctypes.sizeof(ns["ie_conv_desc"]) == 4 * len(names) == ctypes.sizeof(_lib.ConvDesc)
