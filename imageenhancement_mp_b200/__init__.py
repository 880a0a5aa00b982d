"""imageenhancement_mp_b200 - B200-native hot path of hanxuel/ImageEnhancement_MP.

Batched inference through the reference's burst-denoising basis-prediction networks
(``model_library.Simplemodel`` / ``Basis_kpn``) plus the per-pixel quality metrics of
``data_utils`` / ``eval`` - behind the reference's own Python API, computed by hand-written
sm_100a kernels (csrc/) reached through a C ABI (include/imgenh_b200.h).

Importing the package does not need a GPU; calling into it does (there is no CPU fallback).
"""
from . import weights, synth  # noqa: F401
from ._lib import ImgEnhError, LIB_PATH  # noqa: F401

__all__ = ["model_library", "data_utils", "eval", "dist", "ops", "engine", "weights", "synth", "ImgEnhError"]
__version__ = "0.1.0"
