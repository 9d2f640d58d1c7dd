"""carskit_b200 -- B200-native (sm_100a CUDA) engine for the SGD hot path of irecsys/CARSKit.

The product is `libcarskit_b200.so` (C ABI in include/carskit_b200.h, kernels in csrc/).  The Python
modules here only bind that ABI (capi), mirror the reference's recommender classes on top of it
(recommender) and generate synthetic inputs (synth).  There is no CPU implementation of the update.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
