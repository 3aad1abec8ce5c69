"""mppi_tf_b200 — B200-native MPPI update step behind the reference's ControllerBase /
ModelBase / CostBase interface.  The product is libmppi_b200.so (mppi_tf_b200/csrc, C-ABI in
include/mppi_b200.h); this package is its ctypes front-end.  No CPU fallback exists."""
from ._capi import MppiError, LIB_PATH, load  # noqa: F401
from .controller import (ControllerBase, ModelBase, CostBase, actionCostPython, blockDiag, comm_unique_id,  # noqa: F401
                         ellipseStateCost, quatStateCost, ellipse3dStateCost,
                         philox_raw)

__all__ = ["ControllerBase", "ModelBase", "CostBase", "actionCostPython", "blockDiag", "comm_unique_id", "ellipseStateCost", "quatStateCost", "ellipse3dStateCost", "philox_raw",
           "MppiError", "LIB_PATH", "load"]
