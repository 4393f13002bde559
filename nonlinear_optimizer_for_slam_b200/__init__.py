"""B200-native Gauss-Newton / damped-LM normal-equation assembly for the Mahalanobis/NDT and
reprojection-error pose minimizers of nonlinear_optimizer_for_slam.

Product: libnlo_cuda.so (hand-written sm_100a kernels behind the C ABI of include/nlo_cuda.h) and
the C++ drop-in classes under cxx/.  The Python modules here only marshal numpy arrays.
"""
from . import _capi  # noqa: F401
from .api import (Context, NdtProblem, ReprojProblem, NdtMap, Scan, Options, NloError, LOSS_NONE,  # noqa: F401
                  LOSS_EXPONENTIAL, LOSS_HUBER, LOSS_CAUCHY, identity_pose, pose_from_Rt,
                  pose_to_Rt, host_alloc, host_free, guard_report)
