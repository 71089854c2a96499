"""
zopt_b200 -- B200-native (sm_100a CUDA) drop-in for the LQR / iLQR / DDP / LQR-MPC hot path of
zprihoda/zopt.  Module and function names mirror the reference package:

    zopt_b200.lqrUtils   discreteFiniteHorizonLqr, bilinearAffineLqr          (zopt/lqrUtils.py:144-262)
    zopt_b200.ilqrUtils  trajectoryRollout, forwardPass2, riccatiStep_*, backwardPass_*,
                         ensurePositiveDefinite, iterativeLqr, differentialDynamicProgramming
                                                                               (zopt/ilqrUtils.py)
    zopt_b200.mpcUtils   lqrMpc                                                (zopt/mpcUtils.py:12-81)
    zopt_b200.pytrees    Trajectory, AffinePolicy, ... NamedTuples             (zopt/pytrees.py)
    zopt_b200.quadcopter Quadcopter                                            (zopt/quadcopter.py)

Every array argument may carry ONE extra leading batch axis; outputs are torch CUDA tensors in the
reference's pytree types.  All arithmetic runs in hand-written CUDA kernels reached through the C
ABI in include/zopt_b200.h; importing the package without the built library raises.
"""
from . import _lib  # noqa: F401  (fails loudly when libzopt_b200.so is missing)

__all__ = ["lqrUtils", "ilqrUtils", "mpcUtils", "pytrees", "quadcopter"]
__version__ = "0.1.0"
