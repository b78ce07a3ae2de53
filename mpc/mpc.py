"""``mpc.mpc`` - same import path as the reference module (mpc/mpc.py): MPC, QuadCost, LinDx,
GradMethods.  Thin re-export of :mod:`mpc.pytorch_b200.solver`."""
from .pytorch_b200.solver import (MPC, QuadCost, LinDx, GradMethods, SlewRateCost,  # noqa: F401
                                  CtrlPassthroughDynamics)
from .pytorch_b200.step import LQRStep  # noqa: F401
