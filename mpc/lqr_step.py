"""``mpc.lqr_step`` - same import path as the reference module (mpc/lqr_step.py)."""
from .pytorch_b200.step import LQRStep  # noqa: F401
