"""``mpc.pnqp`` - same import path as the reference module (mpc/pnqp.py)."""
from .pytorch_b200.boxqp import pnqp  # noqa: F401
