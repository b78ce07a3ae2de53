"""Drop-in import path: ``from mpc.env_dx.cartpole import CartpoleDx`` (reference mpc/env_dx/cartpole.py:28)."""
from mpc.pytorch_b200.dynamics import CartpoleDx  # noqa: F401
