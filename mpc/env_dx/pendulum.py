"""Drop-in import path: ``from mpc.env_dx.pendulum import PendulumDx`` (reference mpc/env_dx/pendulum.py:17)."""
from mpc.pytorch_b200.dynamics import PendulumDx  # noqa: F401
