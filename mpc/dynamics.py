"""Drop-in import path: ``from mpc.dynamics import NNDynamics, AffineDynamics, CtrlPassthroughDynamics``
(reference mpc/dynamics.py:15, :133, :159); implementations in mpc/pytorch_b200/models.py and solver.py."""
from mpc.pytorch_b200.models import NNDynamics, AffineDynamics  # noqa: F401
from mpc.pytorch_b200.solver import CtrlPassthroughDynamics  # noqa: F401
