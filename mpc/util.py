"""``mpc.util`` - same import path as the reference's helper module (mpc/util.py): the batched
small-matrix helpers and trajectory/cost rollouts.  Inside the CUDA kernels these operations are fused
(see mpc/pytorch_b200/csrc/lqr_step.cuh); the functions here are the host-side equivalents kept for
drop-in compatibility of user code (`from mpc import util`).  Pure torch, any device.
"""
import torch

from .pytorch_b200.solver import get_traj, get_cost, _table_log  # noqa: F401  (get_traj: one CUDA kernel for LinDx)


def bger(x, y):
    """Batched outer product x y' (reference mpc/util.py:40)."""
    return x.unsqueeze(2) * y.unsqueeze(1)


def bmv(X, y):
    """Batched matrix-vector product (reference mpc/util.py:44)."""
    return torch.matmul(X, y.unsqueeze(2)).squeeze(2)


def bquad(x, Q):
    """Batched quadratic form x' Q x (reference mpc/util.py:48)."""
    return (x * bmv(Q, x)).sum(1)


def bdot(x, y):
    """Batched dot product (reference mpc/util.py:52)."""
    return (x * y).sum(1)


def bdiag(d):
    """Batched diag(d) (reference mpc/util.py:30-37)."""
    assert d.ndimension() == 2
    return torch.diag_embed(d)


def eclamp(x, lower, upper):
    """Element-wise clamp that ASSIGNS the bound where violated, in place like the reference
    (mpc/util.py:56-70); bounds are floats or tensors of x's shape."""
    if torch.is_tensor(lower):
        assert x.size() == lower.size()
    if torch.is_tensor(upper):
        assert x.size() == upper.size()
    lo = torch.as_tensor(lower, dtype=x.dtype, device=x.device).expand_as(x)
    hi = torch.as_tensor(upper, dtype=x.dtype, device=x.device).expand_as(x)
    x.copy_(torch.where(x < lo, lo, x))
    x.copy_(torch.where(x > hi, hi, x))
    return x


def get_data_maybe(x):
    return x.data if torch.is_tensor(x) else x


def detach_maybe(x):
    if x is None:
        return None
    return x if not x.requires_grad else x.detach()


def data_maybe(x):
    return None if x is None else x.data


def expandParam(X, n_batch, nDim):
    if X.ndimension() in (0, nDim):
        return X, False
    if X.ndimension() == nDim - 1:
        return X.unsqueeze(0).expand(*([n_batch] + list(X.size()))), True
    raise RuntimeError("Unexpected number of dimensions.")


def jacobian(f, x, eps):
    """Central-difference Jacobian of f at a single point x (reference mpc/util.py:8-18)."""
    if x.ndimension() == 2:
        assert x.size(0) == 1
        x = x.squeeze()
    basis = torch.eye(len(x), dtype=x.dtype, device=x.device)
    cols = [(f(x + eps * basis[i]) - f(x - eps * basis[i])) / (2.0 * eps) for i in range(len(x))]
    return torch.stack(cols).transpose(0, 1)


def table_log(tag, d):
    """Markdown-ish row logger used by MPC(verbose>0) (reference mpc/util.py:77-99)."""
    _table_log(tag, d)
