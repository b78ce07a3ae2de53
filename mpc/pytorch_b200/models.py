"""Learned and affine dynamics modules with closed-form input Jacobians:
host-side mirrors of the reference's ``mpc.dynamics.NNDynamics`` (mpc/dynamics.py:15-131) and
``mpc.dynamics.AffineDynamics`` (mpc/dynamics.py:159-205), the systems ``GradMethods.ANALYTIC`` linearises through
``grad_input`` (reference mpc/mpc.py:495-524).  ``CtrlPassthroughDynamics`` lives in :mod:`.solver`.

These are user models (plain torch Modules, any device); the iLQR step they feed is the CUDA path.  Unlike the
reference, ``grad_input`` does not depend on activations cached by the previous ``forward`` call and never
materialises one weight copy per batch element: the chain rule runs right to left as
``J <- (J * act'(z_i)) @ W_i`` on a [batch, n_state, width] operand.
"""
import torch
from torch import nn

_ACT = {"sigmoid": torch.sigmoid, "relu": torch.relu, "elu": nn.functional.elu}


def _act_slope(name, z):
    """d act / d pre-activation, from the POST-activation value z (what the reference caches, :66-69)."""
    if name == "sigmoid":
        return z * (1.0 - z)                       # :108-110
    if name == "relu":
        return (z > 0).to(z.dtype)                 # :104-107 zeroes the rows with z <= 0
    return torch.where(z > 0, torch.ones_like(z), z + 1.0)     # elu (alpha = 1); the reference asserts here


class NNDynamics(nn.Module):
    """x' = [x +] MLP([x; u]) with ``hidden_sizes`` fully connected layers (reference mpc/dynamics.py:15-37)."""

    def __init__(self, n_state, n_ctrl, hidden_sizes=(100,), activation="sigmoid", passthrough=True):
        super().__init__()
        if activation not in _ACT:
            raise AssertionError(f"activation must be one of {sorted(_ACT)}")
        self.n_state, self.n_ctrl = n_state, n_ctrl
        self.activation, self.passthrough = activation, passthrough
        widths = [n_state + n_ctrl] + list(hidden_sizes) + [n_state]
        self.fcs = nn.ModuleList(nn.Linear(a, b) for a, b in zip(widths[:-1], widths[1:]))

    def _hidden(self, z):
        """Post-activation outputs of the hidden layers and the (linear) output of the last layer."""
        act, zs = _ACT[self.activation], []
        for fc in self.fcs[:-1]:
            z = act(fc(z))
            zs.append(z)
        return zs, self.fcs[-1](z)

    def forward(self, x, u):                                        # :57-79
        squeeze = x.dim() == 1
        if squeeze:
            x = x.unsqueeze(0)
        if u.dim() == 1:
            u = u.unsqueeze(0)
        _, out = self._hidden(torch.cat((x, u), 1))
        if self.passthrough:
            out = out + x
        return out.squeeze(0) if squeeze else out

    def grad_input(self, x, u):
        """(R, S) = (d x'/d x [B,n,n], d x'/d u [B,n,m]) at every row of the batch (:81-130).  Differentiable
        w.r.t. the weights and inputs when called under ``torch.enable_grad()``."""
        squeeze = x.dim() == 1
        if squeeze:
            x, u = x.unsqueeze(0), u.unsqueeze(0)
        zs, _ = self._hidden(torch.cat((x, u), 1))
        J = self.fcs[-1].weight.unsqueeze(0)                        # [1, n, h_last]
        for z, fc in zip(reversed(zs), reversed(self.fcs[:-1])):
            J = (J * _act_slope(self.activation, z).unsqueeze(1)) @ fc.weight
        J = J.expand(x.shape[0], -1, -1)
        n = self.n_state
        R, S = J[:, :, :n], J[:, :, n:]
        if self.passthrough:
            R = R + torch.eye(n, dtype=R.dtype, device=R.device)
        if squeeze:
            R, S = R.squeeze(0), S.squeeze(0)
        return R, S


class AffineDynamics(nn.Module):
    """x' = A x + B u + c with one (A, B, c) shared by the batch (reference mpc/dynamics.py:159-205)."""

    def __init__(self, A, B, c=None):
        super().__init__()
        assert A.dim() == 2 and B.dim() == 2 and (c is None or c.dim() == 1)
        self.A, self.B, self.c = A, B, c

    def forward(self, x, u):
        squeeze = x.dim() == 1
        if squeeze:
            x, u = x.unsqueeze(0), u.unsqueeze(0)
        z = x @ self.A.t() + u @ self.B.t()
        if self.c is not None:
            z = z + self.c
        return z.squeeze(0) if squeeze else z

    def grad_input(self, x, u):
        nb = x.shape[0]
        return self.A.unsqueeze(0).expand(nb, -1, -1), self.B.unsqueeze(0).expand(nb, -1, -1)
