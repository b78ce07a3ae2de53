"""Known nonlinear systems whose step function and exact Jacobians run inside the CUDA kernels
(SURVEY.md section 8(f) rank 2): drop-in stand-ins for the reference's example environments
``mpc.env_dx.cartpole.CartpoleDx`` (mpc/env_dx/cartpole.py:28-96) and ``mpc.env_dx.pendulum.PendulumDx``
(mpc/env_dx/pendulum.py:17-84, ``simple`` parametrisation).

They are ordinary ``nn.Module`` dynamics - ``forward(x, u)`` is plain torch, so they work anywhere a Module
does - but they also carry ``mpcb200_kind`` / ``mpcb200_params()``.  ``MPC.forward`` recognises that and, on
CUDA tensors, replaces
  * ``util.get_traj``            (T-1 Python calls of the Module per iteration)        -> ``mpcb200_dyn_rollout``
  * ``MPC.linearize_dynamics``   ((T-1)*n_state autograd passes in AUTO_DIFF mode)     -> ``mpcb200_dyn_linearize``
  * the Module rollout of ``lqr_forward`` (reference mpc/lqr_step.py:224-225)          -> inside the step kernel
so that one iLQR iteration is three kernel launches plus the best-iterate bookkeeping.
"""
import ctypes

import torch
from torch.nn import Module

from . import _lib
from ._lib import MpcB200Error, check, ptr, stream_handle

DYN_LINEAR, DYN_CARTPOLE, DYN_PENDULUM = 0, 1, 2


def _host_values(owner, params):
    """Python floats of a (possibly CUDA, possibly learnable) parameter tensor, read back only when it changed:
    the kernels take the parameters by value, and a device->host read per kernel call would serialise the GPU."""
    key = (params.data_ptr(), params._version, params.device)
    hit = getattr(owner, "_mpcb200_host_cache", None)
    if hit is None or hit[0] != key:
        hit = (key, tuple(float(v) for v in params.detach().cpu()))
        owner._mpcb200_host_cache = hit
    return hit[1]


class CartpoleDx(Module):
    """state = (x, dx, cos th, sin th, dth), one control (force, clamped to +-force_mag), semi-implicit Euler."""
    mpcb200_kind = DYN_CARTPOLE
    n_state, n_ctrl = 5, 1

    def __init__(self, params=None):
        super().__init__()
        # gravity, masscart, masspole, length
        self.params = torch.tensor((9.8, 1.0, 0.1, 0.5)) if params is None else params
        assert len(self.params) == 4
        self.force_mag = 100.0
        self.dt = 0.05
        self.lower, self.upper = -self.force_mag, self.force_mag
        self.goal_state = torch.tensor([0.0, 0.0, 1.0, 0.0, 0.0])
        self.goal_weights = torch.tensor([0.1, 0.1, 1.0, 1.0, 0.1])
        self.ctrl_penalty = 0.001
        self.mpc_eps = 1e-4
        self.linesearch_decay = 0.5
        self.max_linesearch_iter = 2

    def mpcb200_params(self):
        g, mc, mp, l = _host_values(self, self.params)
        return (g, mc, mp, l, float(self.force_mag), float(self.dt), 0.0, 0.0)

    def forward(self, state, u):
        single = state.dim() == 1
        if single:
            state, u = state.unsqueeze(0), u.unsqueeze(0)
        g, mc, mp, l = self.params.to(state).unbind()
        total, pml = mp + mc, mp * l
        force = u[:, 0].clamp(-self.force_mag, self.force_mag)
        pos, vel, c, s, om = state.unbind(1)
        th = torch.atan2(s, c)
        cart_in = (force + pml * om ** 2 * s) / total
        th_acc = (g * s - c * cart_in) / (l * (4.0 / 3.0 - mp * c ** 2 / total))
        acc = cart_in - pml * th_acc * c / total
        th2 = th + self.dt * om
        out = torch.stack((pos + self.dt * vel, vel + self.dt * acc, torch.cos(th2), torch.sin(th2),
                           om + self.dt * th_acc), 1)
        return out.squeeze(0) if single else out

    def get_true_obj(self):
        q = torch.cat((self.goal_weights, self.ctrl_penalty * torch.ones(self.n_ctrl)))
        p = torch.cat((-torch.sqrt(self.goal_weights) * self.goal_state, torch.zeros(self.n_ctrl)))
        return q, p


class PendulumDx(Module):
    """state = (cos th, sin th, dth), one control (torque, clamped to +-max_torque); the reference's
    ``simple`` parametrisation (g, m, l)."""
    mpcb200_kind = DYN_PENDULUM
    n_state, n_ctrl = 3, 1

    def __init__(self, params=None, simple=True):
        super().__init__()
        if not simple:
            raise NotImplementedError("only the `simple` (g, m, l) pendulum runs inside the kernels")
        self.simple = True
        self.max_torque = 2.0
        self.dt = 0.05
        self.params = torch.tensor((10.0, 1.0, 1.0)) if params is None else params
        assert len(self.params) == 3
        self.goal_state = torch.tensor([1.0, 0.0, 0.0])
        self.goal_weights = torch.tensor([1.0, 1.0, 0.1])
        self.ctrl_penalty = 0.001
        self.lower, self.upper = -2.0, 2.0
        self.mpc_eps = 1e-3
        self.linesearch_decay = 0.2
        self.max_linesearch_iter = 5

    def mpcb200_params(self):
        g, m, l = _host_values(self, self.params)
        return (g, m, l, 0.0, float(self.max_torque), float(self.dt), 0.0, 0.0)

    def forward(self, x, u):
        single = x.dim() == 1
        if single:
            x, u = x.unsqueeze(0), u.unsqueeze(0)
        g, m, l = self.params.to(x).unbind()
        tq = u.clamp(-self.max_torque, self.max_torque)[:, 0]
        c, s, om = x.unbind(1)
        th = torch.atan2(s, c)
        om2 = om + self.dt * (3.0 * g / (2.0 * l) * s + 3.0 * tq / (m * l ** 2))
        th2 = th + om2 * self.dt
        out = torch.stack((torch.cos(th2), torch.sin(th2), om2), 1)
        return out.squeeze(0) if single else out

    def get_true_obj(self):
        q = torch.cat((self.goal_weights, self.ctrl_penalty * torch.ones(self.n_ctrl)))
        p = torch.cat((-torch.sqrt(self.goal_weights) * self.goal_state, torch.zeros(self.n_ctrl)))
        return q, p


def known_kind(dynamics, n_state, n_ctrl, ref_tensor):
    """(kind, params) if `dynamics` is a known system that can run in the kernels for these shapes / this tensor."""
    kind = getattr(dynamics, "mpcb200_kind", DYN_LINEAR)
    if kind == DYN_LINEAR or not ref_tensor.is_cuda or ref_tensor.dtype not in (torch.float32, torch.float64):
        return DYN_LINEAR, None
    if (n_state, n_ctrl) != (dynamics.n_state, dynamics.n_ctrl):
        return DYN_LINEAR, None
    return kind, tuple(dynamics.mpcb200_params())


def _dyn_array(params):
    return (ctypes.c_double * 8)(*params)


def dyn_rollout_raw(kind, params, T, x_init, u):
    """x = get_traj(T, u, x_init, dynamics) for a known system, ONE kernel (reference mpc/util.py:102-126)."""
    if not x_init.is_cuda:
        raise MpcB200Error("mpc.pytorch_b200 runs on CUDA tensors only (no CPU fallback)")
    dtype, dev = x_init.dtype, x_init.device
    B, n = x_init.shape
    m = u.shape[2]
    if tuple(u.shape) != (T, B, m) or u.device != dev:
        raise MpcB200Error(f"u: expected shape {(T, B, m)} on {dev}, got {tuple(u.shape)} on {u.device}")
    x0 = x_init.detach().to(dtype).contiguous()
    u_ = u.detach().to(dtype).contiguous()
    x = torch.empty(T, B, n, dtype=dtype, device=dev)
    L = _lib.lib()
    fn = L.mpcb200_dyn_rollout_f32 if dtype == torch.float32 else L.mpcb200_dyn_rollout_f64
    with torch.cuda.device(dev):
        rc = fn(kind, _dyn_array(params), B, T, ptr(x0), ptr(u_), ptr(x), stream_handle(dev))
    check(rc, "mpcb200_dyn_rollout")
    return x


def dyn_linearize_raw(kind, params, T, x, u):
    """(F[T-1,B,n,n+m], f[T-1,B,n]) = linearisation of a known system along (x, u), ONE kernel
    (reference MPC.linearize_dynamics, mpc/mpc.py:490-601)."""
    dtype, dev = x.dtype, x.device
    _, B, n = x.shape
    m = u.shape[2]
    x_ = x.detach().to(dtype).contiguous()
    u_ = u.detach().to(dtype).contiguous()
    F = torch.empty(T - 1, B, n, n + m, dtype=dtype, device=dev)
    f = torch.empty(T - 1, B, n, dtype=dtype, device=dev)
    L = _lib.lib()
    fn = L.mpcb200_dyn_linearize_f32 if dtype == torch.float32 else L.mpcb200_dyn_linearize_f64
    with torch.cuda.device(dev):
        rc = fn(kind, _dyn_array(params), B, T, ptr(x_), ptr(u_), ptr(F), ptr(f), stream_handle(dev))
    check(rc, "mpcb200_dyn_linearize")
    return F, f
