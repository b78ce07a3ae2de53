"""GPU: known nonlinear systems inside the kernels - rollout, exact Jacobians
and the fused iLQR iteration, against plain torch (the Module's own forward + autograd) and against the
reference's stored cartpole trajectories."""
import pytest
import torch

from tests.helpers import load_golden, maxdiff

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _systems():
    from mpc.env_dx.cartpole import CartpoleDx
    from mpc.env_dx.pendulum import PendulumDx
    return [("cartpole", CartpoleDx), ("pendulum", PendulumDx)]


def _states(name, B, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    th = (torch.rand(B, generator=g, dtype=torch.float64) * 2 - 1) * 3.0
    r = lambda s: (torch.rand(B, generator=g, dtype=torch.float64) - 0.5) * s
    if name == "cartpole":
        x = torch.stack((r(1.0), r(1.0), torch.cos(th), torch.sin(th), r(2.0)), 1)
    else:
        x = torch.stack((torch.cos(th), torch.sin(th), r(2.0)), 1)
    return x.to(dtype)


@pytest.mark.parametrize("dtype", [torch.float64, torch.float32])
@pytest.mark.parametrize("name", ["cartpole", "pendulum"])
def test_rollout_and_jacobians_match_torch(name, dtype):
    from mpc.pytorch_b200.dynamics import dyn_rollout_raw, dyn_linearize_raw
    cls = dict(_systems())[name]
    dx = cls()
    B, T = 37, 11
    x0 = _states(name, B, dtype, 3)
    g = torch.Generator().manual_seed(5)
    scale = 120.0 if name == "cartpole" else 3.0          # some controls beyond the clamp
    u = ((torch.rand(T, B, 1, generator=g, dtype=torch.float64) - 0.5) * 2 * scale).to(dtype)
    kind, params = dx.mpcb200_kind, dx.mpcb200_params()
    x = dyn_rollout_raw(kind, params, T, x0.to(DEV), u.to(DEV)).cpu()
    want = [x0]
    for t in range(T - 1):
        want.append(dx(want[t], u[t]))
    want = torch.stack(want)
    tol = 1e-11 if dtype == torch.float64 else 3e-5
    assert maxdiff(x, want) <= tol * max(1.0, float(want.abs().max()))
    F, f = dyn_linearize_raw(kind, params, T, want.to(DEV), u.to(DEV))
    xs = want[:-1].reshape(-1, dx.n_state).clone().requires_grad_(True)
    us = u[:-1].reshape(-1, 1).clone().requires_grad_(True)
    nx = dx(xs, us)
    rows = [torch.autograd.grad(nx[:, j].sum(), [xs, us], retain_graph=True) for j in range(dx.n_state)]
    R = torch.stack([r[0] for r in rows], 1)
    S = torch.stack([r[1] for r in rows], 1)
    Fw = torch.cat((R, S), 2).view(T - 1, B, dx.n_state, dx.n_state + 1)
    fw = (nx - torch.einsum("bij,bj->bi", R, xs) - torch.einsum("bij,bj->bi", S, us)).view(T - 1, B, -1).detach()
    jt = 1e-10 if dtype == torch.float64 else 2e-4
    assert maxdiff(F, Fw) <= jt * max(1.0, float(Fw.abs().max()))
    assert maxdiff(f, fw) <= jt * max(1.0, float(fw.abs().max()), float(Fw.abs().max()) * scale)


@pytest.mark.parametrize("name", ["cartpole", "pendulum"])
def test_fused_iteration_equals_module_path(name):
    """MPC with a known system (three kernels per iLQR iteration) == MPC with the same physics as an opaque
    nn.Module (autograd linearisation + torch rollout), float64."""
    from mpc.pytorch_b200 import MPC, QuadCost, GradMethods, _lib
    cls = dict(_systems())[name]
    known = cls()

    class Opaque(torch.nn.Module):                      # hides mpcb200_kind: forces the generic Module path
        def forward(self, x, u):
            return known(x, u)

    B, T = 6, 10
    n = known.n_state
    x0 = _states(name, B, torch.float64, 9).to(DEV)
    q, p = known.get_true_obj()
    Q = torch.diag(q).double().expand(T, B, n + 1, n + 1).contiguous().to(DEV)
    pp = p.double().expand(T, B, n + 1).contiguous().to(DEV)
    kw = dict(u_lower=known.lower, u_upper=known.upper, lqr_iter=6, verbose=-1, exit_unconverged=False,
              detach_unconverged=False, linesearch_decay=known.linesearch_decay,
              max_linesearch_iter=known.max_linesearch_iter, grad_method=GradMethods.AUTO_DIFF, eps=1e-9)
    l0 = _lib.launch_count()
    xa, ua, ca = MPC(n, 1, T, **kw)(x0, QuadCost(Q, pp), known.to(DEV))
    fused_launches = _lib.launch_count() - l0
    xb, ub, cb = MPC(n, 1, T, **kw)(x0, QuadCost(Q, pp), Opaque().to(DEV))
    assert maxdiff(ua, ub) < 1e-7 * max(1.0, float(ub.abs().max()))
    assert maxdiff(xa, xb) < 1e-7 * max(1.0, float(xb.abs().max()))
    assert maxdiff(ca, cb) < 1e-8 * max(1.0, float(cb.abs().max()))
    assert fused_launches <= 6 * 3 + 8                  # rollout + linearise + step per iteration (+ final no-op pass)


def test_gradients_flow_through_a_known_system():
    """The differentiable no-op pass still linearises with autograd (parameters of the system get gradients)."""
    from mpc.pytorch_b200 import MPC, QuadCost, GradMethods
    from mpc.env_dx.pendulum import PendulumDx
    params = torch.tensor((10.0, 1.0, 1.0), dtype=torch.float64, device=DEV, requires_grad=True)
    dx = PendulumDx(params=params)
    B, T = 3, 6
    x0 = _states("pendulum", B, torch.float64, 2).to(DEV)
    q, p = dx.get_true_obj()
    Q = torch.diag(q).double().expand(T, B, 4, 4).contiguous().to(DEV)
    pp = p.double().expand(T, B, 4).contiguous().to(DEV)
    x, u, _ = MPC(3, 1, T, u_lower=-2.0, u_upper=2.0, lqr_iter=30, verbose=-1, exit_unconverged=False,
                  detach_unconverged=False, grad_method=GradMethods.AUTO_DIFF, eps=1e-8)(x0, QuadCost(Q, pp), dx)
    (x.sum() + u.sum()).backward()
    assert params.grad is not None and bool(torch.isfinite(params.grad).all()) and float(params.grad.abs().sum()) > 0


def test_pendulum_ilqr_matches_reference_fixture():
    """Pendulum swing-up (reference mpc/env_dx/pendulum.py recipe: bounds +-2, decay 0.2, 5 line-search iterations,
    eps 1e-3), B=16, T=20, float64, 15 iterations, AUTO_DIFF in the reference; here the known system runs inside the
    kernels (rollout, exact Jacobians by dual numbers, line-search rollout).  Fixture: oracle/make_golden_nn.py."""
    from mpc.pytorch_b200 import MPC, QuadCost, GradMethods
    from mpc.env_dx.pendulum import PendulumDx
    from tests.helpers import load_golden
    g = load_golden("pendulum_ilqr_f64")
    T, B = g["x"].shape[0], g["x"].shape[1]
    dx = PendulumDx(params=torch.tensor((10.0, 1.0, 1.0), dtype=torch.float64))
    Q = torch.diag(g["q"]).repeat(T, B, 1, 1).to(DEV)
    p = g["p"].repeat(T, B, 1).to(DEV)
    ctrl = MPC(3, 1, T, u_lower=dx.lower, u_upper=dx.upper, lqr_iter=int(g["lqr_iter"]), verbose=-1,
               exit_unconverged=False, detach_unconverged=False, linesearch_decay=dx.linesearch_decay,
               max_linesearch_iter=dx.max_linesearch_iter, grad_method=GradMethods.AUTO_DIFF, eps=dx.mpc_eps)
    x, u, costs = ctrl(g["x_init"].to(DEV), QuadCost(Q, p), dx)
    rel = (costs.cpu() - g["costs"]).abs() / g["costs"].abs().clamp_min(1.0)
    assert float(rel.max()) < 1e-7, float(rel.max())
    # controls agree at pnqp's own accuracy (it stops at |dx| < 1e-4; the reference couples that test over the batch,
    # INTEGRATION.md section 2); costs - second order in that difference - to 1e-7, the set of saturated torques exactly
    assert maxdiff(u, g["u"]) < 2e-4 * max(1.0, float(g["u"].abs().max()))
    assert maxdiff(x, g["x"]) < 2e-4 * max(1.0, float(g["x"].abs().max()))
    assert torch.equal(u.abs().cpu() == 2.0, g["u"].abs() == 2.0)
