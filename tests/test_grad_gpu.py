"""GPU parity of the KKT-adjoint backward (LQRStepFn.backward, reference mpc/lqr_step.py:312-407):
against the reference's own autograd results (fixtures), the oracle, and float64 finite differences
(the reference's test strategy, tests/test_mpc.py:303-500, atol 1e-4)."""
import glob
import os

import pytest
import torch

from oracle import lqr_oracle as orc
from tests.helpers import GOLD, gen_problem, load_golden, maxdiff

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def _mpc_grads(n, m, T, x0, C, c, F, f, bound, lqr_iter, wx, wu, dtype=torch.float64):
    from mpc import mpc
    lv = [t.to(DEV).to(dtype).requires_grad_(True) for t in (x0, C, c, F, f)]
    ul, uu = (None, None) if bound is None else (-bound, bound)
    ctrl = mpc.MPC(n, m, T, u_lower=ul, u_upper=uu, lqr_iter=lqr_iter, verbose=-1,
                   exit_unconverged=False, detach_unconverged=False, eps=1e-9, back_eps=1e-9)
    xs, us, costs = ctrl(lv[0], mpc.QuadCost(lv[1], lv[2]), mpc.LinDx(lv[3], lv[4]))
    loss = (wx.to(DEV).to(dtype) * xs).sum() + (wu.to(DEV).to(dtype) * us).sum()
    return xs, us, torch.autograd.grad(loss, lv)


@pytest.mark.parametrize("name", sorted(os.path.basename(p)[:-4] for p in
                                        glob.glob(os.path.join(GOLD, "grad_*.npz"))))
def test_mpc_backward_matches_reference_autograd(name):
    g = load_golden(name)
    T, B, p = g["C"].shape[0], g["C"].shape[1], g["C"].shape[2]
    n = g["x_init"].shape[1]
    m = p - n
    xs, us, grads = _mpc_grads(n, m, T, g["x_init"], g["C"], g["c"], g["F"], g["f"], g.get("bound"),
                               int(g["lqr_iter"]), g["wx"], g["wu"])
    # forward: bounded problems iterate a batch-coupled pnqp in the reference -> 1e-6; unbounded exact
    ftol = 1e-9 if g.get("bound") is None else 2e-6
    assert maxdiff(xs, g["x"]) <= ftol and maxdiff(us, g["u"]) <= ftol
    for got, key in zip(grads, ("dx_init", "dC", "dc", "dF", "df")):
        scale = max(1.0, float(g[key].abs().max()))
        assert maxdiff(got, g[key]) <= (1e-8 if g.get("bound") is None else 2e-5) * scale, key


def test_backward_at_the_oracle_solution_is_exact():
    """Feed the oracle's converged (x,u) through no_op_forward; gradients then equal the oracle's adjoint."""
    from mpc.pytorch_b200 import LQRStep, QuadCost, LinDx
    B, T, n, m, bound = 50, 10, 8, 2, 0.3
    C, c, F, f, x0 = gen_problem(23, B, T, n, m, torch.float64, True, True)
    xs, us, _, _ = orc.mpc_forward_lin(n, m, T, x0, C, c, F, f, u_lower=-bound, u_upper=bound,
                                       lqr_iter=12, eps=1e-9, coupled=False)
    gsd = torch.Generator().manual_seed(5)
    wx = torch.randn(T, B, n, generator=gsd, dtype=torch.float64)
    wu = torch.randn(T, B, m, generator=gsd, dtype=torch.float64)
    ref = orc.lqr_step_backward(n, m, T, x0, C, c, F, f, xs, us, wx, wu, u_lower=-bound, u_upper=bound,
                                coupled=False)
    lv = [t.to(DEV).requires_grad_(True) for t in (x0, C, c, F, f)]
    fn = LQRStep(n, m, T, u_lower=-bound, u_upper=bound, true_cost=QuadCost(lv[1], lv[2]),
                 true_dynamics=LinDx(lv[3], lv[4]), current_x=xs.to(DEV), current_u=us.to(DEV),
                 no_op_forward=True)
    xo, uo = fn(*lv)
    grads = torch.autograd.grad((wx.to(DEV) * xo).sum() + (wu.to(DEV) * uo).sum(), lv)
    for got, want, key in zip(grads, ref[:5], ("dx_init", "dC", "dc", "dF", "df")):
        assert maxdiff(got, want) <= 1e-10 * max(1.0, float(want.abs().max())), key
    frac = float(((us.abs() - bound).abs() <= 1e-8).double().mean())
    assert 0.05 < frac < 0.95                  # strictly partially active, like tests/test_mpc.py:453-454


@pytest.mark.parametrize("bound", [None, 0.4])
def test_gradients_against_float64_finite_differences(bound):
    """d u / d {c, F, f, x_init} by central differences through the CUDA solver (atol 1e-4)."""
    from mpc import mpc
    B, T, n, m = 1, 3, 2, 2
    C, c, F, f, x0 = [t.to(DEV) for t in gen_problem(77, B, T, n, m, torch.float64, True, True)]

    def solve(c_, F_, f_, x0_):
        ctrl = mpc.MPC(n, m, T, u_lower=None if bound is None else -bound,
                       u_upper=None if bound is None else bound, lqr_iter=25, verbose=-1,
                       exit_unconverged=False, detach_unconverged=False, eps=1e-11, back_eps=1e-11)
        return ctrl(x0_, mpc.QuadCost(C, c_), mpc.LinDx(F_, f_))[1]

    lv = [t.clone().requires_grad_(True) for t in (c, F, f, x0)]
    u = solve(*lv)
    if bound is not None:
        act = (u.abs() - bound).abs() <= 1e-8
        assert bool(act.any()) and not bool(act.all())
    g = torch.Generator().manual_seed(1)
    w = torch.randn(u.shape, generator=g, dtype=torch.float64).to(DEV)
    grads = torch.autograd.grad((w * u).sum(), lv)
    h = 1e-6
    for k, (leaf, grad) in enumerate(zip((c, F, f, x0), grads)):
        fd = torch.zeros_like(leaf)
        flat = leaf.reshape(-1)
        for i in range(flat.numel()):
            args_p = [c, F, f, x0]
            args_m = [c, F, f, x0]
            e = torch.zeros_like(flat)
            e[i] = h
            args_p[k] = (flat + e).view_as(leaf)
            args_m[k] = (flat - e).view_as(leaf)
            with torch.no_grad():
                fd.view(-1)[i] = (w * (solve(*args_p) - solve(*args_m))).sum() / (2 * h)
        assert maxdiff(grad, fd) < 1e-4, ("c", "F", "f", "x_init")[k]


def test_float32_backward_close_to_float64():
    B, T, n, m, bound = 64, 20, 8, 2, 0.25
    C, c, F, f, x0 = gen_problem(88, B, T, n, m, torch.float64)
    gsd = torch.Generator().manual_seed(9)
    wx = torch.randn(T, B, n, generator=gsd, dtype=torch.float64)
    wu = torch.randn(T, B, m, generator=gsd, dtype=torch.float64)
    _, us64, g64 = _mpc_grads(n, m, T, x0, C, c, F, f, bound, 15, wx, wu, torch.float64)
    _, us32, g32 = _mpc_grads(n, m, T, x0, C, c, F, f, bound, 15, wx, wu, torch.float32)
    same_active = ((us64.abs() - bound).abs() <= 1e-8) == ((us32.double().abs() - bound).abs() <= 1e-6)
    ok = same_active.all(0).all(-1)                 # problems whose active set agrees in both precisions
    assert float(ok.double().mean()) > 0.9
    for a, b in zip(g32, g64):
        a, b = a.double(), b
        if a.dim() >= 2 and a.shape[1] == B:
            a, b = a[:, ok], b[:, ok]
        elif a.shape[0] == B:
            a, b = a[ok], b[ok]
        assert maxdiff(a, b) <= 2e-3 * max(1.0, float(b.abs().max()))
