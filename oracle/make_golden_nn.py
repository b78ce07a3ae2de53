#!/usr/bin/env python3
"""Fixtures for the learned / affine dynamics modules, from the REAL reference.

Run in the build container (where /root/reference exists):   python oracle/make_golden_nn.py
Imports the unmodified reference under the alias ``ref_mpc`` (oracle/make_golden.py: load_reference), builds its
``NNDynamics`` / ``AffineDynamics`` (reference mpc/dynamics.py:15-131, :159-205) with seeded weights, and stores the
weights, one batched step, the analytic Jacobians ``grad_input`` and a short box-constrained iLQR solve with
``GradMethods.ANALYTIC`` as tests/golden/nn_dynamics_*.npz / affine_dynamics_f64.npz.  Only numbers are stored.
"""
import contextlib
import io
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)
from make_golden import load_reference, npz          # noqa: E402


def main():
    rmpc, _, _, _ = load_reference()
    import ref_mpc.dynamics as rdyn
    torch.set_default_dtype(torch.float64)
    n, m, B, T = 3, 2, 4, 8
    p = n + m
    for act in ("sigmoid", "relu"):
        torch.manual_seed(7 if act == "sigmoid" else 8)
        net = rdyn.NNDynamics(n, m, hidden_sizes=[12, 10], activation=act, passthrough=True).double()
        with torch.no_grad():
            for fc in net.fcs:                       # O(1) Jacobians so the iLQR problem is non-trivial
                fc.weight.mul_(1.5)
        xs, us = torch.randn(9, n), torch.randn(9, m)
        nxt = net(xs, us)
        R, S = net.grad_input(xs, us)
        L = torch.randn(T, B, p, p) / p ** 0.5
        C = L @ L.transpose(-1, -2) + torch.eye(p)
        c = 0.5 * torch.randn(T, B, p)
        x0 = torch.randn(B, n)
        with contextlib.redirect_stdout(io.StringIO()):
            x, u, costs = rmpc.MPC(n, m, T, u_lower=-0.6, u_upper=0.6, lqr_iter=12, verbose=-1,
                                   grad_method=rmpc.GradMethods.ANALYTIC, exit_unconverged=False,
                                   detach_unconverged=False, eps=1e-6)(x0, rmpc.QuadCost(C, c), net)
        with contextlib.redirect_stdout(io.StringIO()):          # same problem without bounds: no QP tolerance involved
            xf, uf, cf = rmpc.MPC(n, m, T, lqr_iter=12, verbose=-1, grad_method=rmpc.GradMethods.ANALYTIC,
                                  exit_unconverged=False, detach_unconverged=False, eps=1e-6)(
                x0, rmpc.QuadCost(C, c), net)
        ws = {f"W{i}": fc.weight for i, fc in enumerate(net.fcs)}
        ws.update({f"b{i}": fc.bias for i, fc in enumerate(net.fcs)})
        npz(f"nn_dynamics_{act}_f64", step_x=xs, step_u=us, step_next=nxt, R=R, S=S, C=C, c=c, x_init=x0,
            x=x, u=u, costs=costs, x_free=xf, u_free=uf, costs_free=cf, n_layers=np.int64(len(net.fcs)), **ws)
        print(act, "clamped fraction", float((u.abs() == 0.6).double().mean()), "cost", float(costs.mean()))
    # affine dynamics: x' = A x + B u + c (one system shared by the batch)
    torch.manual_seed(9)
    A = 0.9 * torch.eye(n) + 0.2 * torch.randn(n, n)
    Bm = torch.randn(n, m)
    cc = 0.1 * torch.randn(n)
    aff = rdyn.AffineDynamics(A, Bm, cc)
    L = torch.randn(T, B, p, p) / p ** 0.5
    C = L @ L.transpose(-1, -2) + torch.eye(p)
    c = 0.5 * torch.randn(T, B, p)
    x0 = torch.randn(B, n)
    with contextlib.redirect_stdout(io.StringIO()):
        x, u, costs = rmpc.MPC(n, m, T, u_lower=-0.5, u_upper=0.5, lqr_iter=12, verbose=-1,
                               grad_method=rmpc.GradMethods.ANALYTIC, exit_unconverged=False,
                               detach_unconverged=False, eps=1e-6)(x0, rmpc.QuadCost(C, c), aff)
    npz("affine_dynamics_f64", A=A, B=Bm, c0=cc, C=C, c=c, x_init=x0, x=x, u=u, costs=costs)
    print("affine clamped fraction", float((u.abs() == 0.5).double().mean()))


if __name__ == "__main__":
    main()


def grad_cases():
    """d u* / d c and d u* / d (first-layer bias) of the iLQR solution through NNDynamics, with and without a
    slew-rate penalty, from the reference's own autograd (the quantities its tests
    test_lqr_backward_cost_nn_dynamics_module_constrained[_slew] compare with finite differences,
    tests/test_mpc.py:560-744).  Seeds are searched until the solution is strictly partially on the bounds."""
    rmpc, _, _, _ = load_reference()
    import ref_mpc.dynamics as rdyn
    torch.set_default_dtype(torch.float64)
    n, m, T, B = 2, 2, 3, 1
    p = n + m
    for tag, slew in (("nn_grad_f64", None), ("nn_grad_slew_f64", 1.0)):
        for seed in range(50):
            torch.manual_seed(seed)
            net = rdyn.NNDynamics(n, m, hidden_sizes=[10, 10], activation="sigmoid").double()
            Cf = 10.0 * torch.randn(T, B, p, p)
            C = (Cf.transpose(-1, -2) @ Cf).requires_grad_(True)
            c = (10.0 * torch.randn(T, B, p)).requires_grad_(True)
            x0 = torch.randn(B, n)
            with contextlib.redirect_stdout(io.StringIO()):
                x, u, _ = rmpc.MPC(n, m, T, u_lower=-1.0, u_upper=1.0, lqr_iter=40, verbose=-1,
                                   exit_unconverged=False, max_linesearch_iter=1, slew_rate_penalty=slew,
                                   grad_method=rmpc.GradMethods.ANALYTIC)(x0, rmpc.QuadCost(C, c), net)
            uf = u.reshape(-1)
            on = uf.abs() == 1.0
            if bool(on.any()) and bool((~on).any()):
                break
        else:
            raise RuntimeError("no seed with a partially active solution")
        rows_c, rows_b = [], []
        for i in range(uf.numel()):
            gc, gb = torch.autograd.grad(uf[i], [c, net.fcs[0].bias], retain_graph=True)
            rows_c.append(gc.reshape(-1))
            rows_b.append(gb.reshape(-1))
        ws = {f"W{i}": fc.weight for i, fc in enumerate(net.fcs)}
        ws.update({f"b{i}": fc.bias for i, fc in enumerate(net.fcs)})
        npz(tag, C=C, c=c, x_init=x0, x=x, u=u, du_dc=torch.stack(rows_c), du_db0=torch.stack(rows_b),
            n_layers=np.int64(len(net.fcs)), seed=np.int64(seed), **ws)
        print(tag, "seed", seed, "on bound", int(on.sum()), "of", uf.numel(),
              "max|du/dc|", float(torch.stack(rows_c).abs().max()), "max|du/db0|", float(torch.stack(rows_b).abs().max()))


if __name__ == "__main__" and os.environ.get("GOLDEN_NN_GRAD", "1") == "1":
    grad_cases()


def load_ref_env(modname):
    """Import ref_mpc.env_dx.<modname> (needs matplotlib at import time: stubbed; imports `mpc.util` by its absolute
    name: aliased to the reference only while importing)."""
    import importlib
    import types
    load_reference()
    for mod in ("matplotlib", "matplotlib.pyplot"):
        sys.modules.setdefault(mod, types.ModuleType(mod))
    sys.modules["matplotlib"].use = lambda *a, **k: None
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib.pyplot"].style = types.SimpleNamespace(use=lambda *a, **k: None)
    saved = {k: sys.modules.pop(k) for k in list(sys.modules) if k == "mpc" or k.startswith("mpc.")}
    sys.modules["mpc"] = sys.modules["ref_mpc"]
    sys.modules["mpc.util"] = sys.modules["ref_mpc.util"]
    try:
        env = importlib.import_module("ref_mpc.env_dx." + modname)
    finally:
        for k in [k for k in sys.modules if k == "mpc" or k.startswith("mpc.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    return env


def pendulum_case():
    """Pendulum swing-up iLQR (reference mpc/env_dx/pendulum.py: `simple` model, bounds +-2, decay 0.2, 5 line-search
    iterations), B=16, T=20, float64, AUTO_DIFF Jacobians, 15 iterations: trajectories of the unmodified reference."""
    rmpc, _, _, _ = load_reference()
    rpen = load_ref_env("pendulum")
    torch.set_default_dtype(torch.float64)
    torch.manual_seed(3)
    B, T = 16, 20
    dx = rpen.PendulumDx(params=torch.tensor((10.0, 1.0, 1.0)))
    th = (torch.rand(B) * 2 - 1) * np.pi
    x0 = torch.stack((torch.cos(th), torch.sin(th), torch.rand(B) * 2 - 1), 1)
    q, p = dx.get_true_obj()
    Q = torch.diag(q.double()).repeat(T, B, 1, 1)
    pp = p.double().repeat(T, B, 1)
    with contextlib.redirect_stdout(io.StringIO()):
        x, u, costs = rmpc.MPC(3, 1, T, u_lower=dx.lower, u_upper=dx.upper, lqr_iter=15, verbose=-1,
                               exit_unconverged=False, detach_unconverged=False, linesearch_decay=dx.linesearch_decay,
                               max_linesearch_iter=dx.max_linesearch_iter, grad_method=rmpc.GradMethods.AUTO_DIFF,
                               eps=dx.mpc_eps)(x0, rmpc.QuadCost(Q, pp), dx)
    npz("pendulum_ilqr_f64", x_init=x0, q=q.double(), p=p.double(), x=x, u=u, costs=costs, lqr_iter=np.int64(15))
    print("pendulum: clamped fraction", float((u.abs() == 2.0).double().mean()), "mean cost", float(costs.mean()))


if __name__ == "__main__" and os.environ.get("GOLDEN_PENDULUM", "1") == "1":
    pendulum_case()
