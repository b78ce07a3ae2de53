"""BASELINE config 2: cartpole iLQR MPC (B=128, T=25, bounds +-100, <=50 iterations, eps 1e-2, AUTO_DIFF):
wall time of MPC.forward with the known-system kernels vs the same physics as an opaque nn.Module."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from mpc import mpc
from mpc.env_dx.cartpole import CartpoleDx
from mpc.pytorch_b200 import _lib
from tests.cartpole import initial_states
dev = torch.device("cuda:0")
B, T = 128, 25
dx = CartpoleDx()


class Opaque(torch.nn.Module):
    def forward(self, x, u):
        return dx(x, u)


x0 = initial_states(B, seed=0).to(dev)
q, p = dx.get_true_obj()
Q = torch.diag(q).expand(T, B, 6, 6).contiguous().to(dev)
pp = p.expand(T, B, 6).contiguous().to(dev)
for name, dyn in (("known system (kernels)", dx), ("opaque nn.Module (autograd + torch rollout)", Opaque())):
    ctrl = mpc.MPC(5, 1, T, u_lower=dx.lower, u_upper=dx.upper, lqr_iter=50, verbose=-1, exit_unconverged=False,
                   detach_unconverged=False, linesearch_decay=dx.linesearch_decay,
                   max_linesearch_iter=dx.max_linesearch_iter, grad_method=mpc.GradMethods.AUTO_DIFF, eps=1e-2)
    with torch.no_grad():
        ctrl(x0, mpc.QuadCost(Q, pp), dyn)
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        x, u, costs = ctrl(x0, mpc.QuadCost(Q, pp), dyn)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"cartpole MPC.forward B={B} T={T} <=50 iters, {name}: {dt*1e3:.1f} ms -> {B/dt:.1f} MPC-solves/s, "
          f"{(_lib.launch_count()-l0)//reps} library kernels per solve, mean cost {float(costs.mean()):.4f}", flush=True)
