#!/usr/bin/env python3
"""Developer experiment: time the KKT-adjoint backward path (masked step kernel + grad kernel) at config 3."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from mpc.pytorch_b200.step import lqr_step_raw, lqr_grad_raw
from mpc.pytorch_b200 import LQRStep, QuadCost, LinDx
dev = torch.device("cuda:0")
B, T, n, m = 4096, 20, 8, 2
bounded = len(sys.argv) > 1 and sys.argv[1] == "1"
inp = bench.gen_inputs(3000, B, T, n, m, dev)
kw = dict(u_lower=-0.25, u_upper=0.25) if bounded else {}
o = lqr_step_raw(n, m, T, inp["x_init"], inp["C"], inp["c"], inp["F"], inp["f"], inp["cur_x"], inp["cur_u"], **kw)
nx, nu = o["new_x"], o["new_u"]
wx, wu = torch.randn_like(nx), torch.randn_like(nu)
I = ((nu - (-0.25)).abs() <= 1e-8) | ((nu - 0.25).abs() <= 1e-8) if bounded else None
zx, zu, z0 = torch.zeros_like(nx), torch.zeros_like(nu), torch.zeros_like(inp["x_init"])
r = torch.cat((wx, wu), 2)
def adjoint():
    a = lqr_step_raw(n, m, T, z0, inp["C"], -r, inp["F"], None, zx, zu, u_zero_I=I, want_stats=False)
    return lqr_grad_raw(n, m, T, inp["C"], inp["c"], inp["F"], nx, nu, a["new_x"], a["new_u"], wx, True)
def timeit(fn, reps=50):
    for _ in range(5): fn()
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
t_adj = timeit(adjoint)
a = lqr_step_raw(n, m, T, z0, inp["C"], -r, inp["F"], None, zx, zu, u_zero_I=I, want_stats=False)
t_solve = timeit(lambda: lqr_step_raw(n, m, T, z0, inp["C"], -r, inp["F"], None, zx, zu, u_zero_I=I, want_stats=False))
t_grad = timeit(lambda: lqr_grad_raw(n, m, T, inp["C"], inp["c"], inp["F"], nx, nu, a["new_x"], a["new_u"], wx, True))
# full autograd path through the public API
lv = [inp[k].clone().requires_grad_(True) for k in ("x_init", "C", "c", "F", "f")]
def full():
    fn = LQRStep(n, m, T, true_cost=QuadCost(lv[1], lv[2]), true_dynamics=LinDx(lv[3], lv[4]),
                 current_x=nx, current_u=nu, no_op_forward=True, **kw)
    xo, uo = fn(*lv)
    return torch.autograd.grad((wx * xo).sum() + (wu * uo).sum(), lv)
t_full = timeit(full, 20)
gb = (4 * (T * 100 + T * 8 + (T - 1) * 80 + 2 * T * 10) + 4 * (T * 100 + T * 10 + (T - 1) * 80 + (T - 1) * 8 + 8)) * B / 1e9
print(f"bounded={bounded}: adjoint solve {t_solve:.1f} us, grad assembly {t_grad:.1f} us ({gb / t_grad * 1e6:.0f} GB/s algorithmic), "
      f"both {t_adj:.1f} us -> {B / t_adj:.2f} M adjoint solves/s; autograd API path {t_full:.1f} us")
