#!/usr/bin/env python3
"""Developer smoke: CUDA step/grad kernels vs the CPU oracle on a few seeded cases + a rough timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import lqr_oracle as orc
from oracle.make_golden import gen_problem
from mpc.pytorch_b200.step import lqr_step_raw, lqr_grad_raw

dev = torch.device("cuda:0")

def run_case(name, seed, B, T, n, m, dtype, bounds=None, delta_u=None, tv=False, wf=True, mask=False):
    C, c, F, f, x0 = gen_problem(seed, B, T, n, m, dtype, tv, wf)
    g = torch.Generator().manual_seed(seed + 7)
    u = (0.1 * torch.randn(T, B, m, generator=g, dtype=torch.float64)).to(dtype)
    ul = uu = None
    if bounds == "tensor":
        ul = (-0.5 * torch.rand(T, B, m, generator=g, dtype=torch.float64) - 0.05).to(dtype)
        uu = (0.5 * torch.rand(T, B, m, generator=g, dtype=torch.float64) + 0.05).to(dtype)
        u = torch.maximum(torch.minimum(u, uu), ul)
    elif bounds is not None:
        ul, uu = -float(bounds), float(bounds)
        u = u.clamp(ul, uu)
    zI = None
    if mask:
        zI = torch.rand(T, B, m, generator=g) < 0.3
    x = orc.get_traj(T, u, x0, F, f)
    o = orc.lqr_step_forward(n, m, T, x0, C, c, F, f, x, u, u_lower=ul, u_upper=uu, delta_u=delta_u,
                             u_zero_I=zI, coupled=False)
    cu = lambda t: None if t is None else (t if isinstance(t, float) else t.to(dev))
    r = lqr_step_raw(n, m, T, cu(x0), cu(C), cu(c), cu(F), cu(f), cu(x), cu(u), u_lower=cu(ul), u_upper=cu(uu),
                     u_zero_I=cu(zI), delta_u=delta_u, want_gains=True)
    torch.cuda.synchronize()
    d = lambda a, b: float((a.cpu().double() - b.double()).abs().max())
    msg = (f"{name:28s} dx={d(r['new_x'], o.new_x):.2e} du={d(r['new_u'], o.new_u):.2e} "
           f"dcost={d(r['costs'], o.costs):.2e} dfdn={d(r['full_du_norm'], o.full_du_norm):.2e} "
           f"dalpha={d(r['alphas'], o.alphas):.2e} dK={d(r['Ks'], o.Ks):.2e} dk={d(r['ks'], o.ks):.2e}")
    if bounds is not None or mask:
        mm = int((r['free_mask'].cpu().bool() != o.free_masks).sum())
        msg += f" mask_mismatch={mm}"
    if bounds is not None:
        msg += f" qp_iter_mismatch={int((r['qp_iters'].cpu().long() != o.qp_iters).sum())}"
    msg += f" status={r['status'].cpu().unique().tolist()}"
    print(msg, flush=True)

run_case("cfg1 f32 n3m1", 1, 1, 5, 3, 1, torch.float32, tv=True)
run_case("unb f64 n4m2", 2, 4, 6, 4, 2, torch.float64)
run_case("unb f32 n8m2 B37", 3, 37, 20, 8, 2, torch.float32)
run_case("unb f32 n8m2 B48", 3, 48, 20, 8, 2, torch.float32)
run_case("box f64 n4m2", 4, 8, 8, 4, 2, torch.float64, bounds=0.25)
run_case("box f32 n8m2 B96", 5, 96, 20, 8, 2, torch.float32, bounds=0.25)
run_case("boxT f64 n3m4", 6, 6, 6, 3, 4, torch.float64, bounds="tensor", tv=True)
run_case("delta f64 n3m2", 7, 4, 6, 3, 2, torch.float64, bounds=0.5, delta_u=0.1)
run_case("box f64 n5m1 nof", 8, 6, 9, 5, 1, torch.float64, bounds=0.3, wf=False)
run_case("box f32 n16m4", 9, 5, 12, 16, 4, torch.float32, bounds=0.25)
run_case("mask f64 n4m2", 10, 9, 7, 4, 2, torch.float64, mask=True)
run_case("pad f64 n3m3", 11, 5, 6, 3, 3, torch.float64, bounds=0.3)
run_case("unb f64 n6m2 T60", 12, 7, 60, 6, 2, torch.float64)

# ---- grad
def grad_case(name, seed, B, T, n, m, bounds):
    dtype = torch.float64
    C, c, F, f, x0 = gen_problem(seed, B, T, n, m, dtype, True, True)
    ul, uu = (None, None) if bounds is None else (-bounds, bounds)
    xs, us, _, _ = orc.mpc_forward_lin(n, m, T, x0, C, c, F, f, u_lower=ul, u_upper=uu, lqr_iter=12, eps=1e-9, coupled=False)
    g = torch.Generator().manual_seed(seed + 3)
    wx = torch.randn(T, B, n, generator=g, dtype=dtype); wu = torch.randn(T, B, m, generator=g, dtype=dtype)
    ref = orc.lqr_step_backward(n, m, T, x0, C, c, F, f, xs, us, wx, wu, u_lower=ul, u_upper=uu, coupled=False)
    from mpc.pytorch_b200 import LQRStep, QuadCost, LinDx
    lv = [t.to(dev).requires_grad_(True) for t in (x0, C, c, F, f)]
    fn = LQRStep(n, m, T, u_lower=ul, u_upper=uu, true_cost=QuadCost(lv[1], lv[2]), true_dynamics=LinDx(lv[3], lv[4]),
                 current_x=xs.to(dev), current_u=us.to(dev), no_op_forward=True)
    xo, uo = fn(*lv)
    loss = (wx.to(dev) * xo).sum() + (wu.to(dev) * uo).sum()
    gr = torch.autograd.grad(loss, lv)
    print(f"{name:28s} " + " ".join(f"{nm}={float((a.cpu()-b).abs().max()):.2e}" for a, b, nm in zip(gr, ref[:5], ("dx0","dC","dc","dF","df"))), flush=True)

grad_case("grad unb n3m2", 21, 3, 5, 3, 2, None)
grad_case("grad box n4m2", 22, 4, 6, 4, 2, 0.35)
grad_case("grad box n8m2 B50", 23, 50, 10, 8, 2, 0.3)

# ---- rough timing, config 3
B, T, n, m = 4096, 20, 8, 2
sets = []
for s in range(4):
    C, c, F, f, x0 = [t.to(dev) for t in gen_problem(100 + s, B, T, n, m, torch.float32)]
    u = torch.zeros(T, B, m, device=dev)
    from mpc.pytorch_b200.solver import get_traj, LinDx
    x = get_traj(T, u, x0, LinDx(F, f))
    sets.append((x0, C, c, F, f, x, u))
for bnd in (None, 0.25):
    for _ in range(5):
        for st in sets:
            lqr_step_raw(n, m, T, *st, u_lower=None if bnd is None else -bnd, u_upper=bnd, want_stats=False)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    reps = 20
    for _ in range(reps):
        for st in sets:
            lqr_step_raw(n, m, T, *st, u_lower=None if bnd is None else -bnd, u_upper=bnd, want_stats=False)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * len(sets))
    print(f"config3 bounds={bnd}: {ms*1e3:.1f} us/step  {B/ms*1e3:.3e} solves/s  hbm={17128*B/ms/1e6:.1f} GB/s", flush=True)
