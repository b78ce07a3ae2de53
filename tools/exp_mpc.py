"""Developer experiment: MPC.forward (outer iLQR loop) with QuadCost/LinDx at config 4 and 3 sizes."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from mpc import mpc
dev = torch.device("cuda:0")
for (B, T, n, m, iters, bound) in [(1024, 20, 8, 2, 5, 0.25), (4096, 20, 8, 2, 5, 0.25), (4096, 20, 8, 2, 3, None)]:
    s = bench.gen_inputs(11, B, T, n, m, dev)
    kw = {} if bound is None else dict(u_lower=-bound, u_upper=bound)
    ctrl = mpc.MPC(n, m, T, lqr_iter=iters, verbose=-1, exit_unconverged=False, **kw)
    for _ in range(2):
        ctrl(s["x_init"], mpc.QuadCost(s["C"], s["c"]), mpc.LinDx(s["F"], s["f"]))
    torch.cuda.synchronize(); t0 = time.perf_counter()
    reps = 10
    for _ in range(reps):
        x, u, costs = ctrl(s["x_init"], mpc.QuadCost(s["C"], s["c"]), mpc.LinDx(s["F"], s["f"]))
    torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / reps
    print(f"MPC.forward B={B} lqr_iter={iters} bounds={bound}: {dt*1e3:.2f} ms -> {B/dt:.3e} MPC-solves/s")
