"""Developer aid (lives under tests/ because it checks against the oracle): one test_step_gpu case under a kernel
chosen by MPCB200_KERNEL, per-problem status / qp_iters vs the oracle.  python tests/debug_case.py [B]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import lqr_oracle as orc
from tests.helpers import gen_problem, nominal_controls
from mpc.pytorch_b200.step import lqr_step_raw
seed, B, T, n, m = 19, int(sys.argv[1]) if len(sys.argv) > 1 else 44, 9, 4, 2
C, c, F, f, x0 = gen_problem(seed, B, T, n, m, torch.float32, True, True)
u, ul, uu = nominal_controls(seed, B, T, m, torch.float32, "tensor")
x = orc.get_traj(T, u, x0, F, f)
o = orc.lqr_step_forward(n, m, T, x0, C, c, F, f, x, u, u_lower=ul, u_upper=uu, coupled=False)
d = lambda t: t.cuda()
r = lqr_step_raw(n, m, T, d(x0), d(C), d(c), d(F), d(f), d(x), d(u), u_lower=d(ul), u_upper=d(uu), want_gains=True)
torch.cuda.synchronize()
st = r["status"].cpu()
print("launched", r.get("launched"), "status nonzero at", st.nonzero().flatten().tolist())
dq = r["qp_iters"].cpu().long() - o.qp_iters
print("qp_iters mismatches", dq.nonzero().tolist())
for b in st.nonzero().flatten().tolist():
    print("b", b, "gpu qp", r["qp_iters"][:, b].tolist(), "oracle", o.qp_iters[:, b].tolist())
    print(" ks diff", (r["ks"][:, b].cpu() - o.ks[:, b]).abs().max().item(), "free", r["free_mask"][:, b].tolist(), o.free_masks[:, b].tolist())
    print(" lo", ul[:, b].tolist(), "hi", uu[:, b].tolist())
    print(" ks gpu", r["ks"][:, b].tolist())
    print(" ks orc", o.ks[:, b].tolist())
