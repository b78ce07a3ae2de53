"""Seeded random sweep of the whole MPC.forward loop on LinDx problems (get_traj -> LQRStep -> best-iterate tracking ->
stop test, reference mpc/mpc.py:248-301) against the oracle's loop with per-problem pnqp, float64: shapes, horizons,
batch sizes, bounds, delta_u, u_init, early stop."""
import random

import pytest
import torch

from oracle import lqr_oracle as orc
from tests.helpers import gen_problem, maxdiff

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")


def make_cases(count=14, seed=31):
    rng = random.Random(seed)
    out = []
    for i in range(count):
        n, m = rng.choice([(2, 2), (3, 1), (4, 2), (5, 1), (6, 2), (8, 2), (8, 4), (16, 4), (3, 3)])
        T = rng.choice([2, 4, 7, 10])
        B = rng.choice([1, 4, 9, 20])
        bounds = rng.choice([None, 0.3, "tensor"])
        delta = rng.choice([None, None, 0.2]) if bounds is not None else None
        iters = rng.choice([1, 3, 8])
        out.append((f"m{i}_n{n}m{m}_T{T}_B{B}_{'unb' if bounds is None else 'boxT' if bounds == 'tensor' else 'box'}"
                    f"{'_du' if delta else ''}_it{iters}", 900 + i, B, T, n, m, bounds, delta, iters, rng.random() < 0.4))
    return out


CASES = make_cases()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_random_mpc_loop_matches_oracle(case):
    from mpc.pytorch_b200 import MPC, QuadCost, LinDx
    name, seed, B, T, n, m, bounds, delta, iters, with_init = case
    dt = torch.float64
    C, c, F, f, x0 = gen_problem(seed, B, T, n, m, dt, seed % 2 == 1, True)
    g = torch.Generator().manual_seed(seed)
    if bounds is None:
        ul = uu = None
    elif bounds == "tensor":
        ul = -0.5 * torch.rand(T, B, m, generator=g, dtype=dt) - 0.05
        uu = 0.5 * torch.rand(T, B, m, generator=g, dtype=dt) + 0.05
    else:
        ul, uu = -bounds, bounds
    u_init = 0.05 * torch.randn(T, B, m, generator=g, dtype=dt) if with_init else None
    ox, ou, oc, _ = orc.mpc_forward_lin(n, m, T, x0, C, c, F, f, u_lower=ul, u_upper=uu, u_init=u_init, lqr_iter=iters,
                                        delta_u=delta, coupled=False)
    cu = lambda t: t.to(DEV) if torch.is_tensor(t) else t
    ctrl = MPC(n, m, T, u_lower=cu(ul), u_upper=cu(uu), u_init=cu(u_init), lqr_iter=iters, delta_u=delta, verbose=-1,
               exit_unconverged=False, detach_unconverged=False)
    x, u, costs = ctrl(cu(x0), QuadCost(cu(C), cu(c)), LinDx(cu(F), cu(f)))
    sc = max(1.0, float(ox.abs().max()))
    assert maxdiff(u, ou) <= 1e-8 * sc and maxdiff(x, ox) <= 1e-8 * sc, (name, maxdiff(u, ou), maxdiff(x, ox))
    assert maxdiff(costs, oc) <= 1e-9 * max(1.0, float(oc.abs().max()))
    if ul is not None:
        lo = ul if torch.is_tensor(ul) else torch.full_like(ou, ul)
        hi = uu if torch.is_tensor(uu) else torch.full_like(ou, uu)
        assert torch.equal(u.cpu() == lo, ou == lo) and torch.equal(u.cpu() == hi, ou == hi)
