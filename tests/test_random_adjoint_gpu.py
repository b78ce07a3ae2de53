"""Seeded random sweep of LQRStepFn.backward (the one-call fused KKT adjoint where the column-pair mapping takes the
shape, the masked-step + costate + outer-product kernels elsewhere) against the oracle's adjoint, at a solution
obtained by a few oracle iterations: shapes, horizons, batch tails, bounds (none / scalar / tensor), missing f."""
import random

import pytest
import torch

from oracle import lqr_oracle as orc
from tests.helpers import gen_problem, maxdiff

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda:0")
SHAPES = [(2, 2), (3, 1), (4, 2), (4, 4), (5, 1), (6, 2), (8, 2), (8, 4), (12, 4), (16, 4), (3, 3), (7, 4)]


def make_cases(count=24, seed=77):
    rng = random.Random(seed)
    out = []
    for i in range(count):
        n, m = rng.choice(SHAPES)
        T = rng.choice([2, 3, 5, 9])
        B = rng.choice([1, 3, 8, 17, 33])
        dtype = torch.float32 if i % 4 == 3 else torch.float64
        bounds = rng.choice([None, 0.3, "tensor"])
        wf = rng.random() < 0.7
        out.append((f"a{i}_n{n}m{m}_T{T}_B{B}_{'f32' if dtype == torch.float32 else 'f64'}_"
                    f"{'unb' if bounds is None else 'boxT' if bounds == 'tensor' else 'box'}{'' if wf else '_nof'}",
                    500 + i, B, T, n, m, dtype, bounds, wf))
    return out


CASES = make_cases()


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_random_adjoint_matches_oracle(case):
    from mpc.pytorch_b200 import LQRStep, QuadCost, LinDx
    name, seed, B, T, n, m, dtype, bounds, wf = case
    C, c, F, f, x0 = gen_problem(seed, B, T, n, m, dtype, seed % 2 == 0, wf)
    g = torch.Generator().manual_seed(seed)
    if bounds is None:
        kw = {}
    elif bounds == "tensor":
        kw = dict(u_lower=(-0.5 * torch.rand(T, B, m, generator=g, dtype=torch.float64) - 0.05).to(dtype),
                  u_upper=(0.5 * torch.rand(T, B, m, generator=g, dtype=torch.float64) + 0.05).to(dtype))
    else:
        kw = dict(u_lower=-bounds, u_upper=bounds)
    u = torch.zeros(T, B, m, dtype=dtype)
    x = orc.get_traj(T, u, x0, F, f)
    for _ in range(4):                                   # a (nearly) converged solution with a meaningful active set
        o = orc.lqr_step_forward(n, m, T, x0, C, c, F, f, x, u, coupled=False, **kw)
        x, u = o.new_x, o.new_u
    wx = torch.randn(T, B, n, generator=g, dtype=torch.float64).to(dtype)
    wu = torch.randn(T, B, m, generator=g, dtype=torch.float64).to(dtype)
    ref = orc.lqr_step_backward(n, m, T, x0, C, c, F, f, x, u, wx, wu, coupled=False, **kw)
    cu = lambda t: t.to(DEV) if torch.is_tensor(t) else t
    lv = [cu(t).requires_grad_(True) for t in (x0, C, c, F)] + ([cu(f).requires_grad_(True)] if wf else [])
    fn = LQRStep(n, m, T, true_cost=QuadCost(lv[1], lv[2]), true_dynamics=LinDx(lv[3], lv[4] if wf else None),
                 current_x=cu(x), current_u=cu(u), no_op_forward=True, **{k: cu(v) for k, v in kw.items()})
    xo, uo = fn(*lv) if wf else fn(lv[0], lv[1], lv[2], lv[3])
    grads = torch.autograd.grad((xo, uo), lv, (cu(wx), cu(wu)))
    tol = 1e-8 if dtype == torch.float64 else 3e-4
    for gname, a, b in zip(("dx_init", "dC", "dc", "dF", "df"), grads, ref[:5]):
        sc = max(1.0, float(b.abs().max()))
        assert maxdiff(a, b) <= tol * sc, (name, gname, maxdiff(a, b), sc)
