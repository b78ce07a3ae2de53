"""Shared test helpers: seeded problem generator and fixture loading."""
import os

import numpy as np
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gen_problem(seed, B, T, n, m, dtype, time_varying=False, with_f=True):
    """Well-conditioned LQR instance: C = LL'+I, LTI (or TV) A = 0.9I + 0.1 N/sqrt(n), B = N/sqrt(n)."""
    g = torch.Generator().manual_seed(seed)
    p = n + m
    f64 = torch.float64
    L = torch.randn(T, B, p, p, generator=g, dtype=f64) / p ** 0.5
    C = L @ L.transpose(-1, -2) + torch.eye(p, dtype=f64)
    c = torch.randn(T, B, p, generator=g, dtype=f64)
    if time_varying:
        A = 0.9 * torch.eye(n, dtype=f64) + 0.1 * torch.randn(T - 1, B, n, n, generator=g, dtype=f64) / n ** 0.5
        Bm = torch.randn(T - 1, B, n, m, generator=g, dtype=f64) / n ** 0.5
        F = torch.cat((A, Bm), -1)
    else:
        A = 0.9 * torch.eye(n, dtype=f64) + 0.1 * torch.randn(B, n, n, generator=g, dtype=f64) / n ** 0.5
        Bm = torch.randn(B, n, m, generator=g, dtype=f64) / n ** 0.5
        F = torch.cat((A, Bm), -1).unsqueeze(0).repeat(T - 1, 1, 1, 1)
    f = 0.1 * torch.randn(T - 1, B, n, generator=g, dtype=f64)
    x0 = torch.randn(B, n, generator=g, dtype=f64)
    out = [t.to(dtype).contiguous() for t in (C, c, F, f, x0)]
    if not with_f:
        out[3] = None
    return out


def nominal_controls(seed, B, T, m, dtype, bounds=None):
    """Returns (u, u_lower, u_upper); bounds: None | float | 'tensor'."""
    g = torch.Generator().manual_seed(seed + 7)
    u = (0.1 * torch.randn(T, B, m, generator=g, dtype=torch.float64)).to(dtype)
    if bounds is None:
        return u, None, None
    if bounds == "tensor":
        ul = (-0.5 * torch.rand(T, B, m, generator=g, dtype=torch.float64) - 0.05).to(dtype)
        uu = (0.5 * torch.rand(T, B, m, generator=g, dtype=torch.float64) + 0.05).to(dtype)
        return torch.maximum(torch.minimum(u, uu), ul), ul, uu
    return u.clamp(-float(bounds), float(bounds)), -float(bounds), float(bounds)


def load_golden(name):
    z = np.load(os.path.join(GOLD, name + ".npz"))
    out = {}
    for k in z.files:
        a = z[k]
        out[k] = torch.from_numpy(a) if a.ndim > 0 else a.item()
    return out


def scalar_or_tensor(v):
    return v if not torch.is_tensor(v) else v


def maxdiff(a, b):
    a = torch.as_tensor(a)
    b = torch.as_tensor(b)
    return float((a.detach().cpu().double() - b.detach().cpu().double()).abs().max()) if a.numel() else 0.0
