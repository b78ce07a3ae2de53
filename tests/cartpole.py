"""Cartpole step function used as the INPUT GENERATOR of BASELINE config 2 (test harness only).

Restates the physics of the reference's CartpoleDx.forward (mpc/env_dx/cartpole.py:63-96) and its
quadratic objective (get_true_obj, :116-124): state = (x, dx, cos th, sin th, dth), force clamp
+-100, semi-implicit Euler with dt = 0.05, parameters (g, m_cart, m_pole, l) = (9.8, 1.0, 0.1, 0.5).
"""
import torch


class Cartpole(torch.nn.Module):
    n_state, n_ctrl = 5, 1
    force_mag, dt = 100.0, 0.05
    g, m_cart, m_pole, length = 9.8, 1.0, 0.1, 0.5

    def forward(self, state, u):
        single = state.dim() == 1
        if single:
            state, u = state.unsqueeze(0), u.unsqueeze(0)
        force = u[:, 0].clamp(-self.force_mag, self.force_mag)
        pos, vel, c, s, om = state.unbind(1)
        total = self.m_pole + self.m_cart
        pml = self.m_pole * self.length
        th = torch.atan2(s, c)
        tmp = (force + pml * om ** 2 * s) / total
        th_acc = (self.g * s - c * tmp) / (self.length * (4.0 / 3.0 - self.m_pole * c ** 2 / total))
        acc = tmp - pml * th_acc * c / total
        pos2 = pos + self.dt * vel
        vel2 = vel + self.dt * acc
        th2 = th + self.dt * om
        om2 = om + self.dt * th_acc
        out = torch.stack((pos2, vel2, torch.cos(th2), torch.sin(th2), om2), 1)
        return out.squeeze(0) if single else out

    @staticmethod
    def objective(dtype=torch.float32):
        goal_w = torch.tensor([0.1, 0.1, 1.0, 1.0, 0.1], dtype=dtype)
        goal = torch.tensor([0.0, 0.0, 1.0, 0.0, 0.0], dtype=dtype)
        q = torch.cat((goal_w, torch.tensor([0.001], dtype=dtype)))
        p = torch.cat((-goal_w.sqrt() * goal, torch.zeros(1, dtype=dtype)))
        return q, p


def initial_states(B, seed=0, dtype=torch.float32):
    """Cartpole notebook recipe (examples/Cartpole Control.ipynb cell 1)."""
    g = torch.Generator().manual_seed(seed)
    import math
    th = (torch.rand(B, generator=g) * 2 - 1) * 2 * math.pi
    thdot = (torch.rand(B, generator=g) - 0.5)
    x = (torch.rand(B, generator=g) - 0.5)
    xdot = (torch.rand(B, generator=g) - 0.5)
    return torch.stack((x, xdot, th.cos(), th.sin(), thdot), 1).to(dtype)
