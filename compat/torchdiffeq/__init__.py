"""torchdiffeq stand-in: fixed-step Euler, the only method the reference uses (scene/blce.py:307,
arguments/__init__.py:215)."""
import torch


def odeint(func, y0, t, rtol=None, atol=None, method="euler", **kw):
    ys = [y0]
    for i in range(len(t) - 1):
        ys.append(ys[-1] + (t[i + 1] - t[i]) * func(t[i], ys[-1]))
    return torch.stack(ys, 0)


odeint_adjoint = odeint
