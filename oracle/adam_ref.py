"""ORACLE — test infrastructure only (tests/, smoke(), bench cpu_baseline may import it; the product never does).

Restatement of the optimiser step the reference runs over its Gaussian parameters:
`torch.optim.Adam(l, lr=0.0, eps=1e-15)` (scene/gaussian_model.py:641; betas (0.9, 0.999),
amsgrad=False, weight_decay=0) stepped at train.py:796-800.  Follows torch/optim/adam.py
`_single_tensor_adam` operation by operation in float32 numpy.

PINNED: tests/test_oracle.py::test_adam_oracle_matches_torch checks it against torch.optim.Adam itself
(CPU) over several steps.
"""
import numpy as np


def adam_step(p, g, m, v, step, lr, beta1=0.9, beta2=0.999, eps=1e-15):
    """One in-place Adam step on float32 arrays; `step` is the 1-based step count after the increment."""
    f = np.float32
    m += (g - m) * f(1 - beta1)                                   # exp_avg.lerp_(grad, 1 - beta1)
    v *= f(beta2)                                                 # exp_avg_sq.mul_(beta2)
    v += f(1 - beta2) * g * g                                     #   .addcmul_(grad, grad, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2_sqrt = (1 - beta2 ** step) ** 0.5
    denom = np.sqrt(v) / f(bc2_sqrt) + f(eps)                     # (exp_avg_sq.sqrt() / bc2_sqrt).add_(eps)
    p -= f(lr / bc1) * (m / denom)                                # param.addcdiv_(exp_avg, denom, value=-step_size)
    return p, m, v
