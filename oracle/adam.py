"""Oracle (TEST INFRASTRUCTURE ONLY) for the fused optimiser step ``gn_adam_step``.

numpy fp32 restatement of torch's single-tensor Adam (``torch/optim/adam.py::_single_tensor_adam``,
amsgrad=False, maximize=False, capturable=False) — what the reference scripts run through
``torch.optim.Adam(model.parameters(), lr)`` (``GripNet-pose.py:104,146``).  Pinned in
``tests/test_optim.py`` against ``torch.optim.Adam`` itself (CPU, this image's torch 2.11).
"""
import math

import numpy as np


class AdamOracle:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0):
        self.p = [np.array(p, dtype=np.float32, copy=True) for p in params]
        self.m = [np.zeros_like(p) for p in self.p]
        self.v = [np.zeros_like(p) for p in self.p]
        self.lr, self.b1, self.b2, self.eps, self.wd = lr, betas[0], betas[1], eps, weight_decay
        self.t = 0

    def step(self, grads):
        self.t += 1
        f = np.float32
        bc1 = 1.0 - self.b1 ** self.t                      # python doubles, as in torch
        bc2 = 1.0 - self.b2 ** self.t
        step_size = f(self.lr / bc1)
        bc2_sqrt = f(math.sqrt(bc2))
        for p, m, v, g in zip(self.p, self.m, self.v, grads):
            if g is None:
                continue
            g = np.asarray(g, dtype=np.float32)
            if self.wd != 0.0:
                g = g + f(self.wd) * p                      # grad.add(param, alpha=weight_decay)
            m += (g - m) * f(1.0 - self.b1)                 # exp_avg.lerp_(grad, 1 - beta1)
            v *= f(self.b2)                                 # exp_avg_sq.mul_(beta2)
            v += f(1.0 - self.b2) * g * g                   #   .addcmul_(grad, grad, value=1 - beta2)
            denom = np.sqrt(v) / bc2_sqrt + f(self.eps)     # (exp_avg_sq.sqrt() / bias_correction2_sqrt).add_(eps)
            p -= step_size * (m / denom)                    # param.addcdiv_(exp_avg, denom, value=-step_size)
        return self.p
