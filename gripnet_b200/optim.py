"""Fused optimiser step for the training scripts' ``torch.optim.Adam(model.parameters(), lr)``
(``GripNet-pose.py:104,146``; ``GripNet-aminer.py:113,135``).

``Adam`` keeps torch's constructor / ``step()`` / ``zero_grad()`` surface and its ``state_dict`` layout
(``exp_avg``, ``exp_avg_sq``, ``step`` per parameter), but ``step()`` is ONE ``gn_adam_step`` call: all
parameter tensors are updated by a single multi-tensor kernel, and the step counter lives in device memory so
that a ``step()`` captured in a CUDA graph (``capture.CapturedStep(post_backward=opt.step)``) is a new
optimiser step on every replay.  No CPU fallback: parameters must be CUDA fp32 tensors.
"""
import ctypes as C

import torch

from . import _lib
from .graph import _stream


class Adam:
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, amsgrad=False):
        if amsgrad:
            raise NotImplementedError("gripnet_b200.optim.Adam: amsgrad is not used by the reference scripts")
        if not 0.0 <= lr or not 0.0 <= eps or not 0.0 <= betas[0] < 1.0 or not 0.0 <= betas[1] < 1.0 \
                or not 0.0 <= weight_decay:
            raise ValueError("invalid Adam hyper-parameter")
        self.params = [p for p in params]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        for p in self.params:
            if not p.is_cuda:
                raise RuntimeError("gripnet_b200.optim.Adam: parameters must be CUDA tensors (no CPU fallback exists)")
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise RuntimeError("gripnet_b200.optim.Adam: parameters must be contiguous fp32 tensors")
        self.lr, self.betas, self.eps, self.weight_decay = float(lr), (float(betas[0]), float(betas[1])), float(eps), \
            float(weight_decay)
        dev = self.params[0].device
        self.exp_avg = [torch.zeros_like(p) for p in self.params]
        self.exp_avg_sq = [torch.zeros_like(p) for p in self.params]
        self._step = torch.zeros(1, dtype=torch.int64, device=dev)       # device-resident step counter
        self._lib = _lib.load()

    # -- torch.optim surface -------------------------------------------------------------------
    def zero_grad(self, set_to_none=True):
        for p in self.params:
            if p.grad is None:
                continue
            if set_to_none:
                p.grad = None
            else:
                p.grad.detach_()
                p.grad.zero_()

    @torch.no_grad()
    def step(self):
        """One Adam update of every parameter that has a gradient (parameters without one are skipped,
        as torch does)."""
        live = [(p, m, v) for p, m, v in zip(self.params, self.exp_avg, self.exp_avg_sq) if p.grad is not None]
        if not live:
            return
        table = (_lib.GnAdamTensor * len(live))()
        for i, (p, m, v) in enumerate(live):
            g = p.grad
            if not g.is_cuda or g.dtype != torch.float32 or g.shape != p.shape:
                raise RuntimeError("gripnet_b200.optim.Adam: gradients must be CUDA fp32 tensors of the parameter's shape")
            if not g.is_contiguous():
                g = p.grad = g.contiguous()
            table[i] = _lib.GnAdamTensor(p.data_ptr(), g.data_ptr(), m.data_ptr(), v.data_ptr(), p.numel())
        _lib.check(self._lib.gn_adam_step(C.cast(table, C.c_void_p), len(live), self.lr, self.betas[0], self.betas[1],
                                          self.eps, self.weight_decay, self._step.data_ptr(), _stream()),
                   "gn_adam_step")

    @property
    def step_count(self):
        return int(self._step.item())

    def state_dict(self):
        """torch.optim.Adam's layout, so a checkpoint moves either way."""
        t = self._step.to(torch.float32).cpu().reshape(())
        return {
            "state": {i: {"step": t.clone(), "exp_avg": m, "exp_avg_sq": v}
                      for i, (m, v) in enumerate(zip(self.exp_avg, self.exp_avg_sq))},
            "param_groups": [{"lr": self.lr, "betas": self.betas, "eps": self.eps, "weight_decay": self.weight_decay,
                              "amsgrad": False, "params": list(range(len(self.params)))}],
        }

    def load_state_dict(self, sd):
        g = sd["param_groups"][0]
        self.lr, self.betas, self.eps = float(g["lr"]), (float(g["betas"][0]), float(g["betas"][1])), float(g["eps"])
        self.weight_decay = float(g.get("weight_decay", 0.0))
        steps = set()
        for i, (m, v) in enumerate(zip(self.exp_avg, self.exp_avg_sq)):
            st = sd["state"].get(i)
            if st is None:
                m.zero_()
                v.zero_()
                continue
            m.copy_(st["exp_avg"])
            v.copy_(st["exp_avg_sq"])
            steps.add(int(st["step"]))
        if len(steps) > 1:
            raise RuntimeError("gripnet_b200.optim.Adam keeps one step counter: per-parameter steps differ")
        self._step.fill_(steps.pop() if steps else 0)
