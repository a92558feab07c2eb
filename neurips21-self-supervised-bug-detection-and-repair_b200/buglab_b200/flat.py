"""Flat fp32 parameter / gradient space and the fused Adam that runs on it.

All parameters of the module live in ONE contiguous buffer (each ``p.data`` is a view), all gradients in a
second one (each ``p.grad`` is a view).  That buffer pair is (i) the operand of the fused clip+Adam kernels
(``bl_grad_sqnorm`` / ``bl_adam_step``: 2 launches per step instead of ~10 per parameter) and (ii) the single
NCCL all-reduce bucket of data-parallel training.  Reference optimiser being replaced:
``torch.optim.Adam(p, lr=1e-4)`` + ``clip_grad_norm_(0.5)`` (buglab/models/utils.py:51-52, train.py:104).
"""
from typing import Iterable, List, Optional

import torch

from . import ops


class FlatAdam(torch.optim.Optimizer):
    """Adam (torch defaults, no weight decay / amsgrad) over a flat buffer, with fused global-norm clipping.

    It is a ``torch.optim.Optimizer`` so LR schedulers (the reference's ``LambdaLR`` warm-up,
    buglab/models/utils.py:55-66) drive ``param_groups[0]['lr']`` unchanged.
    """

    fused_clip = True  # the trainer must not clip again

    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 max_grad_norm: Optional[float] = None):
        params = [p for p in params if p.requires_grad]
        if not params:
            raise ValueError("FlatAdam got no parameters")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))
        self.max_grad_norm = max_grad_norm
        self._params: List[torch.nn.Parameter] = params
        device = params[0].device
        if device.type != "cuda":
            raise RuntimeError("FlatAdam needs CUDA parameters; buglab_b200 has no CPU fallback")
        sizes = [p.numel() for p in params]
        # every view starts on a 16-byte boundary so kernels can use 128-bit accesses on it
        self._offsets, total = [], 0
        for n in sizes:
            self._offsets.append(total)
            total += (n + 3) // 4 * 4
        self.flat_param = torch.zeros(total, device=device, dtype=torch.float32)
        self.flat_grad = torch.zeros(total, device=device, dtype=torch.float32)
        self.exp_avg = torch.zeros(total, device=device, dtype=torch.float32)
        self.exp_avg_sq = torch.zeros(total, device=device, dtype=torch.float32)
        self._sqnorm = torch.zeros(1, device=device, dtype=torch.float32)
        self._partial = torch.empty(1024, device=device, dtype=torch.float32)
        self._step = 0
        with torch.no_grad():
            for p, off in zip(params, self._offsets):
                if p.dtype != torch.float32:
                    raise RuntimeError("FlatAdam expects fp32 parameters")
                view = self.flat_param[off: off + p.numel()].view_as(p)
                view.copy_(p.data)
                p.data = view
        self._attach_grads()

    def _attach_grads(self) -> None:
        for p, off in zip(self._params, self._offsets):
            p.grad = self.flat_grad[off: off + p.numel()].view_as(p)

    def zero_grad(self, set_to_none: bool = True) -> None:  # noqa: D102 - keeps the views alive
        self.flat_grad.zero_()
        if any(p.grad is None or p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off
               for p, off in zip(self._params, self._offsets)):
            self._attach_grads()

    @property
    def num_steps(self) -> int:
        return self._step

    def gradient_reducer(self, bucket_bytes: Optional[int] = None):
        """The bucketed all-reduce of this optimiser's flat gradient, overlapped with backward (created once).
        ``BUGLAB_B200_ALLREDUCE_BUCKET_MB`` overrides the bucket size (a value >= the gradient's size gives one bucket,
        reduced after backward)."""
        if getattr(self, "_reducer", None) is None:
            import os

            from .distributed import OverlappedGradientReducer

            if bucket_bytes is None:
                bucket_bytes = int(float(os.environ.get("BUGLAB_B200_ALLREDUCE_BUCKET_MB", "32")) * (1 << 20))

            self._reducer = OverlappedGradientReducer(self.flat_grad, self._params, self._offsets, bucket_bytes)
        return self._reducer

    def grad_norm(self) -> torch.Tensor:
        """Global L2 norm of the current flat gradient (device scalar; no sync)."""
        ops.grad_sqnorm(self.flat_grad, self._sqnorm, self._partial)
        return self._sqnorm.sqrt()

    @torch.no_grad()
    def step(self, closure=None, grad_scale: float = 1.0):
        if closure is not None:
            raise NotImplementedError("FlatAdam does not support closures")
        # autograd may have replaced a .grad view (e.g. first backward after set_to_none): fold it back
        for p, off in zip(self._params, self._offsets):
            if p.grad is not None and p.grad.data_ptr() != self.flat_grad.data_ptr() + 4 * off:
                self.flat_grad[off: off + p.numel()].view_as(p).copy_(p.grad)
                p.grad = self.flat_grad[off: off + p.numel()].view_as(p)
        group = self.param_groups[0]
        self._step += 1
        max_norm = float(self.max_grad_norm) if self.max_grad_norm else 0.0
        if max_norm > 0:
            ops.grad_sqnorm(self.flat_grad, self._sqnorm, self._partial)
        ops.adam_step(self.flat_param, self.flat_grad, self.exp_avg, self.exp_avg_sq, float(group["lr"]),
                      float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]), self._step,
                      max_norm, self._sqnorm if max_norm > 0 else None, float(grad_scale))
        return None

    # pickling an optimizer mid-training is not part of the reference flow (optimizer state is never saved,
    # SURVEY.md §5); state_dict only carries the step count and moments for completeness.
    def state_dict(self):
        return {"step": self._step, "exp_avg": self.exp_avg, "exp_avg_sq": self.exp_avg_sq,
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, state):
        self._step = int(state["step"])
        self.exp_avg.copy_(state["exp_avg"])
        self.exp_avg_sq.copy_(state["exp_avg_sq"])
