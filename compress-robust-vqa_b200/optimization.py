"""Root ``optimization.AdamW`` of the reference (optimization.py:8-129): Adam with decoupled weight
decay plus a running ``state['sum'] += |grad|`` accumulator.

Same constructor and ``step()`` contract.  On CUDA the per-parameter loop of ~9 elementwise kernels is
replaced by ONE fused kernel per parameter tensor (crv_adamw_step), or one launch over the whole score
arena when the parameters are arena views (hg_transformers._engine.ScoreArena).  A clip coefficient
prepared by the trainer (``set_clip``) is folded into the same pass, so clip_grad_norm_ never rewrites
the gradients in HBM."""

import torch
from torch.optim import Optimizer

from crvqa import ops


class AdamW(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-6, weight_decay=0.0, correct_bias=True,
                 initial_accumulator_value=0, grad_mask=None):
        if lr <= 0.0:
            raise ValueError("Invalid learning rate: {} - should be >= 0.0".format(lr))
        if not 0.0 <= betas[0] < 1.0:
            raise ValueError("Invalid beta parameter: {} - should be in [0.0, 1.0[".format(betas[0]))
        if not 0.0 <= betas[1] < 1.0:
            raise ValueError("Invalid beta parameter: {} - should be in [0.0, 1.0[".format(betas[1]))
        if not 0.0 <= eps:
            raise ValueError("Invalid epsilon value: {} - should be >= 0.0".format(eps))
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, correct_bias=correct_bias)
        super().__init__(params, defaults)
        self.grad_mask = grad_mask
        self.initial_accumulator_value = initial_accumulator_value
        self._clip = None      # (device sum-of-squares tensor, max_norm) set by the trainer for one step
        self._arena = None     # ScoreArena whose flat buffers cover a contiguous run of parameters
        self._hyper = None     # device tensor {lr, step_size}: set while a CUDA graph of the step is in use
        for group in self.param_groups:
            for p in group["params"]:
                state = self.state[p]
                state["step"] = 0
                state["sum"] = torch.full_like(p.data, initial_accumulator_value)

    def get_accumulator(self, returnall=True):
        if returnall:
            return self.state
        return {p: {"step": self.state[p]["step"], "sum": self.state[p]["sum"]}
                for group in self.param_groups for p in group["params"]}

    # -- B200 engine hooks ---------------------------------------------------------------------
    def set_clip(self, total_sumsq, max_norm):
        """Fold clip_grad_norm_(max_norm) into the next step: total_sumsq is a device scalar."""
        self._clip = (total_sumsq, float(max_norm))

    def use_device_hyper(self, hyper):
        """All parameter groups share one lr schedule; {lr, step_size} are read from `hyper` on the device."""
        self._hyper = hyper

    def advance_steps(self, n=1):
        """Bookkeeping for steps executed by CUDA-graph replay (no Python optimiser code runs there)."""
        for group in self.param_groups:
            for p in group["params"]:
                self.state[p]["step"] += n

    def hyper_values(self, next_step):
        """{lr, step_size} of the next step for the device hyper vector.  ONE pair serves every parameter group, so
        the groups must agree on lr / betas / correct_bias (they do under the reference's init_optimizer, which
        builds every group from the same TrainingArguments and steps them with one LambdaLR)."""
        g = self.param_groups[0]
        for other in self.param_groups[1:]:
            if (other["lr"], other["betas"], other["correct_bias"]) != (g["lr"], g["betas"], g["correct_bias"]):
                raise RuntimeError("device-side {lr, step_size} (CUDA-graph replay) needs one lr schedule for all "
                                   "parameter groups; set CRVQA_CUDA_GRAPH=0 for per-group schedules")
        return g["lr"], ops.adam_step_size(g["lr"], next_step, g["betas"][0], g["betas"][1], g["correct_bias"])

    def attach_arena(self, arena):
        """Parameters that are views of `arena` are updated by one flat launch; their state tensors
        (sum / exp_avg / exp_avg_sq) become views of the arena's flat state buffers."""
        self._arena = arena
        for group in self.param_groups:
            for p in group["params"]:
                views = arena.state_views(p)
                if views is not None:
                    st = self.state[p]
                    views["sum"].copy_(st["sum"])
                    st["sum"], st["exp_avg"], st["exp_avg_sq"] = views["sum"], views["exp_avg"], views["exp_avg_sq"]

    def ensure_state(self):
        """Create exp_avg / exp_avg_sq of every trainable parameter now (instead of at its first step)."""
        for group in self.param_groups:
            for p in group["params"]:
                if p.requires_grad:
                    self._ensure_state(p)

    def _ensure_state(self, p):
        state = self.state[p]
        if "exp_avg" not in state:
            state["exp_avg"] = torch.zeros_like(p.data)
            state["exp_avg_sq"] = torch.zeros_like(p.data)
        if state["sum"].device != p.device:  # constructed on CPU, model moved later (reference calls .cuda())
            state["sum"] = state["sum"].to(p.device)
        return state

    def step(self, closure=None):
        loss = closure() if closure is not None else None
        clip_sumsq, max_norm = self._clip if self._clip is not None else (None, 1.0)
        self._clip = None
        arena_done = False
        for group in self.param_groups:
            beta1, beta2 = group["betas"]
            for p in group["params"]:
                if p.grad is None:
                    continue
                if p.grad.is_sparse:
                    raise RuntimeError("Adam does not support sparse gradients, please consider SparseAdam instead")
                if self.grad_mask is not None:
                    p.grad.data.mul_(self.grad_mask[p])
                in_arena = self._arena is not None and self._arena.owns(p)
                state = self._ensure_state(p)
                state["step"] += 1
                if in_arena:
                    if not arena_done:
                        self._arena.adamw_step(lr=group["lr"], step=state["step"], beta1=beta1, beta2=beta2,
                                               eps=group["eps"], weight_decay=group["weight_decay"],
                                               correct_bias=group["correct_bias"], clip_sumsq=clip_sumsq,
                                               max_norm=max_norm, with_sum=self.grad_mask is None,
                                               hyper=self._hyper)
                        arena_done = True
                    continue
                if not p.is_cuda:
                    raise RuntimeError("optimization.AdamW updates parameters with CUDA kernels; got a CPU parameter")
                g = p.grad.data
                if not g.is_contiguous():
                    g = g.contiguous()
                ops.adamw_step_flat(p.data, g, state["exp_avg"], state["exp_avg_sq"],
                                    state["sum"] if self.grad_mask is None else None, group["lr"], state["step"],
                                    beta1, beta2, group["eps"], group["weight_decay"], clip_sumsq, max_norm,
                                    group["correct_bias"], self._hyper)
        return loss
