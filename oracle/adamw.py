"""Oracle restatement of clip_grad_norm_ + the reference's root AdamW (optimization.py:66-129)."""
import math

import torch


def clip_coef(grads, max_norm=1.0):
    """torch.nn.utils.clip_grad_norm_ as called at hg_transformers/mask_trainer_VQA.py:649."""
    total = torch.sqrt(sum((g.double() ** 2).sum() for g in grads)).float()
    coef = max_norm / (total + 1e-6)
    return torch.clamp(coef, max=1.0), total


def adamw_step(p, g, state, lr, beta1=0.9, beta2=0.999, eps=1e-8, weight_decay=0.0, correct_bias=True):
    """One parameter of AdamW.step -- optimization.py:78-127.  state: dict(step, sum, exp_avg, exp_avg_sq)."""
    state["step"] += 1
    state["sum"].add_(g.abs())
    state["exp_avg"].mul_(beta1).add_(g, alpha=1.0 - beta1)
    state["exp_avg_sq"].mul_(beta2).addcmul_(g, g, value=1.0 - beta2)
    denom = state["exp_avg_sq"].sqrt().add_(eps)
    step_size = lr
    if correct_bias:
        step_size = lr * math.sqrt(1.0 - beta2 ** state["step"]) / (1.0 - beta1 ** state["step"])
    p.addcdiv_(state["exp_avg"], denom, value=-step_size)
    if weight_decay > 0.0:
        p.add_(p, alpha=-lr * weight_decay)


def new_state(p):
    return {"step": 0, "sum": torch.zeros_like(p), "exp_avg": torch.zeros_like(p), "exp_avg_sq": torch.zeros_like(p)}
