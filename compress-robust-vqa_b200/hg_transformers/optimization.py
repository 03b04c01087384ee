"""Learning-rate schedules used by the stage-2 drivers (reference hg_transformers/optimization.py).
AdamW itself lives in the root ``optimization`` module, as in the reference."""
from torch.optim.lr_scheduler import LambdaLR

from optimization import AdamW  # noqa: F401  (root module of this package tree)


def get_constant_schedule(optimizer, last_epoch=-1):
    return LambdaLR(optimizer, lambda _: 1, last_epoch=last_epoch)


def linear_schedule_factor(step, num_warmup_steps, num_training_steps):
    if step < num_warmup_steps:
        return float(step) / float(max(1, num_warmup_steps))
    return max(0.0, float(num_training_steps - step) / float(max(1, num_training_steps - num_warmup_steps)))


def get_linear_schedule_with_warmup(optimizer, num_warmup_steps, num_training_steps, last_epoch=-1):
    """Linear warm-up from 0 then linear decay to 0 at num_training_steps."""
    return LambdaLR(optimizer, lambda s: linear_schedule_factor(s, num_warmup_steps, num_training_steps), last_epoch)
