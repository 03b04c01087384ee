"""Learning-rate schedule of the mPLUG driver (reference mPLUG/scheduler/: ``create_scheduler`` with the shipped
``sched: cosine`` setting -> ``CosineLRScheduler``, a timm-style scheduler that is TOLD the epoch on every call:
``scheduler.step(i // step_size)`` during the warm-up of epoch 0 and ``lr_scheduler.step(epoch + warmup_steps)`` at
every epoch start, mPLUG/vqa_mplug.py:196,437).  Host-side Python; the tanh / step / plateau variants and the
learning-rate noise option of the reference are not built (no shipped configuration selects them)."""
import math


class CosineLRScheduler:
    """lr(t) = warm-up line from ``warmup_lr_init`` to the group's base lr over ``warmup_t`` calls, then (counted from the
    end of the warm-up) cosine decay from base to ``lr_min`` over ``t_initial``, restarted ``cycle_limit`` times with
    period x ``t_mul`` and amplitude x ``decay_rate``; ``lr_min`` afterwards."""

    def __init__(self, optimizer, t_initial, t_mul=1.0, lr_min=0.0, decay_rate=1.0, warmup_t=0, warmup_lr_init=0,
                 warmup_prefix=True, cycle_limit=0, t_in_epochs=True, noise_range_t=None, noise_pct=0.67,
                 noise_std=1.0, noise_seed=42, initialize=True):
        if noise_range_t is not None:
            raise NotImplementedError("learning-rate noise is not built")
        assert t_initial > 0 and lr_min >= 0
        self.optimizer = optimizer
        for i, group in enumerate(optimizer.param_groups):
            if initialize:
                if "lr" not in group:
                    raise KeyError(f"lr missing from param_groups[{i}]")
                group.setdefault("initial_lr", group["lr"])
            elif "initial_lr" not in group:
                raise KeyError(f"initial_lr missing from param_groups[{i}]")
        self.base_values = [g["initial_lr"] for g in optimizer.param_groups]
        self.t_initial, self.t_mul, self.lr_min, self.decay_rate = t_initial, t_mul, lr_min, decay_rate
        self.cycle_limit, self.warmup_t, self.warmup_lr_init = cycle_limit, warmup_t, warmup_lr_init
        self.warmup_prefix, self.t_in_epochs = warmup_prefix, t_in_epochs
        self.metric = None
        if warmup_t:
            self.warmup_steps = [(v - warmup_lr_init) / warmup_t for v in self.base_values]
            self.update_groups(warmup_lr_init)
        else:
            self.warmup_steps = [1 for _ in self.base_values]
            self.update_groups(self.base_values)

    def update_groups(self, values):
        if not isinstance(values, (list, tuple)):
            values = [values] * len(self.optimizer.param_groups)
        for group, value in zip(self.optimizer.param_groups, values):
            group["lr"] = value

    def _cycle(self, t):
        """(index of the cosine cycle that contains t, its length, position inside it)."""
        period, growth = self.t_initial, self.t_mul
        if growth == 1:
            index = t // period
            return index, period, t - period * index
        index = math.floor(math.log(1 - t / period * (1 - growth), growth))
        start = (1 - growth ** index) / (1 - growth) * period
        return index, growth ** index * period, t - start

    def _get_lr(self, t):
        if t < self.warmup_t:                                   # the warm-up line
            return [self.warmup_lr_init + t * slope for slope in self.warmup_steps]
        index, length, pos = self._cycle(t - self.warmup_t if self.warmup_prefix else t)
        if self.cycle_limit != 0 and index >= self.cycle_limit:  # past the last restart
            return [self.lr_min] * len(self.base_values)
        shrink = self.decay_rate ** index
        floor = self.lr_min * shrink
        wave = 0.5 * (1 + math.cos(math.pi * pos / length))
        return [floor + (base * shrink - floor) * wave for base in self.base_values]

    def step(self, epoch, metric=None):
        self.metric = metric
        if self.t_in_epochs:
            self.update_groups(self._get_lr(epoch))

    def step_update(self, num_updates, metric=None):
        self.metric = metric
        if not self.t_in_epochs:
            self.update_groups(self._get_lr(num_updates))

    def get_cycle_length(self, cycles=0):
        cycles = max(1, cycles or self.cycle_limit)
        if self.t_mul == 1.0:
            return self.t_initial * cycles
        return int(math.floor(-self.t_initial * (self.t_mul ** cycles - 1) / (1 - self.t_mul)))

    def state_dict(self):
        return {k: v for k, v in self.__dict__.items() if k != "optimizer"}

    def load_state_dict(self, state):
        self.__dict__.update(state)


def create_scheduler(args, optimizer):
    """(scheduler, number of epochs) from the ``schedular`` block of the YAML config (attribute access)."""
    if getattr(args, "lr_noise", None) is not None:
        raise NotImplementedError("learning-rate noise is not built")
    if args.sched != "cosine":
        raise NotImplementedError(f"sched={args.sched!r}: only the shipped 'cosine' schedule is built")
    scheduler = CosineLRScheduler(
        optimizer, t_initial=args.epochs, t_mul=getattr(args, "lr_cycle_mul", 1.0), lr_min=args.min_lr,
        decay_rate=args.decay_rate, warmup_lr_init=args.warmup_lr, warmup_t=args.warmup_epochs,
        cycle_limit=getattr(args, "lr_cycle_limit", 1), t_in_epochs=True)
    return scheduler, scheduler.get_cycle_length() + args.cooldown_epochs
