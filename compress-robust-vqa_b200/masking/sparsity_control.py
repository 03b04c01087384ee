"""Host-side sparsity schedule (reference masking/sparsity_control.py:10-240).  Pure Python: in the
LXMERT / VisualBERT flows only ``MaskerScheduler.init_sparsity`` is read; ``step`` drives mPLUG."""
from ._core import MaskedLinearX


def automated_gradual_sparsity(init_sparsity, final_sparsity, interval_epoch, init_epoch, final_epoch):
    """Cubic ramp of Zhu & Gupta 2017 from init_sparsity (at init_epoch) to final_sparsity (at final_epoch)."""
    span = final_epoch - init_epoch

    def f(current_epoch, current_sparsity):
        if current_epoch > final_epoch or span == 0:
            return final_sparsity
        remaining = 1.0 - (1.0 * (current_epoch - init_epoch) / span)
        return final_sparsity + (init_sparsity - final_sparsity) * remaining ** 3

    return f


def stepwise_sparsity(init_sparsity, final_sparsity, interval_epoch, init_epoch, final_epoch,
                      sparsity_incremental_ratio, with_safety_check=True):
    """Every interval_epoch, prune a fixed ratio of what is left."""

    def f(current_epoch, current_sparsity):
        if current_epoch < init_epoch:
            return init_sparsity
        if current_epoch >= final_epoch:
            return final_sparsity
        if (current_epoch - init_epoch) % interval_epoch <= 1e-5:
            return current_sparsity + (1 - current_sparsity) * sparsity_incremental_ratio
        return current_sparsity

    if with_safety_check:
        reached = init_sparsity
        for epoch in range(init_epoch, final_epoch, interval_epoch):
            reached = f(epoch, reached)
        if (final_epoch - init_epoch) % interval_epoch <= 1e-5:
            reached += (1 - reached) * sparsity_incremental_ratio
        if reached < final_sparsity:
            raise ValueError(
                "Increase initial sparsity and/or incremental ratio,"
                + "current final sparsity is {}, required value is {}".format(reached, final_sparsity))
    return f


class MaskerScheduler(object):
    def __init__(self, conf):
        self.conf = conf
        self.masking_scheduler_conf_ = conf.masking_scheduler_conf_
        self._current_sparsity = 0
        c = conf.masking_scheduler_conf_
        if c is not None:
            assert "final_sparsity" in c
            assert "sparsity_warmup_interval_epoch" in c
            self.init_sparsity = c["init_sparsity"] if "init_sparsity" in c else c["final_sparsity"]
            self.get_sparsity_fn = self._get_pruner()
        else:
            self.init_sparsity = 0.5
            self.get_sparsity_fn = None

    @property
    def is_skip(self):
        c = self.conf.masking_scheduler_conf_
        return self.get_sparsity_fn is None or ("lambdas_lr" in c and c["lambdas_lr"] == 0)

    def _epoch_bounds(self):
        c = self.masking_scheduler_conf_
        init_epoch = c["init_epoch"] if "init_epoch" in c else int(self.conf.num_epochs * 0.1)
        final_epoch = c["final_epoch"] if "final_epoch" in c else int(self.conf.num_epochs * 0.8)
        return init_epoch, final_epoch

    def _get_pruner(self):
        c = self.masking_scheduler_conf_
        kind = c.get("sparsity_warmup", "automated_gradual_sparsity")
        init_epoch, final_epoch = self._epoch_bounds()
        common = dict(init_sparsity=self.init_sparsity, final_sparsity=c["final_sparsity"],
                      interval_epoch=c["sparsity_warmup_interval_epoch"], init_epoch=init_epoch,
                      final_epoch=final_epoch)
        if kind == "automated_gradual_sparsity":
            self.conf.logger.info("use automated_gradual_sparsity.")
            return automated_gradual_sparsity(**common)
        if kind == "stepwise_sparsity":
            self.conf.logger.info("use stepwise pruner.")
            assert "sparsity_incremental_ratio" in c
            return stepwise_sparsity(sparsity_incremental_ratio=c["sparsity_incremental_ratio"], **common)
        raise NotImplementedError

    def step(self, cur_epoch):
        self.cur_epoch = cur_epoch
        target = self.get_sparsity_fn(cur_epoch, self._current_sparsity)
        final = self.masking_scheduler_conf_["final_sparsity"]
        lo, hi = (self.init_sparsity, final) if final > self.init_sparsity else (final, self.init_sparsity)
        self.target_sparsity = min(hi, max(target, lo))
        incremental = (self.target_sparsity - self._current_sparsity) / (1 - self._current_sparsity)
        return incremental, self.target_sparsity, self.is_sparsity_change()

    def is_meet_sparsity(self):
        return self.target_sparsity >= self.masking_scheduler_conf_["final_sparsity"]

    def is_sparsity_change(self):
        if self._current_sparsity == self.target_sparsity:
            return False
        self._current_sparsity = self.target_sparsity
        return True

    def get_sparsity_over_whole_model(self, model, masker, trainable=True):
        """1 - nnz/total over the masks of every MaskedLinearX in the model."""

        def collect(mod):
            found = []
            for child in mod.children():
                if isinstance(child, MaskedLinearX):
                    found.append(child)
                else:
                    found.extend(collect(child))
            return found

        nnz, tot = 0, 0
        for module in collect(model):
            for mask in module.get_masks():
                if mask is not None:
                    nnz = nnz + mask.sum()
                    tot = tot + mask.numel()
        return 1 - nnz / tot
