"""Settings of mPLUG mask training: the attribute bag ``vqa_mplug.py`` creates and edits in place (reference
mPLUG/masking/mask_config.py:1-21 -- same field names and default values)."""

_SCHEDULER = ",".join(["lambdas_lr=0", "sparsity_warmup=automated_gradual_sparsity",
                       "sparsity_warmup_interval_epoch=0.1", "init_epoch=0", "final_epoch=1"])

# controlled_init is one of: magnitude, uniform, magnitude_and_uniform, double_uniform, magnitude_soft
_DEFAULTS = dict(
    zero_rate=0.5, threshold=1e-2, init_scale=2e-2,
    mask_classifier=False, mask_biases=False, train_classifier=True, global_prune=False,
    force_masking="bert", controlled_init="magnitude_soft", name_of_masker="MaskedLinear1",
    structured_masking=None, structured_masking_types=None,
    masking_scheduler_conf=_SCHEDULER, init_sparsity=None, final_sparsity_epoch=1, masker_update_step=100,
    load_mask_from=None,
)


class MaskConfigs:
    def __init__(self) -> None:
        self.__dict__.update(_DEFAULTS)
