"""Drop-in for the reference's ``hg_transformers/global_mask_trainer_VQA.py``: the stage-2 trainer whose
``reset_threshold`` takes ONE threshold for every masked module, the k-th smallest score of the union of all score
tensors with k = int(total * tgt_sparsity) (reference :421-443; requires ``model_args.global_prune``)."""
from masking.global_maskers import Masker  # noqa: F401

from ._trainer_core import (  # noqa: F401
    CosineLoss, LPF_loss, RUBI_loss, SequentialDistributedSampler, TrainerCore, is_apex_available,
    is_tensorboard_available, is_wandb_available, set_seed,
)
from .trainer_utils import PREFIX_CHECKPOINT_DIR, EvalPrediction, PredictionOutput, TrainOutput  # noqa: F401
from .training_args import TrainingArguments, is_tpu_available  # noqa: F401
from .vqa_debias_loss_functions import *  # noqa: F401,F403


class Trainer(TrainerCore):
    threshold_mode = "union"
    forward_style = "lxmert"
