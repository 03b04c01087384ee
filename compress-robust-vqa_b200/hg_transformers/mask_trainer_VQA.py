"""Drop-in for the reference's ``hg_transformers/mask_trainer_VQA.py``: stage-2 trainer with ONE zero
rate for every masked module (``reset_threshold`` uses ``masker_scheduler.init_sparsity``, reference
:470-477).  Also drives stage-1/3 style runs through ``training_type``."""
from masking.maskers import Masker  # noqa: F401

from ._trainer_core import (  # noqa: F401
    CosineLoss, LPF_loss, RUBI_loss, SequentialDistributedSampler, TrainerCore, is_apex_available,
    is_tensorboard_available, is_wandb_available, set_seed,
)
from .trainer_utils import PREFIX_CHECKPOINT_DIR, EvalPrediction, PredictionOutput, TrainOutput  # noqa: F401
from .training_args import TrainingArguments, is_tpu_available  # noqa: F401
from .vqa_debias_loss_functions import *  # noqa: F401,F403


class Trainer(TrainerCore):
    threshold_mode = "global"
    forward_style = "lxmert"
