"""Drop-in for the reference's ``hg_transformers/mask_trainer_visualBERT_VQA.py``: the baseline
trainer calling ``model(input_ids=..., visual_embeds=..., labels=...)`` (reference :820, :1150)."""
from masking.maskers_visualBert import Masker  # noqa: F401

from ._trainer_core import (  # noqa: F401
    CosineLoss, LPF_loss, RUBI_loss, SequentialDistributedSampler, TrainerCore, is_apex_available,
    is_tensorboard_available, is_wandb_available, set_seed,
)
from .trainer_utils import PREFIX_CHECKPOINT_DIR, EvalPrediction, PredictionOutput, TrainOutput  # noqa: F401
from .training_args import TrainingArguments, is_tpu_available  # noqa: F401
from .vqa_debias_loss_functions import *  # noqa: F401,F403


class Trainer(TrainerCore):
    threshold_mode = "global"
    forward_style = "visualbert"
