"""Drop-in for the reference's ``hg_transformers/mask_trainer_Robust_VQA.py``: per-modality zero rates.
``reset_threshold`` reads ``masker.name_in_module[name]`` -> modality -> ``masker.hpmodel.zerorate_dict``
(reference :467-482) and ``save_model_mask`` logs the zero rate of each modality (:943-991)."""
from masking.maskers import Masker  # noqa: F401

from ._trainer_core import (  # noqa: F401
    CosineLoss, LPF_loss, RUBI_loss, SequentialDistributedSampler, TrainerCore, is_apex_available,
    is_tensorboard_available, is_wandb_available, set_seed,
)
from .trainer_utils import PREFIX_CHECKPOINT_DIR, EvalPrediction, PredictionOutput, TrainOutput  # noqa: F401
from .training_args import TrainingArguments, is_tpu_available  # noqa: F401
from .vqa_debias_loss_functions import *  # noqa: F401,F403


class Trainer(TrainerCore):
    threshold_mode = "modal"
    forward_style = "lxmert"
