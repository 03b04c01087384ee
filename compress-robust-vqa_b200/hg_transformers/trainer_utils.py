"""Result tuples of the trainers (reference hg_transformers/trainer_utils.py)."""
from typing import Any, Dict, NamedTuple, Optional

PREFIX_CHECKPOINT_DIR = "checkpoint"


class EvalPrediction(NamedTuple):
    predictions: Any
    label_ids: Any


class PredictionOutput(NamedTuple):
    predictions: Any
    label_ids: Optional[Any]
    metrics: Optional[Dict[str, float]]
    q_ids: Optional[Any]


class TrainOutput(NamedTuple):
    global_step: int
    training_loss: float
