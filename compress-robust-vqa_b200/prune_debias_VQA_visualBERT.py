"""Stage-2 driver pieces of the reference's ``prune_debias_VQA_visualBERT.py`` (BASELINE config 3), plus a
synthetic-data entry point.

Kept with the reference's names and semantics: ``init_masker(conf, model, logger)`` (:125-190 -- ONE uniform zero rate
from the scheduler, weight types K/Q/V/AO/I/O/P/E over ``layers_to_mask``, the VisualBERT masker) and, shared with the
LXMERT driver, ``init_optimizer`` / ``ModelArguments``.  The reference driver needs the VQA-CP features and a stage-1
checkpoint; ``main()`` runs the same flow on synthetic VQA-shaped tensors and random init through the visualBERT trainer.
"""
import argparse
import logging

import torch

from hg_transformers.training_args import TrainingArguments
from masking import maskers_visualBert as maskers
from masking import sparsity_control as sp_control
from prune_debias_VQA import DEFAULT_SCHEDULER_CONF, ModelArguments, SyntheticVQADataset, init_optimizer  # noqa: F401
from utils import param_parser

logger = logging.getLogger(__name__)

VISUALBERT_WEIGHT_TYPES = ["K", "Q", "V", "AO", "I", "O", "P", "E"]


def init_masker(conf, model, logger):
    conf.masking_scheduler_conf_ = (param_parser.dict_parser(conf.masking_scheduler_conf)
                                    if conf.masking_scheduler_conf is not None else None)
    conf.masking_scheduler_conf_["final_sparsity"] = conf.zero_rate
    for k, v in conf.masking_scheduler_conf_.items():
        setattr(conf, f"masking_scheduler_{k}", v)
    conf.logger = logger
    masker = maskers.Masker(
        masker_scheduler=sp_control.MaskerScheduler(conf), logger=logger, mask_biases=conf.mask_biases,
        structured_masking_info={"structured_masking": conf.structured_masking,
                                 "structured_masking_types": conf.structured_masking_types,
                                 "force_masking": conf.force_masking},
        threshold=conf.threshold, init_scale=conf.init_scale, which_ptl=conf.model_type,
        controlled_init=conf.controlled_init)
    assert conf.layers_to_mask is not None, "Please specify which BERT layers to mask."
    conf.layers_to_mask_ = [int(x) for x in str(conf.layers_to_mask).split(",")]
    names_tobe_masked = maskers.chain_module_names(conf.model_type, conf.layers_to_mask_, VISUALBERT_WEIGHT_TYPES)
    if conf.mask_classifier:
        # the reference reads the non-existent `conf.type` here and dies with AttributeError (:176-181); say why instead
        raise AssertionError("--mask_classifier is not reachable in the reference (it reads conf.type, which does not "
                             "exist); the classifier is trained, not masked")
    masker.patch_modules(model=model, names_tobe_masked=names_tobe_masked, name_of_masker=conf.name_of_masker)
    return masker


def build_stage2(ans_num=3129, model_args=None, device=None, seed=49, config_kwargs=None, quiet=True):
    """Model + masker for a synthetic VisualBERT stage-2 run: random init under `seed`, moved to `device`, then patched
    (so the magnitude init runs on the GPU)."""
    from hg_transformers.modeling_visualbert import VisualBertForMultipleChoice, visualBERTConfig
    model_args = model_args or ModelArguments(model_type="visual_bert")
    torch.manual_seed(seed)
    model = VisualBertForMultipleChoice(visualBERTConfig(ans_num=ans_num, **(config_kwargs or {})))
    if device is not None:
        model = model.to(device)
    log = logging.getLogger("crvqa.masker")
    if quiet:
        log.setLevel(logging.WARNING)
    return model, init_masker(model_args, model, log), model_args


def main(argv=None):
    ap = argparse.ArgumentParser(description="synthetic stage-2 mask training (VisualBERT, uniform zero rate)")
    ap.add_argument("--per_gpu_train_batch_size", type=int, default=32)
    ap.add_argument("--max_steps", type=int, default=4)
    ap.add_argument("--logging_steps", type=int, default=2)
    ap.add_argument("--ans_num", type=int, default=3129)
    ap.add_argument("--zero_rate", type=float, default=0.7)
    ap.add_argument("--learning_rate", type=float, default=5e-5)
    ap.add_argument("--num_examples", type=int, default=256)
    ap.add_argument("--output_dir", default="./out_stage2_visualbert")
    ap.add_argument("--seed", type=int, default=49)
    a = ap.parse_args(argv)
    from hg_transformers.data.data_collator import TrimCollator
    from hg_transformers.data.metrics import vqa_compute_metrics
    from hg_transformers.mask_trainer_visualBERT_VQA import Trainer
    targs = TrainingArguments(output_dir=a.output_dir, per_gpu_train_batch_size=a.per_gpu_train_batch_size,
                              max_steps=a.max_steps, logging_steps=a.logging_steps, seed=a.seed, Masker_type="normal",
                              training_type="Masker", save_steps=0, learning_rate=a.learning_rate)
    margs = ModelArguments(model_type="visual_bert", zero_rate=a.zero_rate)
    model, masker, margs = build_stage2(a.ans_num, model_args=margs, device=targs.device, seed=a.seed)
    data = SyntheticVQADataset(a.num_examples, a.ans_num, seed=a.seed)
    opt = init_optimizer(model, targs, len(data))
    trainer = Trainer(model=model, args=targs, model_args=margs, data_collator=TrimCollator(), train_dataset=data,
                      eval_dataset=None, compute_metrics=vqa_compute_metrics, optimizers=opt, masker=masker)
    out = trainer.train()
    print(out[0])


if __name__ == "__main__":
    main()
