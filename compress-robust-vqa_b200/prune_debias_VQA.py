"""Stage-2 driver pieces of the reference's ``prune_debias_VQA.py`` that the hot path needs, plus a
synthetic-data entry point.

Kept with the reference's names and semantics: ``HPmodel_modal`` (:369-384), ``init_masker`` (:261-336),
``init_optimizer`` (:612-631), the ``ModelArguments`` fields the masker reads (:390-520).  The
reference driver itself needs the VQA-CP datasets and stage-1 checkpoints (not shipped, SURVEY.md
section 2.1 #13/#15); ``main()`` here runs the same flow on synthetic VQA-shaped tensors and random
init, which is also what bench.py and the parity tests use.
"""
import argparse
import logging
from dataclasses import dataclass
from typing import Optional

import torch
from torch.utils.data import Dataset

from hg_transformers.optimization import get_linear_schedule_with_warmup
from hg_transformers.training_args import TrainingArguments
from masking import maskers_Robust as maskers
from masking import sparsity_control_Robust as sp_control
from optimization import AdamW
from utils import param_parser

logger = logging.getLogger(__name__)

LXMERT_WEIGHT_TYPES = ["E", "VV", "VB", "lK", "lQ", "lV", "lAO", "lI", "lO", "vK", "vQ", "vV", "vAO", "vI", "vO",
                       "vlVK", "vlVQ", "vlVV", "vlVAO", "vlLaK", "vlLaQ", "vlLaV", "vlLaAO", "vlVaK", "vlVaQ",
                       "vlVaV", "vlVaAO", "vlLi", "vlLo", "vlVi", "vlVo", "P"]

DEFAULT_SCHEDULER_CONF = ("lambdas_lr=0,sparsity_warmup=automated_gradual_sparsity,"
                          "sparsity_warmup_interval_epoch=0.1,init_epoch=0,final_epoch=1")


@dataclass
class ModelArguments:
    """The masker-related fields of the reference's ModelArguments (:390-520), same names and defaults."""
    model_type: str = "lxmert"
    masker_level: str = "modal"
    zero_rate: float = 0.7
    Lang_comp: float = 0.3
    Vis_comp: float = 0.3
    Fus_comp: float = 0.3
    threshold: float = 1e-2
    init_scale: float = 2e-2
    mask_classifier: bool = False
    mask_biases: bool = False
    force_masking: str = "bert"
    controlled_init: Optional[str] = "magnitude"
    structured_masking: Optional[str] = None
    structured_masking_types: Optional[str] = None
    structured: bool = False
    global_prune: bool = False
    name_of_masker: str = "MaskedLinear1"
    layers_to_mask: str = "0,1,2,3,4,5,6,7,8,9,10,11"
    masking_scheduler_conf: Optional[str] = DEFAULT_SCHEDULER_CONF
    output_mask_dir: Optional[str] = None


class HPmodel_modal(torch.nn.Module):
    """Per-modality zero rates: Lang / Vis / Fus from the compression ratios, P from the global zero rate."""

    def __init__(self, Lang, Vis, Fus, P):
        super().__init__()
        self.Lang, self.Vis, self.Fus, self.P = Lang, Vis, Fus, P
        self.module_list = [Lang, Vis, Fus, P]
        self.weight_types = ["Lang", "Vis", "Fus", "P"]
        self.zerorate_dict = dict(zip(self.weight_types, self.module_list))


def init_masker(conf, model, logger, hpmodel, model_args):
    conf.masking_scheduler_conf_ = (param_parser.dict_parser(conf.masking_scheduler_conf)
                                    if conf.masking_scheduler_conf is not None else None)
    conf.masking_scheduler_conf_["final_sparsity"] = conf.zero_rate
    for k, v in conf.masking_scheduler_conf_.items():
        setattr(conf, f"masking_scheduler_{k}", v)
    conf.logger = logger
    masker_scheduler = sp_control.MaskerScheduler(conf)
    masker = maskers.Masker(
        hpmodel=hpmodel, masker_scheduler=masker_scheduler, logger=logger, mask_biases=conf.mask_biases,
        structured_masking_info={"structured_masking": conf.structured_masking,
                                 "structured_masking_types": conf.structured_masking_types,
                                 "force_masking": conf.force_masking},
        threshold=conf.threshold, init_scale=conf.init_scale, which_ptl=conf.model_type,
        controlled_init=conf.controlled_init)
    assert conf.layers_to_mask is not None, "Please specify which BERT layers to mask."
    conf.layers_to_mask_ = [int(x) for x in str(conf.layers_to_mask).split(",")]
    names_tobe_masked, name_in_modal, name_in_module, name_in_layer = maskers.chain_module_names(
        conf.model_type, conf.layers_to_mask_, LXMERT_WEIGHT_TYPES)
    masker.names_tobe_masked = names_tobe_masked
    if model_args.masker_level == "modal":
        masker.name_in_module = name_in_modal
    else:
        raise AssertionError("only masker_level='modal' is reachable in the reference (:323-330)")
    masker.name_of_masker = conf.name_of_masker
    masker.patch_modules(model=model, names_tobe_masked=names_tobe_masked, name_of_masker=conf.name_of_masker)
    return masker


def init_optimizer(model, training_args, num_train_data):
    """One param group per trainable tensor + linear schedule (reference :612-631; max(1, n_gpu) avoids the
    reference's division by zero on a CPU-only host)."""
    params = [{"params": [v], "name": k, "weight_decay": training_args.weight_decay, "param_size": v.size(),
               "nelement": v.nelement(), "lr": training_args.learning_rate}
              for k, v in model.named_parameters() if v.requires_grad]
    optimizer = AdamW(params, lr=training_args.learning_rate, eps=training_args.adam_epsilon)
    world = torch.distributed.get_world_size() if training_args.local_rank != -1 else 1
    per_step = max(1, training_args.n_gpu) * world * training_args.per_gpu_train_batch_size
    num_training_steps = int(int(num_train_data / per_step + 1) * training_args.num_train_epochs)
    scheduler = get_linear_schedule_with_warmup(optimizer, num_warmup_steps=training_args.warmup_steps,
                                                num_training_steps=num_training_steps)
    return optimizer, scheduler


class SyntheticVQADataset(Dataset):
    """VQA-CP-shaped examples as the 8-tuple of dataset_LXM.VQAFeatureDataset.__getitem__ (:282):
    (ids[T], feats[R,2048], pos[R,4], target[A], qid, img_id, bias[A], max_label)."""

    def __init__(self, n, ans_num, seed=49, tokens=20, regions=36, feat_dim=2048, vocab=30522):
        g = torch.Generator().manual_seed(seed)
        self.ids = torch.randint(0, vocab, (n, tokens), generator=g)
        self.feats = torch.randn(n, regions, feat_dim, generator=g)
        self.pos = torch.rand(n, regions, 4, generator=g)
        self.target = (torch.rand(n, ans_num, generator=g) > 0.999).float() * torch.rand(n, ans_num, generator=g)
        self.bias = torch.rand(n, ans_num, generator=g) * 0.01
        self.max_label = self.target.argmax(1)

    def __len__(self):
        return self.ids.shape[0]

    def __getitem__(self, i):
        return (self.ids[i], self.feats[i], self.pos[i], self.target[i], torch.tensor(i), torch.tensor(i),
                self.bias[i], self.max_label[i])


def synthetic_batch(B, ans_num, seed=49, tokens=20, regions=36, feat_dim=2048, vocab=30522):
    """One synthetic VQA-CP-shaped batch as a dict (the recipe of SURVEY.md section 8(d); the order of the draws
    matters -- it is the batch the reference goldens were generated on): ids [B,T] int64, feats [B,R,feat] fp32,
    pos [B,R,4] in [0,1), target [B,A] sparse soft labels, bias [B,A], max_label [B]."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab, (B, tokens), generator=g)
    feats = torch.randn(B, regions, feat_dim, generator=g)
    pos = torch.rand(B, regions, 4, generator=g)
    target = (torch.rand(B, ans_num, generator=g) > 0.999).float() * torch.rand(B, ans_num, generator=g)
    bias = torch.rand(B, ans_num, generator=g) * 0.01
    return {"ids": ids, "feats": feats, "pos": pos, "target": target, "bias": bias, "max_label": target.argmax(1)}


def batch_tuple(batch):
    """The dict above as the 8-tuple the trainers index (dataset_LXM.VQAFeatureDataset.__getitem__, :282)."""
    qid = torch.arange(batch["ids"].shape[0])
    return [batch["ids"], batch["feats"], batch["pos"], batch["target"], qid, qid.clone(), batch["bias"],
            batch["max_label"]]


def build_stage2(ans_num=2274, model_args=None, device=None, seed=49, config_kwargs=None, quiet=True):
    """Model + masker for a synthetic stage-2 run: random-init LXMERT under `seed`, moved to `device`,
    then patched (so the magnitude init runs on the GPU)."""
    from hg_transformers.modeling_lxmert import LxmertConfig, LxmertForMultipleChoice
    model_args = model_args or ModelArguments()
    torch.manual_seed(seed)
    model = LxmertForMultipleChoice(LxmertConfig(ans_num=ans_num, **(config_kwargs or {})))
    if device is not None:
        model = model.to(device)
    hp = HPmodel_modal(Lang=1 - model_args.Lang_comp, Vis=1 - model_args.Vis_comp, Fus=1 - model_args.Fus_comp,
                       P=model_args.zero_rate)
    log = logging.getLogger("crvqa.masker")
    if quiet:
        log.setLevel(logging.WARNING)
    masker = init_masker(model_args, model, log, hp, model_args)
    return model, masker, model_args


def main(argv=None):
    ap = argparse.ArgumentParser(description="synthetic stage-2 mask training (LXMERT)")
    ap.add_argument("--Masker_type", default="lpf", choices=["normal", "lmh", "lpf", "rubi"])
    ap.add_argument("--per_gpu_train_batch_size", type=int, default=32)
    ap.add_argument("--max_steps", type=int, default=4)
    ap.add_argument("--logging_steps", type=int, default=2)
    ap.add_argument("--ans_num", type=int, default=2274)
    ap.add_argument("--num_examples", type=int, default=256)
    ap.add_argument("--output_dir", default="./out_stage2")
    ap.add_argument("--seed", type=int, default=49)
    a = ap.parse_args(argv)
    from hg_transformers.data.data_collator import TrimCollator
    from hg_transformers.data.metrics import vqa_compute_metrics
    from hg_transformers.mask_trainer_Robust_VQA import Trainer
    targs = TrainingArguments(output_dir=a.output_dir, per_gpu_train_batch_size=a.per_gpu_train_batch_size,
                              max_steps=a.max_steps, logging_steps=a.logging_steps, seed=a.seed,
                              Masker_type=a.Masker_type, training_type="Masker", save_steps=0)
    model, masker, margs = build_stage2(a.ans_num, device=targs.device, seed=a.seed)
    data = SyntheticVQADataset(a.num_examples, a.ans_num, seed=a.seed)
    opt = init_optimizer(model, targs, len(data))
    trainer = Trainer(model=model, args=targs, model_args=margs, data_collator=TrimCollator(), train_dataset=data,
                      eval_dataset=None, compute_metrics=vqa_compute_metrics, optimizers=opt, masker=masker)
    out = trainer.train()
    print(out[0])


if __name__ == "__main__":
    main()
