"""Shared body of the three stage-2 trainers (reference: hg_transformers/mask_trainer_VQA.py,
mask_trainer_Robust_VQA.py, mask_trainer_visualBERT_VQA.py -- 1200-line near copies of an HF-2.10
Trainer fork).  The public surface (constructor, train / evaluate / predict / _training_step /
reset_threshold / save_model_mask / ...) is kept; what runs underneath is different:

  * one process per GPU; data parallel = asynchronous bucketed all-reduce of score + classifier
    gradients only (hg_transformers._engine.GradSync) instead of nn.DataParallel's per-step broadcast
    of 207 M scores + 207 M frozen weights (reference mask_trainer_VQA.py:533-543);
  * every masked Linear is a fused tcgen05 GEMM; score gradients accumulate in a flat arena;
  * loss (+ dlogits + batch score) is one fused kernel; loss / score stay on the device and are read
    back only at logging steps (the reference synchronises three times per step, :855-886);
  * global-norm clip + AdamW is one streaming pass; reset_threshold is one batched exact radix select
    (reference: 168 single-CTA torch.kthvalue calls).
"""
import json
import logging
import os
import random
import re
import shutil
from pathlib import Path
from typing import Callable, Dict, List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn
from torch.utils.data.dataloader import DataLoader
from torch.utils.data.dataset import Dataset
from torch.utils.data.distributed import DistributedSampler
from torch.utils.data.sampler import Sampler

from crvqa import ops

from ._engine import execution_order, GradSync, GraphedStep, InputPrefetcher, ScoreArena, masked_modules_of
from .data.data_collator import DataCollator, DefaultDataCollator, TrimCollator  # noqa: F401
from .optimization import AdamW, get_constant_schedule, get_linear_schedule_with_warmup  # noqa: F401
from .trainer_utils import PREFIX_CHECKPOINT_DIR, EvalPrediction, PredictionOutput, TrainOutput
from .training_args import TrainingArguments, is_tpu_available  # noqa: F401
from .vqa_debias_loss_functions import LearnedMixin

logger = logging.getLogger(__name__)

try:
    from torch.utils.tensorboard import SummaryWriter
    _has_tensorboard = True
except Exception:  # pragma: no cover
    _has_tensorboard = False


def is_tensorboard_available():
    return _has_tensorboard


def is_wandb_available():
    return False


def is_apex_available():
    return False


def set_seed(seed: int):
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


class SequentialDistributedSampler(Sampler):
    """Sequential shard of a dataset per rank, padded so every rank gets the same count
    (reference mask_trainer_VQA.py:137-170)."""

    def __init__(self, dataset, num_replicas=None, rank=None):
        if num_replicas is None:
            num_replicas = torch.distributed.get_world_size()
        if rank is None:
            rank = torch.distributed.get_rank()
        self.dataset, self.num_replicas, self.rank = dataset, num_replicas, rank
        self.num_samples = (len(dataset) + num_replicas - 1) // num_replicas
        self.total_size = self.num_samples * num_replicas

    def __iter__(self):
        idx = list(range(len(self.dataset)))
        idx += idx[: self.total_size - len(idx)]
        return iter(idx[self.rank * self.num_samples: (self.rank + 1) * self.num_samples])

    def __len__(self):
        return self.num_samples


def LPF_loss(logits, bias, max_label, device, gamma):
    """mean_b (1 - q[b,y])^gamma * -log p[b,y]  (reference mask_trainer_VQA.py:111-129) -- fused kernel."""
    return ops.vqa_loss_lpf(logits, bias, max_label, gamma)[0]


def RUBI_loss(logits, bias, max_label):
    """CE(logits * sigmoid(bias), y) (reference :131-135): fused kernel on CUDA tensors, torch graph on CPU ones."""
    if logits.is_cuda:
        return ops.vqa_loss_rubi(logits, bias, max_label)[0]
    return F.cross_entropy(logits * torch.sigmoid(bias), max_label)


class CosineLoss(nn.Module):
    def forward(self, a, b):
        return (1 - F.cosine_similarity(a, b, dim=-1)).mean()


class TrainerCore:
    """See the module docstring.  Subclasses choose `threshold_mode` ('global' | 'modal') and
    `forward_style` ('lxmert' | 'visualbert')."""

    threshold_mode = "global"
    forward_style = "lxmert"

    def __init__(self, model, args: TrainingArguments, model_args, data_collator: Optional[DataCollator] = None,
                 train_dataset: Optional[Dataset] = None, eval_dataset: Optional[Dataset] = None,
                 compute_metrics: Optional[Callable[[EvalPrediction], Dict]] = None, prediction_loss_only=False,
                 tb_writer=None, optimizers=None, masker=None, head_mask_weight=None, ffn_mask_weight=None,
                 threshold_fn_head=None, threshold_fn_ffn=None, teacher_model=None):
        self.model = model.to(args.device)
        self.teacher_model = teacher_model.to(args.device) if teacher_model is not None else None
        if self.teacher_model is not None:
            self.kd_loss_fn = CosineLoss()
        self.masker = masker
        self.lpf_loss = LPF_loss
        self.rubi_loss = RUBI_loss
        hidden = getattr(getattr(model, "config", None), "hidden_size", 768)
        self.debias_loss_fn = LearnedMixin(0.36, hidden_size=hidden).to(args.device)
        self.head_mask_weight, self.ffn_mask_weight = head_mask_weight, ffn_mask_weight
        self.threshold_fn_head, self.threshold_fn_ffn = threshold_fn_head, threshold_fn_ffn
        self.args, self.model_args = args, model_args
        self.data_collator = data_collator if data_collator is not None else DefaultDataCollator()
        self.train_dataset, self.eval_dataset = train_dataset, eval_dataset
        self.compute_metrics = compute_metrics
        self.prediction_loss_only = prediction_loss_only
        self.optimizers = optimizers
        self.tb_writer = tb_writer
        self.global_step: Optional[int] = None
        self.epoch: Optional[float] = None
        self.tr_rep_loss = 0.0
        self.logging_rep_loss = 0.0
        self.arena: Optional[ScoreArena] = None
        self.grad_sync: Optional[GradSync] = None

        assert self.args.training_type in ["Masker", "FTonly", "FTlmh", "FTlpf", "FTrubi", "FT_trainedMask",
                                           "FT_randMask"]
        self.debiasing = self.args.Masker_type if self.args.training_type == "Masker" else self.args.FT_type
        assert self.debiasing in ["normal", "lmh", "lpf", "rubi"]
        if getattr(model_args, "structured", False):
            raise NotImplementedError("structured (head / ffn) masking is not part of the stage-2 VQA scripts")
        if self.args.use_kd:
            raise NotImplementedError("use_kd is off in every stage-2 script and is not provided")
        if self.tb_writer is None and is_tensorboard_available() and self.is_world_master() and self.args.logging_dir:
            self.tb_writer = SummaryWriter(log_dir=self.args.logging_dir)
        set_seed(self.args.seed)
        # The dense (unmasked, trainable) answer head goes through torch matmul: let it use TF32 tensor cores rather
        # than SIMT fp32 (0.3 ms of a 18 ms step); every masked layer already multiplies in bf16 with fp32
        # accumulation, so this is the most precise GEMM of the step either way.  CRVQA_TF32=0 keeps strict fp32.
        if os.environ.get("CRVQA_TF32", "1") != "0":
            torch.backends.cuda.matmul.allow_tf32 = True
        if self.is_world_master():
            os.makedirs(self.args.output_dir, exist_ok=True)

    # ------------------------------------------------------------------ data
    def _loader(self, dataset, batch_size, shuffle):
        sampler = None
        if self.args.local_rank != -1 and shuffle:
            # one process per GPU: every rank trains on its own shard (the reference leaves its samplers
            # commented out, mask_trainer_VQA.py:317-324, which under DDP would repeat the full set on every rank)
            sampler = DistributedSampler(dataset, shuffle=True, seed=self.args.seed)
        return DataLoader(dataset, batch_size=batch_size, sampler=sampler, shuffle=shuffle and sampler is None,
                          num_workers=self.args.dataloader_num_workers, collate_fn=self.data_collator.collate_batch,
                          pin_memory=torch.cuda.is_available())

    def get_train_dataloader(self) -> DataLoader:
        if self.train_dataset is None:
            raise ValueError("Trainer: training requires a train_dataset.")
        return self._loader(self.train_dataset, self.args.train_batch_size, True)

    def get_eval_dataloader(self, eval_dataset: Optional[Dataset] = None) -> DataLoader:
        if eval_dataset is None and self.eval_dataset is None:
            raise ValueError("Trainer: evaluation requires an eval_dataset.")
        return self._loader(eval_dataset if eval_dataset is not None else self.eval_dataset,
                            self.args.eval_batch_size, False)

    def get_test_dataloader(self, test_dataset: Dataset) -> DataLoader:
        return self._loader(test_dataset, self.args.eval_batch_size, False)

    def num_examples(self, dataloader: DataLoader) -> int:
        return len(dataloader.dataset)

    # ------------------------------------------------------------------ optimiser
    def get_optimizers(self, num_training_steps: int):
        if self.optimizers is not None:
            return self.optimizers
        no_decay = ["bias", "LayerNorm.weight"]
        named = [(n, p) for n, p in self.model.named_parameters() if p.requires_grad]
        groups = [
            {"params": [p for n, p in named if not any(nd in n for nd in no_decay)], "weight_decay": self.args.weight_decay},
            {"params": [p for n, p in named if any(nd in n for nd in no_decay)], "weight_decay": 0.0},
        ]
        optimizer = AdamW([g for g in groups if g["params"]], lr=self.args.learning_rate, eps=self.args.adam_epsilon)
        scheduler = get_linear_schedule_with_warmup(optimizer, num_warmup_steps=self.args.warmup_steps,
                                                    num_training_steps=num_training_steps)
        self.optimizers = optimizer, scheduler
        return optimizer, scheduler

    # ------------------------------------------------------------------ thresholds / masks
    def _sparsity_of(self, name, init_sparsity):
        if self.threshold_mode == "modal":
            modal = self.masker.name_in_module[name]
            return self.masker.hpmodel.zerorate_dict[modal]
        return init_sparsity

    def reset_threshold(self, model, init_sparsity):
        """Per module: threshold = k-th smallest score, k = int(numel * rate) (>= 1); returns the mean
        threshold as a Python float (reference mask_trainer_Robust_VQA.py:467-482, mask_trainer_VQA.py:470-477).
        All modules are selected by ONE batched launch sequence; module.threshold becomes a 0-dim device
        tensor exactly as in the reference."""
        mods = masked_modules_of(model)
        if not mods:
            return float("nan")
        if getattr(self, "grad_sync", None) is not None:
            self.grad_sync.sync_scores()          # sharded optimiser: every rank selects over every current score
        if self.threshold_mode == "union":
            return self._reset_threshold_union(mods, init_sparsity)
        ks = []
        for name, module in mods:
            k = int(module.weight.nelement() * self._sparsity_of(name, init_sparsity))
            ks.append(1 if k == 0 else k)
        tensors = [m.weight_mask.data for _, m in mods]
        sig = tuple(t.data_ptr() for t in tensors)
        plan = getattr(self, "_kth_plan", None)
        if plan is None or plan[0] != sig:       # scores live in the arena: same storage every call
            plan = (sig, ops.KthPlan(tensors))
            self._kth_plan = plan
        thr = plan[1](ks)
        arena = getattr(self, "arena", None)
        modules = [m for _, m in mods]
        if arena is not None and len(modules) == len(arena.modules) and set(map(id, modules)) == set(map(id, arena.modules)):
            # one device vector in ARENA order (the arena follows execution order, not named_modules order);
            # module.threshold = 0-dim views of it
            pos = {id(m): i for i, m in enumerate(modules)}
            perm = [pos[id(m)] for m in arena.modules]
            if perm != list(range(len(perm))):
                thr_arena = thr[torch.tensor(perm, device=thr.device)]
            else:
                thr_arena = thr
            arena.set_thresholds(thr_arena)
            arena.refresh_masked()
        else:
            for i, (_, module) in enumerate(mods):
                module.threshold = thr[i]
        return float(thr.mean())

    def _reset_threshold_union(self, mods, tgt_sparsity):
        """global_mask_trainer_VQA.py:421-443: one threshold = k-th smallest of ALL scores together."""
        if not getattr(self.model_args, "global_prune", False):
            raise AssertionError("this function is designed for GLOBAL_MASKER, plz check it again whether run the "
                                 "wrong py-files")
        from masking._core import global_kth_value
        tensors = [m.weight_mask.data for _, m in mods]
        k = int(sum(t.numel() for t in tensors) * tgt_sparsity)
        thr = global_kth_value(tensors, k)                      # 1-element device tensor
        arena = getattr(self, "arena", None)
        modules = [m for _, m in mods]
        if arena is not None and len(modules) == len(arena.modules) and set(map(id, modules)) == set(map(id, arena.modules)):
            arena.set_thresholds(thr.expand(len(modules)).contiguous())
            arena.refresh_masked()
        else:
            shared = thr[0]
            for m in modules:
                m.threshold = shared
        # the reference returns the fp32 mean of the per-module list (all entries identical)
        return float(torch.tensor([float(thr[0])] * len(modules)).mean())

    def binarizer_fn1(self, inputs, threshold):
        return ops.binarize(inputs, threshold)

    def binarizer_fn2(self, inputs):
        outputs = inputs.clone()
        inputs.data.clamp_(-1, 1)
        outputs.data = (torch.sign(outputs.data) + 1) / 2
        return outputs

    def binarizer_fn3(self, inputs):
        return torch.bernoulli(torch.sigmoid(inputs))

    def save_model_mask(self, output_dir: Optional[str] = None):
        """mask.pt = {module_name + '.weight': BoolTensor(cpu)}; returns the overall zero rate in percent
        (reference mask_trainer_VQA.py:930-949; per-modality logging of mask_trainer_Robust_VQA.py:943-991)."""
        if getattr(self, "grad_sync", None) is not None:
            self.grad_sync.make_consistent()
        mask_dict = {}
        zeros = {"all": 0, "Lang": 0, "Vis": 0, "Fus": 0, "P": 0}
        elems = dict(zeros)
        logger.info("Collecting mask...")
        for name, module in self.model.named_modules():
            if not hasattr(module, "threshold"):
                continue
            mask, kept = ops.binarize(module.weight_mask, module.threshold, want_count=True, as_bool=True)
            mask_dict[name + ".weight"] = mask.cpu()
            n, z = mask.numel(), mask.numel() - int(kept)
            zeros["all"] += z
            elems["all"] += n
            if self.threshold_mode == "modal":
                modal = self.masker.name_in_module[name[7:] if name.startswith("module.") else name]
                assert modal in ("Lang", "Vis", "Fus", "P")
                zeros[modal] += z
                elems[modal] += n
        output_dir = output_dir if output_dir is not None else self.args.output_dir
        os.makedirs(output_dir, exist_ok=True)
        zero_rate = torch.tensor(100.0 * zeros["all"] / max(1, elems["all"]))
        logger.info("Saving model mask to %s", output_dir)
        logger.info("Zero rate = %.2f", zero_rate)
        if self.threshold_mode == "modal":
            for modal in ("Lang", "Vis", "Fus", "P"):
                if elems[modal]:
                    logger.info("Zero rate %s = %.2f", modal.upper(), 100.0 * zeros[modal] / elems[modal])
        torch.save(mask_dict, os.path.join(output_dir, "mask.pt"))
        return zero_rate

    # ------------------------------------------------------------------ engine set-up
    def _setup_engine(self, optimizer):
        """Move the scores into a flat arena and hook up the data-parallel gradient exchange."""
        if self.arena is not None:
            return
        if self.masker is None and hasattr(optimizer, "attach_weight_arena"):
            return self._setup_finetune_engine(optimizer)
        mods = [(n, m) for n, m in masked_modules_of(self.model)
                if m.weight_mask.requires_grad and m.weight_mask.is_cuda and getattr(m, "unstructured_masked", False)]
        if not mods:
            return
        self.arena = ScoreArena(execution_order(mods))
        if os.environ.get("CRVQA_MASK_MODE", "cached") == "cached":
            self.arena.enable_mask_cache()
        if hasattr(optimizer, "attach_arena"):
            optimizer.attach_arena(self.arena)
        self.grad_sync = GradSync(self.arena)
        self.grad_sync.defer = self.args.gradient_accumulation_steps > 1
        self.grad_sync.attach_loose(self._loose_params())

    def _setup_finetune_engine(self, optimizer):
        """Stage 3 (FT_trainedMask / FT_randMask, run_vqa_stage3.py): every trainable tensor moves into a WeightArena,
        the optimiser (run_vqa_stage3.ArenaAdam) steps it with one launch.  CRVQA_FT_ENGINE=0 keeps per-tensor Adam."""
        if os.environ.get("CRVQA_FT_ENGINE", "1") == "0" or self.args.device.type != "cuda":
            return
        from ._engine_ft import WeightArena
        owned = {id(p) for g in optimizer.param_groups for p in g["params"]}
        if owned != {id(p) for p in self.model.parameters() if p.requires_grad}:
            return          # a hand-made optimiser over a subset: leave it alone
        try:
            self.arena = WeightArena(self.model)
        except ValueError:
            return
        optimizer.attach_weight_arena(self.arena)
        self.grad_sync = GradSync(self.arena, mode="allreduce")
        self.grad_sync.defer = self.args.gradient_accumulation_steps > 1

    def _prefetched(self, dataloader):
        """Yield batches whose host -> device copy was started one step ahead on a side stream."""
        if not torch.cuda.is_available() or self.args.device.type != "cuda":
            yield from dataloader
            return
        pre = InputPrefetcher(self.args.device)
        it = iter(dataloader)
        try:
            handle = pre.stage(next(it))
        except StopIteration:
            return
        while handle is not None:
            batch = pre.take(handle)
            cur = handle
            try:
                handle = pre.stage(next(it))
            except StopIteration:
                handle = None
            yield batch
            pre.release(cur)

    def _make_graphed_step(self, model, optimizer, scheduler):
        """CUDA-graph replay of the whole step when it is safe: arena engine, our AdamW, no accumulation.
        CRVQA_CUDA_GRAPH=0 keeps the eager loop; multi-rank runs capture the NCCL all-reduces too."""
        if os.environ.get("CRVQA_CUDA_GRAPH", "1") == "0" or self.arena is None or not torch.cuda.is_available():
            return None
        if self.args.gradient_accumulation_steps != 1 or not hasattr(optimizer, "use_device_hyper"):
            return None
        return GraphedStep(self, model, optimizer, scheduler)

    def _loose_params(self):
        return [p for p in self.model.parameters() if p.requires_grad and not (self.arena and self.arena.owns(p))]

    def _zero_grad(self, optimizer):
        if self.arena is not None:
            self.arena.begin_step()
            for p in self._loose_params():
                p.grad = None
        else:
            optimizer.zero_grad()

    def _clip_and_optimizer_step(self, model, optimizer):
        """clip_grad_norm_(max_grad_norm) + optimizer.step() (reference :646-653): device work only."""
        max_norm = self.args.max_grad_norm
        if self.grad_sync is not None and self.grad_sync.sharded and hasattr(optimizer, "set_clip"):
            optimizer.set_clip(self.grad_sync.global_sumsq([p.grad for p in self._loose_params()]), max_norm)
        elif self.arena is not None and hasattr(optimizer, "set_clip"):
            sumsq = torch.zeros((), dtype=torch.float32, device=self.args.device)
            self.arena.grad_sumsq_into(sumsq)
            for p in self._loose_params():
                if p.grad is not None:
                    ops.sumsq_into(p.grad.contiguous(), sumsq)
            optimizer.set_clip(sumsq, max_norm)
        else:
            torch.nn.utils.clip_grad_norm_(model.parameters(), max_norm)
        optimizer.step()

    def _clip_and_step(self, model, optimizer, scheduler):
        self._clip_and_optimizer_step(model, optimizer)
        scheduler.step()

    def _device_step(self, model, inputs, optimizer):
        """Everything of one optimisation step that runs on the GPU (gradient_accumulation_steps == 1):
        forward, loss, backward, gradient exchange, clip + AdamW (+ mask-cache refresh), zero_grad.
        This is the unit hg_transformers._engine.GraphedStep captures as one CUDA graph."""
        if torch.cuda.is_available():
            from crvqa.fused import RngState
            RngState.get(self.args.device).advance()   # fresh dropout masks for the fused kernels, graph-safe
        loss, score = self._training_step(model, inputs, optimizer)
        if self.grad_sync is not None:
            self.grad_sync.finish([p.grad for p in self._loose_params()] + self.arena.loose_grads())
        self._clip_and_optimizer_step(model, optimizer)
        self._zero_grad(optimizer)
        return loss, score

    # ------------------------------------------------------------------ the loop
    def train(self, model_path: Optional[str] = None):
        train_dataloader = self.get_train_dataloader()
        accum = self.args.gradient_accumulation_steps
        if self.args.max_steps > 0:
            t_total = self.args.max_steps
            num_train_epochs = self.args.max_steps // max(1, len(train_dataloader) // accum) + 1
        else:
            t_total = int(len(train_dataloader) // accum * self.args.num_train_epochs)
            num_train_epochs = self.args.num_train_epochs
        optimizer, scheduler = self.get_optimizers(num_training_steps=t_total)
        if (model_path is not None and os.path.isfile(os.path.join(model_path, "optimizer.pt"))
                and os.path.isfile(os.path.join(model_path, "scheduler.pt"))):
            optimizer.load_state_dict(torch.load(os.path.join(model_path, "optimizer.pt"), map_location=self.args.device))
            scheduler.load_state_dict(torch.load(os.path.join(model_path, "scheduler.pt")))
        model = self.model
        if self.args.fp16:
            raise ImportError("fp16 through apex is not provided; the masked GEMMs already run bf16 MMAs")
        self._setup_engine(optimizer)

        world = torch.distributed.get_world_size() if self.args.local_rank != -1 else 1
        logger.info("***** Running training *****")
        logger.info("  Num examples = %d", self.num_examples(train_dataloader))
        logger.info("  Num Epochs = %d", num_train_epochs)
        logger.info("  Instantaneous batch size per device = %d", self.args.per_gpu_train_batch_size)
        logger.info("  Total train batch size (w. parallel, distributed & accumulation) = %d",
                    self.args.train_batch_size * accum * world)
        logger.info("  Total optimization steps = %d", t_total)

        self.global_step, self.epoch = 0, 0
        dev = self.args.device
        tr_loss = torch.zeros((), device=dev)
        tr_score = torch.zeros((), device=dev)
        logging_loss = logging_score = 0.0
        self.tr_rep_loss = self.logging_rep_loss = 0.0
        best_eval_loss, best_score, results_at_best_score = 100.0, 0.0, None
        self._zero_grad(optimizer)
        if self.eval_dataset is not None:
            # the reference always evaluates before step 0 and asks the user to check it (:592-594)
            result_start, _ = self.evaluate()
            print(result_start)
            print("\n\n\n!!!!PLZ check the results of the models loaded checkpoint+mask+clf!\n\n\n\n")

        graphed = self._make_graphed_step(model, optimizer, scheduler)
        stop = False
        for epoch in range(int(np.ceil(num_train_epochs))):
            if isinstance(train_dataloader.sampler, DistributedSampler):
                train_dataloader.sampler.set_epoch(epoch)
            n_batches = len(train_dataloader)
            for step, inputs in enumerate(self._prefetched(train_dataloader)):
                if graphed is not None:
                    loss_batch, score_batch = graphed.step(inputs)
                else:
                    loss_batch, score_batch = self._training_step(model, inputs, optimizer)
                tr_loss += loss_batch
                tr_score += score_batch
                if (step + 1) % accum == 0 or (n_batches <= accum and (step + 1) == n_batches):
                    if graphed is None:
                        if self.grad_sync is not None:
                            self.grad_sync.finish([p.grad for p in self._loose_params()] + self.arena.loose_grads())
                        self._clip_and_step(model, optimizer, scheduler)
                        self._zero_grad(optimizer)
                    self.global_step += 1
                    self.epoch = epoch + (step + 1) / n_batches

                    if (self.args.logging_steps > 0 and self.global_step % self.args.logging_steps == 0) or (
                            self.global_step == 1 and self.args.logging_first_step):
                        cur_loss, cur_score = float(tr_loss), float(tr_score)  # the only host sync of the loop
                        logs: Dict[str, float] = {
                            "loss": (cur_loss - logging_loss) / self.args.logging_steps,
                            "rep_loss": (self.tr_rep_loss - self.logging_rep_loss) / self.args.logging_steps,
                            "score": 100 * (cur_score - logging_score) / (self.args.logging_steps * self.args.train_batch_size),
                            "learning_rate": scheduler.get_last_lr()[0],
                        }
                        logging_loss, logging_score = cur_loss, cur_score
                        self.logging_rep_loss = self.tr_rep_loss
                        if self.masker is not None:
                            self.reset_threshold(model, self.masker.masker_scheduler.init_sparsity)
                        self._log(logs)

                    if self.args.save_steps > 0 and self.global_step % self.args.save_steps == 0 \
                            and self.args.evaluate_during_training and self.eval_dataset is not None:
                        results, eval_output = self.evaluate()
                        value = results.get("eval_acc")
                        if value is not None:
                            if best_score < value:
                                best_score, results_at_best_score = value, results
                                self._save_best(model, eval_output)
                            elif value == 0 and best_score == 0:
                                results_at_best_score = results
                            self._log({"best_score": best_score})
                if 0 < self.args.max_steps <= self.global_step:
                    stop = True
                    break
            if stop:
                break
        if self.tb_writer:
            self.tb_writer.close()
        if self.grad_sync is not None:
            self.grad_sync.make_consistent()      # sharded optimiser: leave every rank with the full final state
        logger.info("\n\nTraining completed.\n\n")
        return (TrainOutput(self.global_step, float(tr_loss) / max(1, self.global_step)), best_eval_loss, best_score,
                results_at_best_score)

    def _save_best(self, model, eval_output):
        """What the reference does on a new best eval accuracy (mask_trainer_VQA.py:697-742).  The threshold refresh
        changes the masks every rank trains with, so ALL ranks run it; only the file writes are the master's."""
        if self.args.training_type == "Masker" and self.masker:
            self.reset_threshold(model, self.masker.masker_scheduler.init_sparsity)
        if not self.is_world_master():
            return
        try:
            loader = self.get_eval_dataloader(self.eval_dataset)
            with open(os.path.join(self.args.output_dir, "test.json"), "w") as f:
                json.dump(self.make_json(eval_output[0], eval_output[3], loader), f)
        except (AttributeError, TypeError):
            pass  # synthetic datasets carry no label2ans table
        kind = self.args.training_type
        if kind == "Masker":
            if not self.masker:
                raise AssertionError("When you are training the masker, please pass the initialed masker into trainer()")
            self.save_model_mask(self.args.output_dir)
            head = getattr(model, "classifier", None) or getattr(model, "cls", None)
            if head is None:
                raise AssertionError("cannot find classifier in model!")
            torch.save(head, os.path.join(self.args.output_dir, "classifier4masker.bin"))
        else:
            suffix = {"FTonly": "_FTonly.bin", "FTlmh": "_FTlmh_only.bin", "FTlpf": "_FTlpf_only.bin",
                      "FTrubi": "_FTrubi_only.bin", "FT_trainedMask": "_FT_trainedMask.bin",
                      "FT_randMask": "FT_randMask.bin"}[kind]
            torch.save(model, os.path.join(self.args.output_dir, str(self.args.label4save) + suffix))

    def _log(self, logs: Dict[str, float], iterator=None) -> None:
        if self.epoch is not None:
            logs["epoch"] = self.epoch
        if self.tb_writer:
            for k, v in logs.items():
                self.tb_writer.add_scalar(k, v, self.global_step)
        output = json.dumps({**{k: float(v) for k, v in logs.items()}, **{"step": self.global_step}})
        if self.is_world_master():
            print(output)

    # ------------------------------------------------------------------ one step
    def _forward(self, model, inputs):
        dev = self.args.device
        nb = torch.cuda.is_available()
        if self.forward_style == "visualbert":
            return model(input_ids=inputs[0].to(dev, non_blocking=nb), visual_embeds=inputs[1].to(dev, non_blocking=nb),
                         labels=inputs[3].to(dev, non_blocking=nb))
        return model(inputs[0].to(dev, non_blocking=nb), inputs[1].to(dev, non_blocking=nb),
                     inputs[2].to(dev, non_blocking=nb), labels=inputs[3].to(dev, non_blocking=nb))

    def _loss_and_score(self, outputs, inputs):
        """Loss dispatch on Masker_type / FT_type (reference _training_step :797-842) + batch VQA score."""
        dev = self.args.device
        logits, pool_out = outputs[1], outputs[2]
        labels = inputs[3].to(dev, non_blocking=True)
        kind = self.debiasing
        if kind == "normal":
            if self.forward_style == "visualbert":
                loss = outputs[0]  # VisualBERT's own soft-label cross entropy
                score = labels.gather(1, logits.detach().max(1)[1].view(-1, 1)).sum()
            else:
                loss, score = ops.vqa_loss_bce(logits, labels)
        elif kind == "lmh":
            loss = self.debias_loss_fn(pool_out, logits, inputs[6].to(dev, non_blocking=True), labels, dev)
            score = self.debias_loss_fn.last_score
        elif kind == "lpf":
            loss, score = ops.vqa_loss_lpf(logits, inputs[6].to(dev, non_blocking=True),
                                           inputs[7].to(dev, non_blocking=True), self.args.gamma, labels)
        else:  # rubi
            loss, score = ops.vqa_loss_rubi(logits, inputs[6].to(dev, non_blocking=True),
                                            inputs[7].to(dev, non_blocking=True), labels)
        return loss, score

    def _training_step(self, model: nn.Module, inputs, optimizer) -> Tuple[torch.Tensor, torch.Tensor]:
        """forward -> loss -> backward.  Returns (loss, batch score) as 0-dim DEVICE tensors: unlike the
        reference (:855-886) nothing here synchronises with the host."""
        model.train()
        if self.grad_sync is not None and self.grad_sync._pending is None:
            self.grad_sync.begin_step()
        outputs = self._forward(model, inputs)
        loss, score = self._loss_and_score(outputs, inputs)
        if self.args.gradient_accumulation_steps > 1:
            loss = loss / self.args.gradient_accumulation_steps
        loss.backward()
        if hasattr(optimizer, "accumulate_grad"):
            optimizer.accumulate_grad()
        return loss.detach(), score.detach()

    # ------------------------------------------------------------------ evaluation
    def evaluate(self, eval_dataset: Optional[Dataset] = None, prediction_loss_only: Optional[bool] = None):
        eval_dataloader = self.get_eval_dataloader(eval_dataset)
        output = self._prediction_loop(eval_dataloader, description="Evaluation")
        self._log(output.metrics)
        return output.metrics, output

    def predict(self, test_dataset: Dataset) -> PredictionOutput:
        return self._prediction_loop(self.get_test_dataloader(test_dataset), description="Prediction")

    def _prediction_loop(self, dataloader: DataLoader, description: str,
                         prediction_loss_only: Optional[bool] = None) -> PredictionOutput:
        prediction_loss_only = prediction_loss_only if prediction_loss_only is not None else self.prediction_loss_only
        if getattr(self, "grad_sync", None) is not None:
            self.grad_sync.make_consistent()
        model = self.model
        logger.info("***** Running %s *****", description)
        logger.info("  Num examples = %d", self.num_examples(dataloader))
        logger.info("  Batch size = %d", dataloader.batch_size)
        eval_losses: List[torch.Tensor] = []
        preds, label_ids, q_ids = [], [], []
        model.eval()
        dev = self.args.device
        for inputs in dataloader:
            inputs = [v.to(dev) if torch.is_tensor(v) else v for v in inputs]
            with torch.no_grad():
                outputs = self._forward(model, inputs)
                loss, _ = self._loss_and_score(outputs, inputs)
                eval_losses.append(loss.detach().mean())
                logits = outputs[1]
            if not prediction_loss_only:
                preds.append(logits.detach())
                if inputs[3] is not None:
                    label_ids.append(inputs[3].detach())
                if len(inputs) > 4 and inputs[4] is not None and torch.is_tensor(inputs[4]):
                    q_ids.append(inputs[4].detach())
        preds = torch.cat(preds) if preds else None
        label_ids = torch.cat(label_ids) if label_ids else None
        q_ids = torch.cat(q_ids) if q_ids else None
        if self.args.local_rank != -1:
            n = self.num_examples(dataloader)
            if preds is not None:
                preds = self.distributed_concat(preds, num_total_examples=n)
            if label_ids is not None:
                label_ids = self.distributed_concat(label_ids, num_total_examples=n)
            if q_ids is not None:
                q_ids = self.distributed_concat(q_ids, num_total_examples=n)
        if label_ids is not None:
            label_ids = label_ids.cpu()
        if q_ids is not None:
            q_ids = q_ids.cpu()
        metrics = {}
        if self.compute_metrics is not None and preds is not None and label_ids is not None:
            metrics = dict(self.compute_metrics(EvalPrediction(predictions=preds, label_ids=label_ids)))
            metrics["acc"] = float(100 * metrics["acc"] / len(preds))
        if eval_losses:
            metrics["eval_loss"] = float(torch.stack(eval_losses).mean())
        for key in list(metrics.keys()):
            if not key.startswith("eval_"):
                metrics[f"eval_{key}"] = metrics.pop(key)
        return PredictionOutput(predictions=preds, label_ids=label_ids, metrics=metrics, q_ids=q_ids)

    def distributed_concat(self, tensor: torch.Tensor, num_total_examples: int) -> torch.Tensor:
        assert self.args.local_rank != -1
        out = [torch.empty_like(tensor) for _ in range(torch.distributed.get_world_size())]
        torch.distributed.all_gather(out, tensor.contiguous())
        return torch.cat(out, dim=0)[:num_total_examples]

    # ------------------------------------------------------------------ small helpers kept from the reference
    def get_answer(self, p, dataloader):
        _m, idx = p.max(0)
        return dataloader.dataset.label2ans[idx.item()]

    def make_json(self, logits, qIds, dataloader):
        assert logits.size(0) == len(qIds)
        return [{"question_id": qIds[i].item(), "answer": self.get_answer(logits[i], dataloader)}
                for i in range(logits.size(0))]

    def freeze_model(self, model):
        for _, param in model.named_parameters():
            param.requires_grad = False
        return model

    def is_local_master(self) -> bool:
        return self.args.local_rank in [-1, 0]

    def is_world_master(self) -> bool:
        return self.args.local_rank == -1 or torch.distributed.get_rank() == 0

    def save_model(self, output_dir: Optional[str] = None):
        if self.is_world_master():
            self._save(output_dir)

    def _save(self, output_dir: Optional[str] = None):
        output_dir = output_dir if output_dir is not None else self.args.output_dir
        os.makedirs(output_dir, exist_ok=True)
        logger.info("Saving model checkpoint to %s", output_dir)
        torch.save(self.model.state_dict(), os.path.join(output_dir, "pytorch_model.bin"))

    def _sorted_checkpoints(self, checkpoint_prefix=PREFIX_CHECKPOINT_DIR, use_mtime=False) -> List[str]:
        found = []
        for path in (str(x) for x in Path(self.args.output_dir).glob(f"{checkpoint_prefix}-*")):
            if use_mtime:
                found.append((os.path.getmtime(path), path))
            else:
                m = re.match(f".*{checkpoint_prefix}-([0-9]+)", path)
                if m and m.groups():
                    found.append((int(m.groups()[0]), path))
        return [p for _, p in sorted(found)]

    def _rotate_checkpoints(self, use_mtime=False) -> None:
        limit = self.args.save_total_limit
        if limit is None or limit <= 0:
            return
        ckpts = self._sorted_checkpoints(use_mtime=use_mtime)
        for ckpt in ckpts[: max(0, len(ckpts) - limit)]:
            logger.info("Deleting older checkpoint [%s] due to args.save_total_limit", ckpt)
            shutil.rmtree(ckpt)
