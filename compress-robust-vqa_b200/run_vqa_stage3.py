"""Stage-3 driver pieces of the reference's ``run_vqa_stage3.py`` that the frozen-mask fine-tune needs
(SURVEY.md section 8(f) rank 2): the same function names, arguments and results, with the pruned modules of
``masking/pruned.py`` underneath instead of ``torch.nn.utils.prune`` hooks.

    pruning_model_with_mask(model, mask_dict, model_type)    run_vqa_stage3.py:227-297   (FT_trainedMask)
    mag_pruning(model, px)                                   run_vqa_stage3.py:205-225   (FT_randMask)
    see_weight_rate(model, model_type)                       run_vqa_stage3.py:75-178
    init_optimizer(model, training_args, num_train_data)     run_vqa_stage3.py:577-598

The trainer is the stage-2 one (``hg_transformers.mask_trainer_VQA.Trainer``, as the reference imports at :45)
with ``training_type`` in {FT_trainedMask, FT_randMask} and no masker; dataset / argument parsing of the
reference driver is out of scope.
"""
import torch
from torch.optim._functional import adam as _adam_functional

from crvqa import ops
from hg_transformers.optimization import get_linear_schedule_with_warmup
from masking.pruned import PrunedEmbedding, PrunedLinear, custom_from_mask, l1_unstructured_mask  # noqa: F401 (re-exported)

_ATT = ("attention.self.query", "attention.self.key", "attention.self.value", "attention.output.dense",
        "intermediate.dense", "output.dense")
_XSUB = ("visual_attention.att.query", "visual_attention.att.key", "visual_attention.att.value",
         "visual_attention.output.dense", "lang_self_att.self.query", "lang_self_att.self.key",
         "lang_self_att.self.value", "lang_self_att.output.dense", "visn_self_att.self.query",
         "visn_self_att.self.key", "visn_self_att.self.value", "visn_self_att.output.dense",
         "lang_inter.dense", "lang_output.dense", "visn_inter.dense", "visn_output.dense")


def trained_mask_module_names():
    """The modules pruning_model_with_mask reparametrises, in the reference's order (:231-294)."""
    names = [f"encoder.layer.{i}.{s}" for i in range(9) for s in _ATT]
    names += [f"encoder.r_layers.{i}.{s}" for i in range(5) for s in _ATT]
    names += [f"encoder.x_layers.{i}.{s}" for i in range(5) for s in _XSUB]
    return names + ["pooler.dense", "embeddings.word_embeddings", "encoder.visn_fc.visn_fc", "encoder.visn_fc.box_fc"]


def pruning_model_with_mask(model, mask_dict, model_type):
    """`model` is the bare encoder (``model.lxmert``); `mask_dict` is what stage 2's save_model_mask wrote
    (keys '<model_type>.<module>.weight_mask', bool) or a state_dict-style dict (suffix '.weight')."""
    suffix = ".weight_mask" if "_mask" in list(mask_dict.keys())[0] else ".weight"
    for name in trained_mask_module_names():
        custom_from_mask(model, name, mask_dict["%s.%s%s" % (model_type, name, suffix)])


def mag_pruning(model, px):
    """prune.l1_unstructured(amount=px) on the language layers that exist, the pooler and the word embeddings --
    exactly the module list of the reference (it never names r_layers / x_layers)."""
    print("Start magnitude pruning with zero rate %.2f" % px)
    wanted = [f"encoder.layer.{i}.{s}" for i in range(12) for s in _ATT] + ["pooler.dense"]
    existing = dict(model.named_modules())
    for name in wanted:
        if name in existing:
            custom_from_mask(model, name, l1_unstructured_mask(existing[name].weight, px))
    emb = model.embeddings.word_embeddings
    custom_from_mask(model, "embeddings.word_embeddings", l1_unstructured_mask(emb.weight, px))


def see_weight_rate(model, model_type):
    """Percentage of zeros over the weight masks of the trained-mask module set (read from state_dict, :75-178)."""
    sd = model.state_dict()
    total = zeros = 0.0
    for name in trained_mask_module_names():
        m = sd["%s.%s.weight_mask" % (model_type, name)]
        total += float(m.nelement())
        zeros += float(torch.sum(m == 0))
    return 100 * zeros / total


class GroupedAdam(torch.optim.Adam):
    """torch.optim.Adam whose step() updates every parameter group with the same hyper-parameters in ONE fused
    multi-tensor call.  The reference builds one group per tensor (~500 groups); stepped group by group, each launch
    covers a single small tensor with a few dozen CTAs (measured: 15 ms for 6 GB of optimiser traffic).  The update
    rule, the state layout (`exp_avg`, `exp_avg_sq`, `step`) and `param_groups` are torch.optim.Adam's."""

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        buckets = {}
        for group in self.param_groups:
            key = (group["lr"] if not torch.is_tensor(group["lr"]) else id(group["lr"]), group["betas"], group["eps"],
                   group["weight_decay"], group["amsgrad"], group["maximize"], group.get("decoupled_weight_decay", False))
            b = buckets.setdefault(key, (group, [], [], [], [], [], []))
            self._init_group(group, b[1], b[2], b[3], b[4], b[5], b[6])
        for group, params, grads, exp_avgs, exp_avg_sqs, max_sqs, steps in buckets.values():
            if not params:
                continue
            beta1, beta2 = group["betas"]
            _adam_functional(params, grads, exp_avgs, exp_avg_sqs, max_sqs, steps, amsgrad=group["amsgrad"],
                                  has_complex=False, beta1=beta1, beta2=beta2, lr=group["lr"],
                                  weight_decay=group["weight_decay"], eps=group["eps"], maximize=group["maximize"],
                                  foreach=group["foreach"], capturable=group["capturable"],
                                  differentiable=group["differentiable"], fused=group["fused"],
                                  grad_scale=getattr(self, "grad_scale", None), found_inf=getattr(self, "found_inf", None),
                                  decoupled_weight_decay=group.get("decoupled_weight_decay", False))
        return loss


class ArenaAdam(GroupedAdam):
    """GroupedAdam that, once a hg_transformers._engine_ft.WeightArena is attached (Trainer._setup_engine), performs
    the whole step -- global-norm clip, Adam on every trainable tensor, refresh of the bf16 GEMM operands, gradient
    clearing -- as ONE launch over the arena (crv_adamw_segmented, torch.optim.Adam rule).  `param_groups`, the
    per-parameter state (`step`, `exp_avg`, `exp_avg_sq`: views of the arena's flat buffers) and the LambdaLR
    scheduler contract are torch.optim.Adam's; without an arena it IS GroupedAdam."""

    def __init__(self, params, **kw):
        super().__init__(params, **kw)
        self._arena = None
        self._clip = None
        self._hyper = None

    def attach_weight_arena(self, arena):
        for group in self.param_groups:
            if group["amsgrad"] or group["maximize"]:
                raise ValueError("the arena optimiser pass implements plain Adam (no amsgrad / maximize)")
            for p in group["params"]:
                views = arena.state_views(p)
                if views is None:
                    raise ValueError("every parameter of the optimiser must live in the arena")
                st = self.state[p]
                for k in ("exp_avg", "exp_avg_sq"):
                    if k in st:
                        views[k].copy_(st[k])
                    st[k] = views[k]
                if "step" not in st:
                    st["step"] = torch.tensor(0.0, dtype=torch.float32)
        self._arena = arena

    # -- engine hooks (same contract as optimization.AdamW) ----------------------------------------------------
    def set_clip(self, total_sumsq, max_norm):
        self._clip = (total_sumsq, float(max_norm))

    def use_device_hyper(self, hyper):
        self._hyper = hyper

    def ensure_state(self):
        pass

    def _steps(self):
        return [self.state[p]["step"] for g in self.param_groups for p in g["params"]]

    def advance_steps(self, n=1):
        n = int(n)
        if n:
            torch._foreach_add_(self._steps(), float(n))

    def _shared(self):
        g = self.param_groups[0]
        key = (g["lr"], g["betas"], g["eps"], g["weight_decay"])
        for other in self.param_groups[1:]:
            if (other["lr"], other["betas"], other["eps"], other["weight_decay"]) != key:
                raise RuntimeError("the arena optimiser pass needs one lr / betas / eps / weight_decay for all groups "
                                   "(true of init_optimizer + LambdaLR); set CRVQA_FT_ENGINE=0 otherwise")
        return g

    def hyper_values(self, next_step):
        g = self._shared()
        return ops.adam_hyper(ops.ADAM_TORCH, g["lr"], int(next_step), g["betas"][0], g["betas"][1])

    @torch.no_grad()
    def step(self, closure=None):
        if self._arena is None:
            return super().step(closure)
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        g = self._shared()
        p0 = g["params"][0]
        t = int(self.state[p0]["step"]) + 1
        clip_sumsq, max_norm = self._clip if self._clip is not None else (None, 1.0)
        self._clip = None
        self._arena.adam_step(g["lr"], t, g["betas"][0], g["betas"][1], g["eps"], g["weight_decay"], clip_sumsq,
                              max_norm, self._hyper)
        self.advance_steps(1)
        return loss


def init_optimizer(model, training_args, num_train_data):
    """torch.optim.Adam with one param group per tensor + linear schedule (:577-598)."""
    params = [{"params": [value], "name": key, "weight_decay": training_args.weight_decay,
               "param_size": value.size(), "nelement": value.nelement(), "lr": training_args.learning_rate}
              for key, value in model.named_parameters() if value.requires_grad]
    # same update rule as the reference's torch.optim.Adam(params, lr, betas, eps): fused implementation on CUDA, and
    # all groups stepped by one multi-tensor call (GroupedAdam)
    on_cuda = all(g["params"][0].is_cuda for g in params) and len(params) > 0
    optimizer = ArenaAdam(params, lr=training_args.learning_rate, betas=(0.9, 0.999),
                          eps=training_args.adam_epsilon, **({"fused": True} if on_cuda else {}))
    num_training_steps = int(int(num_train_data / (max(1, training_args.n_gpu) * training_args.per_gpu_train_batch_size) + 1)
                             * training_args.num_train_epochs)
    scheduler = get_linear_schedule_with_warmup(optimizer, num_warmup_steps=training_args.warmup_steps,
                                                num_training_steps=num_training_steps)
    return optimizer, scheduler


__all__ = ["ArenaAdam", "GroupedAdam", "PrunedLinear", "PrunedEmbedding", "pruning_model_with_mask", "mag_pruning", "see_weight_rate",
           "init_optimizer", "trained_mask_module_names"]
