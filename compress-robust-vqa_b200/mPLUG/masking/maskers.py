"""Drop-in for the reference's ``mPLUG/masking/maskers.py`` (mPLUG masked training, BASELINE config 5).

Same public names as the reference module: the four name tables and ``chain_module_names`` (:16-82, every name
also gets its ``<ptl>_m`` momentum twin), ``MaskedLinearX`` / ``MaskedLinear0-3`` with the ``global_prune`` /
``magnitude_soft`` options (:85-284), ``Masker`` (:509-678) and the module-level ``reset_threshold`` (:680-703),
``see_sparsity`` (:708-722) and ``save_model_mask`` (:724-745).  The masked-module bodies are the shared ones of
``masking/_core.py`` (tcgen05 masked GEMMs, exact batched select, CUDA binariser -- libcrvqa.so); there is no CPU
implementation.

Score precision.  The reference trains under DeepSpeed bf16 (mPLUG/configs/ds_config.json): the model copy of every
score is bf16(master score), thresholds are compared after rounding to bf16, and ``reset_threshold`` rounds the new
threshold to bf16 (:697).  Here scores stay fp32 (the master copy, which is what the optimiser updates);
``set_score_dtype(model, torch.bfloat16)`` makes every masked module reproduce the bf16 comparison bit for bit by
comparing the fp32 score with the one fp32 value T for which ``bf16(S) > bf16(thr)  <=>  S > T`` (round-to-nearest-
even is monotone; ``bf16_score_threshold``).
"""
import json
import os

import numpy as np
import torch
import torch.nn as nn

from . import _core as core
from ._core import (  # noqa: F401
    MaskedLinear0, _Binarizer1, _Binarizer2, _Binarizer3, _get_nnz_from, _scheme_idx_to_fn, binarizer_fn1,
    binarizer_fn2, binarizer_fn3, reshape_mask_for_sp,
)

ops = core.ops

# --------------------------------------------------------------------------- name tables (reference :16-62)
_BLOCK = {"AO": "attn.out_proj", "I": "mlp.c_fc", "O": "mlp.c_proj"}
_visual_spec = {a + "_visual": "visual.transformer.resblocks.{l}." + p for a, p in _BLOCK.items()}
_visual_spec.update({a: "transformer.resblocks.{l}." + p for a, p in _BLOCK.items()})
_visual_spec["E"] = "token_embedding"
_visual_encoder_names = core._table(_visual_spec)

_text_encoder_names = core._table({**{a: "encoder.layer.{l}." + p for a, p in core._ATT.items()},
                                   "E": "embeddings.word_embeddings"})

_CROSS = {"SK": "attention.self.key", "SQ": "attention.self.query", "SV": "attention.self.value",
          "SAO": "attention.output.dense", "CK": "crossattention.self.key", "CQ": "crossattention.self.query",
          "CV": "crossattention.self.value", "CAO": "crossattention.output.dense", "I": "intermediate.dense",
          "O": "output.dense"}
_fusion_encoder_names = core._table({**{a: "encoder.layer.{l}." + p for a, p in _CROSS.items()},
                                     "E": "embeddings.word_embeddings"})
_text_decoder_names = core._table({**{a: "bert.encoder.layer.{l}." + p for a, p in _CROSS.items()},
                                   "E": "bert.embeddings.word_embeddings"})

_TABLES = {"visual_encoder": _visual_encoder_names, "text_encoder": _text_encoder_names,
           "text_decoder": _text_decoder_names, "fusion_encoder": _fusion_encoder_names}


def chain_module_names(which_ptl, layer_idices, abbres):
    """Module names of one tower plus the same names under ``<which_ptl>_m`` (the distillation twin)."""
    names = core.chain_names_plain(_TABLES[which_ptl], which_ptl, layer_idices, abbres)
    return names | {n.replace(which_ptl, which_ptl + "_m") for n in names}


# --------------------------------------------------------------------------- bf16 score comparison
def bf16_score_threshold(thr):
    """fp32 tensor T (same shape as ``thr``) with  bf16(S) > bf16(thr)  <=>  S > T  for every finite fp32 S.

    With a bf16 score tensor and a 0-dim threshold, ``scores > thr`` is evaluated in bf16 (type promotion keeps the
    dimensioned operand's dtype), i.e. on RNE-rounded values.  Rounding is monotone, so the set {S : bf16(S) > t16}
    is a half line that starts at the midpoint between t16 and the next bf16 above it; the midpoint itself rounds
    to the neighbour whose last mantissa bit is even."""
    t16 = thr.detach().to(torch.float32).to(torch.bfloat16)
    b = t16.view(torch.int16).to(torch.int32) & 0xFFFF
    mag = b & 0x7FFF
    up = ((b & 0x8000) == 0) | (mag == 0)          # value >= 0: the next bf16 above has the larger magnitude
    nxt = torch.where(up, mag + 1, mag - 1)
    mid = torch.where(up, (mag << 16) + 0x8000, (mag << 16) - 0x8000)   # fp32 magnitude bits of the midpoint
    tie_goes_up = (nxt & 1) == 0                   # then S == midpoint already counts: take the fp32 just below it
    mid = torch.where(tie_goes_up, torch.where(up, mid - 1, mid + 1), mid).to(torch.int64)
    bits = torch.where(up, mid, mid - (1 << 31))   # sign bit of an int32 pattern (two's complement)
    return bits.to(torch.int32).view(torch.float32)


def _compute_device(t):
    """Where a set-up-time op on ``t`` runs: its own device, or the current GPU for a model still on the host."""
    if t.is_cuda or not torch.cuda.is_available():
        return t.device
    return torch.device("cuda", torch.cuda.current_device())


def set_score_dtype(model, dtype):
    """Select how every masked module of ``model`` compares scores: ``torch.float32`` (plain fp32 compare) or
    ``torch.bfloat16`` (the reference's DeepSpeed-bf16 behaviour, see the module docstring)."""
    if dtype not in (torch.float32, torch.bfloat16):
        raise ValueError("score dtype must be torch.float32 or torch.bfloat16")
    for m in model.modules():
        if isinstance(m, MaskedLinearX):
            m.score_dtype = dtype
            m._thr_dev = None
            m._thr_key = None


# --------------------------------------------------------------------------- masked modules
class MaskedLinearX(core.MaskedLinearX):
    """The reference's constructor signature (``global_prune`` positional after ``mask_biases``, :86-88) and its two
    extra initialisations: ``magnitude_soft`` (score = |W|, threshold = k-th |W|, :215-220) and the global magnitude
    cut (:222-234).  A rank of 0 (``init_sparsity`` 0, the default configuration) means "cut at 0" (:208,:218)."""

    score_dtype = torch.float32

    def __init__(self, scheme_idx, weight, bias, mask_biases, global_prune=False, **kwargs):
        # read by controlled_init / _uses_magnitude while the base constructor runs
        object.__setattr__(self, "global_prune", global_prune)
        object.__setattr__(self, "global_threshold", kwargs.get("global_threshold"))
        core.MaskedLinearX.__init__(self, scheme_idx, weight, bias, mask_biases, **kwargs)

    def _uses_magnitude(self, kind):
        return kind in ("magnitude", "magnitude_soft") or (kind == "magnitude_and_uniform" and "bert" in self.name)

    def controlled_init(self, weight, init_sparsity, threshold, controlled_init_type):
        if controlled_init_type not in ("magnitude", "magnitude_soft"):
            return super().controlled_init(weight, init_sparsity, threshold, controlled_init_type)
        k = self.num_zero_elements(weight, init_sparsity)
        thr = float(threshold)
        if controlled_init_type == "magnitude" and self.global_prune:
            assert self.global_threshold is not None, \
                "Compute the global magnitude threshold before initializating the weight_mask!"
            cut = ops.as_thr(self.global_threshold, _compute_device(weight)).reshape(1)
        elif k > 0:
            cut = ops.kth_value_batched([weight.detach().float()], [k], use_abs=True)[0:1]
        else:
            cut = None
        if controlled_init_type == "magnitude_soft":
            self.threshold = cut[0].to(weight.device) if cut is not None else 0
            return nn.Parameter(weight.detach().abs())
        if cut is None:
            cut = torch.zeros(1)
        return nn.Parameter(ops.magnitude_init(weight, cut, 2.0 * thr, 0.0 * thr))


class MaskedLinear1(MaskedLinearX, core.MaskedLinear1):
    """Scheme 1 (:334-358): forward == F.linear(x, weight * (weight_mask > threshold), bias) or, for names containing
    'embedding', F.embedding of the masked table -- the fused sm_100a kernels of the shared core."""

    def __init__(self, weight, bias, mask_biases, global_prune=False, **kwargs):
        MaskedLinearX.__init__(self, "MaskedLinear1", weight, bias, mask_biases, global_prune, **kwargs)
        self._w16 = self._w16_key = self._thr_dev = self._thr_key = None

    def _threshold_on(self, device):
        if self.score_dtype != torch.bfloat16:
            return core.MaskedLinear1._threshold_on(self, device)
        t = self.threshold
        key = (id(t), t._version if torch.is_tensor(t) else t, device, "bf16")
        if self._thr_dev is None or self._thr_key != key:
            self._thr_dev = bf16_score_threshold(ops.as_thr(t, device))
            self._thr_key = key
        return self._thr_dev


class MaskedLinear2(MaskedLinearX, core.MaskedLinear2):
    def __init__(self, weight, bias, mask_biases, global_prune=False, **kwargs):
        MaskedLinearX.__init__(self, "MaskedLinear2", weight, bias, mask_biases, global_prune, **kwargs)


class MaskedLinear3(MaskedLinearX, core.MaskedLinear3):
    def __init__(self, weight, bias, mask_biases, global_prune=False, **kwargs):
        MaskedLinearX.__init__(self, "MaskedLinear3", weight, bias, mask_biases, global_prune, **kwargs)


_MASKED_CLASSES = {"MaskedLinear0": MaskedLinear0, "MaskedLinear1": MaskedLinear1,
                   "MaskedLinear2": MaskedLinear2, "MaskedLinear3": MaskedLinear3}
_REPLACED_TYPES = (nn.Linear, nn.Embedding, nn.modules.linear.NonDynamicallyQuantizableLinear)


def _finish_deferred_init(modules, global_cut):
    """|W| selects of all modules whose initialisation was deferred, as ONE batched exact select."""
    pend = [m for m in modules if getattr(m, "_pending_magnitude", False)]
    ranked = [(m, MaskedLinearX.num_zero_elements(m.weight, m._init_sparsity)) for m in pend]
    def global_cut_applies(m):
        return m._controlled_init == "magnitude" and m.global_prune

    need = [(m, k) for m, k in ranked if k > 0 and not global_cut_applies(m)]
    cuts = {}
    if need:
        vals = ops.kth_value_batched([m.weight.detach().float() for m, _ in need], [k for _, k in need], use_abs=True)
        cuts = {id(m): vals[i:i + 1] for i, (m, _) in enumerate(need)}
    for m, k in ranked:
        thr = float(m.threshold)
        cut = cuts.get(id(m))
        if m._controlled_init == "magnitude_soft":
            m.weight_mask.data = m.weight.detach().abs()
            m.threshold = cut[0].to(m.weight.device) if cut is not None else 0
        else:
            if global_cut_applies(m):
                cut = global_cut.reshape(1)
            elif cut is None:
                cut = torch.zeros(1)
            m.weight_mask.data = ops.magnitude_init(m.weight, cut.contiguous(), 2.0 * thr, 0.0 * thr)
        m._pending_magnitude = False


# --------------------------------------------------------------------------- the Masker
class Masker(core.MaskerBase):
    def __init__(self, masker_scheduler, logger, mask_biases, structured_masking_info, threshold, init_scale,
                 controlled_init, train_classifier=False, global_prune=False):
        self._setup(masker_scheduler, logger, mask_biases, structured_masking_info, threshold, init_scale,
                    "text_encoder", controlled_init)
        self.train_classifier = train_classifier
        self.global_prune = global_prune
        self.global_threshold = None

    def compute_global_threshold(self, model, names_tobe_masked):
        """ONE |W| cut over the union of all weights to be masked (:535-545)."""
        self.logger.info("Computing global threshold...")
        weights = [m.weight for n, m in model.named_modules() if n in names_tobe_masked]
        k = int(sum(w.numel() for w in weights) * self.masker_scheduler.init_sparsity)
        self.global_threshold = core.global_kth_value([w.float() for w in weights], k, use_abs=True)[0]

    def patch_modules(self, model, names_tobe_masked, name_of_masker="MaskedLinear1"):
        self.ptl_config = model.text_encoder.config
        masked_linear_cls = _MASKED_CLASSES[name_of_masker]
        if self.global_prune:
            self.compute_global_threshold(model, names_tobe_masked)
        self._created = []
        self.replace(model, "", names_tobe_masked, masked_linear_cls)
        _finish_deferred_init(self._created, self.global_threshold)
        self.masked_linear_cls = masked_linear_cls

        self.logger.info("Check the masking status.")
        for m_name, m in model.named_modules():
            if m_name not in names_tobe_masked:
                continue
            if isinstance(m, masked_linear_cls):
                param_info = {}
                for _name, param in m.named_parameters():
                    if "mask" in _name:
                        kept = _get_nnz_from(self.eval_binarizer_fn(name_of_masker, param.detach(), m.threshold))
                        param_info[_name] = float(1.0 - kept / np.prod(param.shape))
                        self.init_masks[f"{m_name}_{_name}"] = self.eval_binarizer_fn(
                            name_of_masker, param.detach(), self.threshold).cpu()
                print(f"\t {m_name} is MASKED -> {json.dumps(param_info)}")
            else:
                print(f"\t {m_name} is NOT MASKED")

    def _freeze(self, name, module):
        """Freeze rules of :603-616: everything but the LM head ('predictions') and, with train_classifier, the
        classifier; fused attention projections and the ViT position / class embeddings as well."""
        spared = "predictions" in name or ("classifier" in name and self.train_classifier)
        for pname in ("weight", "bias"):
            p = getattr(module, pname, None)
            if isinstance(p, torch.Tensor) and not spared:
                p.requires_grad = False
        for pname in ("in_proj_weight", "positional_embedding", "class_embedding"):
            p = getattr(module, pname, None)
            if isinstance(p, torch.Tensor):
                p.requires_grad = False

    def replace(self, m, root_name, names_tobe_masked, masked_linear_cls, **_ignored):
        for attr_str in dir(m):
            try:
                target_attr = getattr(m, attr_str)
            except Exception:
                continue
            if not isinstance(target_attr, nn.Module):
                continue
            name = root_name + "." + attr_str if root_name else attr_str
            self._freeze(name, target_attr)
            if type(target_attr) not in _REPLACED_TYPES:
                continue
            masked = False
            if name in names_tobe_masked:
                masked_linear = masked_linear_cls(
                    name=name, weight=target_attr.weight, bias=getattr(target_attr, "bias", None),
                    padding_idx=getattr(target_attr, "padding_idx", None), mask_biases=self.mask_biases,
                    threshold=self.threshold, init_sparsity=self.masker_scheduler.init_sparsity,
                    init_scale=self.init_scale, controlled_init=self.controlled_init,
                    structured_masking_info={"ptl_config": self.ptl_config, **self.structured_masking_info},
                    global_threshold=self.global_threshold, global_prune=self.global_prune,
                    _defer_magnitude_init=True)
                for _name, param in masked_linear.named_parameters():
                    if "mask" not in _name:
                        param.requires_grad = False
                if masked_linear.unstructured_masked or masked_linear.structured_masked:
                    masked = True
                    setattr(m, attr_str, masked_linear)
                    self._created.append(masked_linear)
                    kind = (f"structured masking for layer type="
                            f"{self.structured_masking_info['structured_masking_types']}"
                            if masked_linear.structured_masked else "unstructured masking")
                    self.logger.info(f"\t {name} is MASKED: {kind}")
            if not masked:
                self.logger.info(f"\t {name} is NOT MASKED")

        for child_name, child in m.named_children():
            self.replace(child, root_name + "." + child_name if root_name else child_name, names_tobe_masked,
                         masked_linear_cls)


# --------------------------------------------------------------------------- threshold refresh, reports
def _masked_modules(model):
    return [(n, m) for n, m in model.named_modules() if hasattr(m, "threshold")]


def reset_threshold(model, tgt_sparsity, global_prune=False):
    """New thresholds for the target sparsity (:680-703); returns their mean.

    Per module: k = int(weight.numel() * tgt_sparsity); nothing happens for k == 0; otherwise the k-th smallest score
    is rounded to bf16 (the reference selects in float64 -- exact for fp32 / bf16 inputs -- and casts the result to
    bfloat16) and becomes the module's threshold unless it is not below the largest score ("all the values in
    weight_mask are the same": the old threshold is kept).  All modules go through ONE batched exact select that
    returns rank k and rank n (the maximum) of every score tensor.  ``global_prune``: one fp32 threshold, rank
    int(total * tgt_sparsity) of the union of all scores."""
    mods = _masked_modules(model)
    if global_prune:
        scores = [m.weight_mask.detach().float() for _, m in mods]
        k = int(sum(s.numel() for s in scores) * tgt_sparsity)
        cut = core.global_kth_value(scores, k)[0]
        for _, m in mods:
            m.threshold = cut
        return _mean_threshold([cut] * len(mods))      # the fp32 mean of n equal values need not be that value

    ranked = [(m, int(m.weight.nelement() * tgt_sparsity)) for _, m in mods]
    live = [(m, k) for m, k in ranked if k > 0]
    if live:
        scores = [m.weight_mask.detach().float() for m, _ in live]
        both = ops.kth_value_batched(scores + scores, [k for _, k in live] + [s.numel() for s in scores])
        kth16 = both[:len(live)].to(torch.bfloat16)
        top = both[len(live):]
        bf16_scores = torch.tensor([m.score_dtype == torch.bfloat16 for m, _ in live], device=top.device)
        top = torch.where(bf16_scores, top.to(torch.bfloat16).float(), top)
        moved = (kth16.float() < top).tolist()                      # the one host sync of a refresh
        for i, (m, _) in enumerate(live):
            if moved[i]:
                m.threshold = kth16[i]
    return _mean_threshold([m.threshold for m, _ in ranked])


def _mean_threshold(thresholds):
    """``float(torch.tensor(thresholds).mean())`` of the reference (:703).  The list's dtype is the promotion of its
    entries, so a list of bf16 thresholds is averaged and rounded in bf16; device entries come back in ONE copy.
    A list of Python ints only (every rank was 0 and no threshold has ever been set) makes the reference raise
    inside ``mean()``; here it is averaged as floats."""
    on_dev = [t for t in thresholds if torch.is_tensor(t) and t.is_cuda]
    if on_dev:
        host = iter(torch.stack([t.detach().float() for t in on_dev]).cpu())    # bf16 -> fp32 is exact
        thresholds = [next(host).to(t.dtype) if (torch.is_tensor(t) and t.is_cuda) else t for t in thresholds]
    listed = torch.tensor([t.detach() if torch.is_tensor(t) else t for t in thresholds])
    return float((listed if listed.is_floating_point() else listed.float()).mean())


# towers the VQA model never runs (:705-706)
exclude_prefix = ["visual_encoder.transformer"] + [f"fusion_encoder.encoder.layer.{i}" for i in range(6)]


def _mask_and_zeros(module):
    """(0/1 weight mask, number of zeros): ``module.get_masks()[0]`` of the reference, for scheme 1 through the CUDA
    binariser with its fused kept-count."""
    if isinstance(module, core.MaskedLinear1) and not module.structured_masked:
        scores = module.weight_mask.detach()
        mask, kept = ops.binarize(scores.float(), module._threshold_on(_compute_device(scores)), want_count=True)
        return mask, mask.numel() - int(kept)
    mask = module.get_masks()[0].detach()
    return mask, int((mask == 0).sum())


def see_sparsity(model):
    """Zeroed weights of the masked modules over all parameters that are neither scores, embeddings nor part of an
    unused tower (:708-722).  Prints like the reference and also returns the percentage."""
    print("\n\n")
    print("Checking zero rate...")
    num_zero = sum(_mask_and_zeros(m)[1] for _, m in _masked_modules(model))
    num_total = sum(p.nelement() for n, p in model.named_parameters()
                    if not n.endswith(".weight_mask") and "embedding" not in n
                    and not any(prefix in n for prefix in exclude_prefix))
    print("=" * 100)
    sparsity = 100 * num_zero / num_total
    print(f"Sparsity of entire model = {sparsity:.2f}")
    print("\n\n")
    return sparsity


def save_model_mask(model, output_dir=None, is_save=True):
    """Collect ``{<module name>.weight: 0/1 mask (CPU)}``, print the zero rate, optionally write mask.pt (:724-745)."""
    mask_dict = {}
    zero_sum, elem_sum = 0, 0
    print("\n\n")
    print("Collecting mask...")
    for name, module in _masked_modules(model):
        mask, zeros = _mask_and_zeros(module)
        zero_sum += zeros
        elem_sum += mask.numel()
        mask_dict[name + ".weight"] = mask.cpu()
    zero_rate = 100 * zero_sum / max(elem_sum, 1)
    print(f"Zero rate of entire model = {zero_rate:.2f}")
    if is_save:
        print("Saving model mask to %s", output_dir)
        os.makedirs(output_dir, exist_ok=True)
        torch.save(mask_dict, os.path.join(output_dir, "mask.pt"))
    print("\n\n")
    return zero_rate
