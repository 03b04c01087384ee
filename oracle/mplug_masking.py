"""Oracle restatement (torch-CPU, fp32 / float64) of the reference's mPLUG masking path:
mPLUG/masking/maskers.py and the mask-update block of mPLUG/vqa_mplug.py:202-210.

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Pinned against outputs of the unmodified reference modules on the
miniature mPLUG-shaped network: tests/golden/make_golden_mplug.py -> tests/golden/mplug_skeleton.pt, checked by
tests/test_mplug_cpu.py.
"""
import numpy as np
import torch
import torch.nn as nn

from . import masked_ops as o

# --------------------------------------------------------------------------- names (mPLUG/masking/maskers.py:16-82)
_BERT_SELF = {"K": "attention.self.key", "Q": "attention.self.query", "V": "attention.self.value",
              "AO": "attention.output.dense", "I": "intermediate.dense", "O": "output.dense"}
_BERT_CROSS = {"SK": "attention.self.key", "SQ": "attention.self.query", "SV": "attention.self.value",
               "SAO": "attention.output.dense", "CK": "crossattention.self.key",
               "CQ": "crossattention.self.query", "CV": "crossattention.self.value",
               "CAO": "crossattention.output.dense", "I": "intermediate.dense", "O": "output.dense"}
_CLIP = {"AO": "attn.out_proj", "I": "mlp.c_fc", "O": "mlp.c_proj"}


def module_name(tower, abbre, layer):
    """One entry of the reference's four lambda tables."""
    if tower == "visual_encoder":
        if abbre == "E":
            return f"{tower}.token_embedding"
        if abbre.endswith("_visual"):
            return f"{tower}.visual.transformer.resblocks.{layer}.{_CLIP[abbre[:-7]]}"
        return f"{tower}.transformer.resblocks.{layer}.{_CLIP[abbre]}"
    prefix = f"{tower}.bert." if tower == "text_decoder" else f"{tower}."
    if abbre == "E":
        return prefix + "embeddings.word_embeddings"
    table = _BERT_SELF if tower == "text_encoder" else _BERT_CROSS
    return f"{prefix}encoder.layer.{layer}.{table[abbre]}"


def chain_module_names(tower, layers, abbres):
    """:64-82 -- every name also under `<tower>_m` (str.replace of the tower name, as the reference does)."""
    names = {module_name(tower, a, l) for a in abbres for l in layers}
    return names | {n.replace(tower, tower + "_m") for n in names}


# --------------------------------------------------------------------------- initialisation (:188-234)
def rank(numel, sparsity):
    return int(numel * sparsity)


def magnitude_soft_init(weight, init_sparsity):
    """_magnitude_soft (:215-220): score = |W|; threshold = kthvalue(|W|, k), or the int 0 when k == 0."""
    k = rank(weight.numel(), init_sparsity)
    thr = torch.tensor(float(o.kth_value(weight, k, use_abs=True))) if k > 0 else 0
    return weight.detach().abs(), thr


def magnitude_init(weight, init_sparsity, threshold, global_cut=None):
    """_magnitude (:204-213, cut 0 when k == 0) / _magnitude_global (:222-234): 2*thr where |W| > cut else 0."""
    if global_cut is None:
        k = rank(weight.numel(), init_sparsity)
        global_cut = float(o.kth_value(weight, k, use_abs=True)) if k > 0 else 0.0
    s = torch.zeros_like(weight)
    s[weight.abs() > float(global_cut)] = 2.0 * float(threshold)
    return s


def global_weight_cut(weights, init_sparsity):
    """Masker.compute_global_threshold (:535-545)."""
    k = rank(sum(w.numel() for w in weights), init_sparsity)
    return float(o.global_kth_value(weights, k, use_abs=True))


# --------------------------------------------------------------------------- binarisation in the two precisions
def mask_of(scores, threshold, score_dtype=torch.float32):
    """binarizer_fn1 (:326-328) `(inputs > threshold).type(inputs.type())` as torch evaluates it: with fp32 scores
    the 0-dim threshold is promoted to fp32; with bf16 scores (the DeepSpeed-bf16 model copy) BOTH operands are
    compared in bf16, the threshold being rounded first."""
    thr = threshold if torch.is_tensor(threshold) else torch.tensor(float(threshold))
    if score_dtype == torch.bfloat16:
        return (scores.detach().to(torch.bfloat16) > thr.to(torch.bfloat16)).float()
    return (scores.detach().float() > thr.float()).float()


# --------------------------------------------------------------------------- threshold refresh (:680-703)
def reset_threshold_module(scores, weight_numel, tgt_sparsity, old_threshold, score_dtype=torch.float32):
    """One module of reset_threshold: float64 k-th value -> bfloat16; kept only if below the largest score (compared
    in the promoted dtype: bf16 against fp32 scores -> fp32; against bf16 scores -> bf16)."""
    k = rank(weight_numel, tgt_sparsity)
    if k == 0:
        return old_threshold
    s = scores.detach().to(score_dtype)
    kth = torch.kthvalue(s.reshape(-1).to(torch.float64), k).values.to(torch.bfloat16)
    return kth if bool(kth < s.max()) else old_threshold


def reset_threshold_global(scores, tgt_sparsity):
    flat = torch.cat([s.detach().reshape(-1) for s in scores])
    return torch.kthvalue(flat, rank(flat.numel(), tgt_sparsity)).values


def thresholds_mean(thresholds):
    """`float(torch.tensor(thresholds).mean())` (:703): the list's dtype is the promotion of its entries -- all-bf16
    lists average in bf16."""
    return float(torch.tensor([t.cpu() if torch.is_tensor(t) else t for t in thresholds]).mean())


# --------------------------------------------------------------------------- reports (:705-745)
EXCLUDE = ["visual_encoder.transformer"] + [f"fusion_encoder.encoder.layer.{i}" for i in range(6)]


def see_sparsity(masks, named_param_sizes):
    zeros = sum(int((m == 0).sum()) for m in masks)
    total = sum(n for name, n in named_param_sizes
                if not name.endswith(".weight_mask") and "embedding" not in name
                and not any(p in name for p in EXCLUDE))
    return 100.0 * zeros / total


def zero_rate(masks):
    return 100.0 * sum(int((m == 0).sum()) for m in masks) / sum(m.numel() for m in masks)


# --------------------------------------------------------------------------- a masked module for CPU forward/backward
class OracleMasked(nn.Module):
    """MaskedLinear1 (:334-358) on the oracle's masked_ops, with the attributes the reference's helpers look for."""

    def __init__(self, name, weight, bias, padding_idx, scores, threshold, operand="fp32"):
        super().__init__()
        self.name, self.weight, self.bias, self.padding_idx = name, weight, bias, padding_idx
        self.weight_mask = nn.Parameter(scores)
        self.threshold = threshold
        self.operand = operand
        self.score_dtype = torch.float32

    def get_masks(self):
        return mask_of(self.weight_mask, self.threshold, self.score_dtype), None

    def forward(self, x):
        m = self.get_masks()[0]
        if "embedding" in self.name:
            return _MaskedEmbeddingWithMask.apply(x, self.weight_mask, self.weight, m, self.padding_idx)
        return _MaskedLinearWithMask.apply(x, self.weight_mask, self.weight, m, self.bias, self.operand)


class _MaskedLinearWithMask(torch.autograd.Function):
    """o.MaskedLinear with the mask given (the bf16 score mode changes how the mask is derived, nothing else)."""

    @staticmethod
    def forward(ctx, x, scores, weight, mask, bias, operand):
        xr = o._round(x.detach().reshape(-1, x.shape[-1]), operand)
        wr = o._round(weight.detach(), operand)
        wm = wr * mask
        ctx.save_for_backward(xr, wr, wm)
        ctx.operand, ctx.x_shape = operand, x.shape
        return torch.nn.functional.linear(xr, wm, bias).view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        xr, wr, wm = ctx.saved_tensors
        dyr = o._round(dy.reshape(-1, dy.shape[-1]), ctx.operand)
        dx = (dyr @ wm).view(ctx.x_shape) if ctx.needs_input_grad[0] else None
        return dx, (dyr.t() @ xr) * wr, None, None, None, None


class _MaskedEmbeddingWithMask(torch.autograd.Function):
    @staticmethod
    def forward(ctx, ids, scores, weight, mask, padding_idx):
        ctx.save_for_backward(ids, weight)
        ctx.padding_idx = padding_idx
        return torch.nn.functional.embedding(ids, weight * mask, padding_idx=padding_idx)

    @staticmethod
    def backward(ctx, dout):
        ids, weight = ctx.saved_tensors
        flat = ids.reshape(-1)
        d = dout.reshape(-1, dout.shape[-1]).clone()
        if ctx.padding_idx is not None:
            d[flat == ctx.padding_idx] = 0
        return None, torch.zeros_like(weight).index_add_(0, flat, d) * weight, None, None, None


REPLACED = (nn.Linear, nn.Embedding, nn.modules.linear.NonDynamicallyQuantizableLinear)


def patch(model, names, *, init_sparsity, threshold=1e-2, controlled_init="magnitude_soft", global_prune=False,
          train_classifier=True, operand="fp32"):
    """Masker.patch_modules / replace (:547-678) for MaskedLinear1, unstructured: freeze, swap, initialise.
    Returns the global |W| cut (or None)."""
    cut = None
    if global_prune:
        cut = global_weight_cut([m.weight for n, m in model.named_modules() if n in names], init_sparsity)

    def walk(m, root):
        for attr in dir(m):
            try:
                child = getattr(m, attr)
            except Exception:
                continue
            if not isinstance(child, nn.Module):
                continue
            name = f"{root}.{attr}" if root else attr
            spared = "predictions" in name or ("classifier" in name and train_classifier)
            for pn in ("weight", "bias"):
                p = getattr(child, pn, None)
                if isinstance(p, torch.Tensor) and not spared:
                    p.requires_grad = False
            for pn in ("in_proj_weight", "positional_embedding", "class_embedding"):
                p = getattr(child, pn, None)
                if isinstance(p, torch.Tensor):
                    p.requires_grad = False
            if type(child) in REPLACED and name in names:
                w = child.weight
                if controlled_init == "magnitude_soft":
                    s, thr = magnitude_soft_init(w, init_sparsity)
                else:
                    s, thr = magnitude_init(w, init_sparsity, threshold, cut), torch.tensor(threshold)
                setattr(m, attr, OracleMasked(name, w, getattr(child, "bias", None),
                                              getattr(child, "padding_idx", None), s, thr, operand))
        for cn, c in m.named_children():
            walk(c, f"{root}.{cn}" if root else cn)

    walk(model, "")
    return cut


def masked_modules(model):
    return [(n, m) for n, m in model.named_modules() if hasattr(m, "threshold")]


def reset_threshold(model, tgt_sparsity, global_prune=False):
    mods = masked_modules(model)
    if global_prune:
        cut = reset_threshold_global([m.weight_mask for _, m in mods], tgt_sparsity)
        for _, m in mods:
            m.threshold = cut
    else:
        for _, m in mods:
            m.threshold = reset_threshold_module(m.weight_mask, m.weight.numel(), tgt_sparsity, m.threshold,
                                                 m.score_dtype)
    return thresholds_mean([m.threshold for _, m in mods])


def packed(mask):
    return np.packbits((mask.detach().cpu().float() != 0).numpy().reshape(-1))
