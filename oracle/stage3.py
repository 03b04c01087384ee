"""ORACLE (test infrastructure only -- see oracle/__init__.py): CPU restatement of the stage-3 frozen-mask
fine-tune.

The reference applies `torch.nn.utils.prune.CustomFromMask` to every masked module
(run_vqa_stage3.py:227-297 `pruning_model_with_mask`) or `prune.l1_unstructured` (`mag_pruning`,
run_vqa_stage3.py:205-225).  Both reparametrise a module as

    weight_orig (Parameter, trainable)   weight_mask (buffer, 0/1 fp32)   weight = weight_orig * weight_mask

recomputed by a forward pre-hook on every call, so autograd gives  dW_orig = dW (.) mask,  every other parameter
of the network (biases, LayerNorms, embeddings, head) trains as usual, and the optimiser is torch.optim.Adam
(run_vqa_stage3.py:577-598).  torch.nn.utils.prune is part of PyTorch (the reference's only dependency on this
path); its semantics restated here: CustomFromMask -> mask = ones * given mask; l1_unstructured(amount=px) ->
round(px * n) entries of smallest |w| (torch.topk(..., largest=False)) get mask 0.

Pinned by tests/golden/stage3_full.pt (tests/golden/make_golden_stage3.py runs the reference functions).
"""
import numpy as np
import torch

from . import lxmert_oracle as lxo

ATT = ("attention.self.query", "attention.self.key", "attention.self.value", "attention.output.dense",
       "intermediate.dense", "output.dense")
XSUB = ("visual_attention.att.query", "visual_attention.att.key", "visual_attention.att.value",
        "visual_attention.output.dense", "lang_self_att.self.query", "lang_self_att.self.key",
        "lang_self_att.self.value", "lang_self_att.output.dense", "visn_self_att.self.query",
        "visn_self_att.self.key", "visn_self_att.self.value", "visn_self_att.output.dense",
        "lang_inter.dense", "lang_output.dense", "visn_inter.dense", "visn_output.dense")


def trained_mask_modules():
    """Modules `pruning_model_with_mask` reparametrises, in its order (run_vqa_stage3.py:231-294)."""
    out = [f"encoder.layer.{i}.{s}" for i in range(9) for s in ATT]
    out += [f"encoder.r_layers.{i}.{s}" for i in range(5) for s in ATT]
    out += [f"encoder.x_layers.{i}.{s}" for i in range(5) for s in XSUB]
    return out + ["pooler.dense", "embeddings.word_embeddings", "encoder.visn_fc.visn_fc", "encoder.visn_fc.box_fc"]


def mag_pruning_modules(existing):
    """Modules `mag_pruning` prunes (run_vqa_stage3.py:205-225): encoder.layer.0..11 x six sub-modules (only the
    ones that exist), pooler.dense, then embeddings.word_embeddings."""
    names = [f"encoder.layer.{i}.{s}" for i in range(12) for s in ATT] + ["pooler.dense"]
    return [n for n in names if n in existing] + ["embeddings.word_embeddings"]


def l1_unstructured_mask(weight, amount):
    """prune.l1_unstructured: k = round(amount * n); the k entries of smallest |w| are masked out."""
    n = weight.numel()
    k = round(amount * n)
    mask = torch.ones_like(weight)
    if k > 0:
        idx = torch.topk(weight.detach().abs().reshape(-1), k=k, largest=False).indices
        mask.view(-1)[idx] = 0
    return mask


def magnitude_mask(weight, zero_rate):
    """The stand-in for a trained stage-2 mask used by the golden file: |W| > kthvalue(|W|, max(1, int(rate n)))."""
    k = max(1, int(weight.numel() * zero_rate))
    a = np.abs(weight.detach().cpu().numpy().reshape(-1))
    thr = np.partition(a, k - 1)[k - 1]
    return (weight.detach().abs() > float(thr))


class PrunedLinear(torch.autograd.Function):
    """y = x (W_orig (.) M)^T + b with the gradients autograd derives for the reparametrised module:
    dX = dY (W_orig (.) M),  dW_orig = (dY^T X) (.) M,  db = sum dY.  operand='bf16' rounds the MMA operands the
    way the CUDA path does (fp32 accumulate)."""

    @staticmethod
    def forward(ctx, x, weight_orig, mask, bias, operand):
        w = weight_orig * mask
        if operand == "bf16":
            x2, w2 = x.bfloat16().float(), w.bfloat16().float()
        else:
            x2, w2 = x, w
        ctx.save_for_backward(x2, w2, mask)
        ctx.operand, ctx.has_bias = operand, bias is not None
        y = x2.reshape(-1, x.shape[-1]) @ w2.t()
        if bias is not None:
            y = y + bias
        return y.view(*x.shape[:-1], w.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, w2, mask = ctx.saved_tensors
        d = dy.reshape(-1, dy.shape[-1])
        db = d.sum(0) if ctx.has_bias else None
        if ctx.operand == "bf16":
            d = d.bfloat16().float()
        dx = (d @ w2).view(x2.shape)
        dw = (d.t() @ x2.reshape(-1, x2.shape[-1])) * mask
        return dx, dw, None, db, None


def pruned_linear(x, weight_orig, mask, bias=None, operand="fp32"):
    return PrunedLinear.apply(x, weight_orig, mask, bias, operand)


def see_weight_rate(masks):
    """100 * zeros / elements over the given masks (run_vqa_stage3.py:75-178 walks the same set)."""
    total = sum(float(m.numel()) for m in masks.values())
    zeros = sum(float((m == 0).sum()) for m in masks.values())
    return 100 * zeros / total


def effective_params(params, masks):
    """state_dict-style params with '<module>.weight_orig' -> the dict lxmert_oracle.forward reads
    ('<module>.weight' = weight_orig * mask, an autograd node so gradients reach weight_orig)."""
    eff = {}
    for k, v in params.items():
        if k.endswith(".weight_orig"):
            mod = k[: -len(".weight_orig")]
            eff[mod + ".weight"] = v * masks[mod].to(v.dtype)
        else:
            eff[k] = v
    return eff


def forward(params, masks, batch):
    """(logits, pooled) of the pruned LXMERT in eval mode (fp32)."""
    c = lxo.Ctx(effective_params(params, masks), {}, {})
    return lxo.forward(c, batch["ids"], batch["feats"], batch["pos"])


def adam_step(p, g, m, v, step, lr=5e-5, b1=0.9, b2=0.999, eps=1e-8):
    """torch.optim.Adam (run_vqa_stage3.py:590): no weight decay, bias-corrected, eps added after the sqrt of
    the corrected second moment."""
    m.mul_(b1).add_(g, alpha=1 - b1)
    v.mul_(b2).addcmul_(g, g, value=1 - b2)
    bc1, bc2 = 1 - b1 ** step, 1 - b2 ** step
    denom = (v.sqrt() / (bc2 ** 0.5)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)
    return p
