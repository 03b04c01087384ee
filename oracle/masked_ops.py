"""Oracle restatement of the masking primitives (numpy / torch-CPU, fp32).

`operand` selects how GEMM operands are rounded before the fp32 multiply-accumulate:
  'fp32' -- exactly the reference arithmetic (F.linear on fp32 tensors);
  'bf16' -- X, W and dY rounded to bf16 first (what the tcgen05 path feeds its MMAs), products and
            accumulation still fp32.  This separates "bf16 operand rounding" from "kernel bug".
"""
import numpy as np
import torch
import torch.nn.functional as F


def binarize(scores, threshold):
    """binarizer_fn1 -- masking/maskers.py:325-329: out = clone; out[s <= thr] = 0; out[s > thr] = 1."""
    thr = float(threshold)
    out = scores.clone()
    out[scores <= thr] = 0.0
    out[scores > thr] = 1.0
    return out


def num_zero_elements(numel, rate):
    """k of every select: int(numel * rate) in Python double arithmetic; 0 -> 1 in reset_threshold
    (masking/maskers.py:201-202, hg_transformers/mask_trainer_Robust_VQA.py:476-479)."""
    return int(numel * rate)


def kth_value(values, k, use_abs=False):
    """k-th smallest (1-based) -- torch.kthvalue as called at masking/maskers.py:211 and
    mask_trainer_Robust_VQA.py:480.  Exact order statistic via numpy's introselect."""
    a = np.asarray(values.detach().cpu().numpy() if torch.is_tensor(values) else values, dtype=np.float32).reshape(-1)
    if use_abs:
        a = np.abs(a)
    if not 1 <= k <= a.size:
        raise ValueError("k out of range")
    return np.float32(np.partition(a, k - 1)[k - 1])


def global_kth_value(tensors, k, use_abs=False):
    """k-th smallest over the union of the tensors -- torch.cat([...]).kthvalue(k) as called at
    masking/global_maskers.py:536-541 (|W|) and hg_transformers/global_mask_trainer_VQA.py:424-429 (scores)."""
    flat = torch.cat([t.detach().reshape(-1) for t in tensors])
    return kth_value(flat, k, use_abs=use_abs)


def magnitude_init_global(weight, w_cut, threshold):
    """_magnitude_global -- masking/global_maskers.py:219-231: S = 2*thr where |W| > the global cut, else 0."""
    thr = float(threshold)
    s = torch.zeros_like(weight)
    keep = weight.abs() > float(w_cut)
    s[keep] = 2.0 * thr
    return s


def magnitude_init(weight, init_sparsity, threshold):
    """MaskedLinearX.controlled_init._magnitude -- masking/maskers.py:204-215:
    S = 2*thr where |W| > kthvalue(|W|, int(n*sparsity)) else 0*thr."""
    k = num_zero_elements(weight.numel(), init_sparsity)
    w_thr = kth_value(weight, k, use_abs=True)
    thr = float(threshold)
    s = torch.zeros_like(weight)
    keep = weight.abs() > float(w_thr)
    s[keep] = 2.0 * thr
    s[~keep] = 0.0 * thr
    return s, w_thr


def reset_threshold(scores, rate):
    """One module of Trainer.reset_threshold -- mask_trainer_Robust_VQA.py:476-480."""
    k = num_zero_elements(scores.numel(), rate)
    if k == 0:
        k = 1
    return kth_value(scores, k)


def _round(t, operand):
    return t.bfloat16().float() if operand == "bf16" else t


class MaskedLinear(torch.autograd.Function):
    """MaskedLinear1.forward and its autograd -- masking/maskers.py:337-339 (STE), :359-366 (forward),
    :564-569 (weights frozen => no dW, no db):
        Y  = X . (W (.) M)^T + b,   M = binarize(S, thr)
        dX = dY . (W (.) M)
        dS = (dY^T . X) (.) W
    """

    @staticmethod
    def forward(ctx, x, scores, weight, threshold, bias, operand):
        mask = binarize(scores.detach(), threshold)
        xr = _round(x.detach().reshape(-1, x.shape[-1]), operand)
        wr = _round(weight.detach(), operand)
        wm = wr * mask
        y = F.linear(xr, wm, bias)
        # operand == "bf16" models the MMA OPERANDS only (X, dY, W (.) M); the (.) W of dS multiplies by the fp32
        # weight, as the reference does and as the CUDA epilogue does
        ctx.save_for_backward(xr, weight.detach(), wm)
        ctx.operand = operand
        ctx.x_shape = x.shape
        return y.view(*x.shape[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        xr, wr, wm = ctx.saved_tensors
        dyr = _round(dy.reshape(-1, dy.shape[-1]), ctx.operand)
        dx = (dyr @ wm).view(ctx.x_shape) if ctx.needs_input_grad[0] else None
        ds = (dyr.t() @ xr) * wr
        return dx, ds, None, None, None, None


def masked_linear(x, scores, weight, threshold, bias, operand="fp32"):
    return MaskedLinear.apply(x, scores, weight, threshold, bias, operand)


def masked_linear_reference_form(x, scores, weight, threshold, bias):
    """The literal reference graph (binarizer STE -> weight * M -> F.linear) on torch autograd; used to
    confirm that MaskedLinear's hand-written backward equals what autograd derives."""

    class _STE(torch.autograd.Function):
        @staticmethod
        def forward(ctx, s):
            return binarize(s, threshold)

        @staticmethod
        def backward(ctx, g):
            return g

    return F.linear(x, weight * _STE.apply(scores), bias)


class MaskedEmbedding(torch.autograd.Function):
    """Embedding branch -- masking/maskers.py:362-363: F.embedding(ids, W * M, padding_idx);
    backward scatters dOut (.) W into the looked-up rows, rows == padding_idx get no gradient."""

    @staticmethod
    def forward(ctx, ids, scores, weight, threshold, padding_idx):
        mask = binarize(scores.detach(), threshold)
        ctx.save_for_backward(ids, weight)
        ctx.padding_idx = padding_idx
        return F.embedding(ids, weight * mask, padding_idx=padding_idx)

    @staticmethod
    def backward(ctx, dout):
        ids, weight = ctx.saved_tensors
        flat = ids.reshape(-1)
        d = dout.reshape(-1, dout.shape[-1]).clone()
        if ctx.padding_idx is not None:
            d[flat == ctx.padding_idx] = 0
        dwm = torch.zeros_like(weight).index_add_(0, flat, d)
        return None, dwm * weight, None, None, None


def masked_embedding(ids, scores, weight, threshold, padding_idx):
    return MaskedEmbedding.apply(ids, scores, weight, threshold, padding_idx)
