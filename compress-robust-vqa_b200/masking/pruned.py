"""Stage-3 frozen-mask modules.

The reference fine-tunes the sub-network found in stage 2 by reparametrising every masked module with
``torch.nn.utils.prune`` (run_vqa_stage3.py:205-297): ``weight_orig`` (trainable Parameter), ``weight_mask``
(fp32 0/1 buffer) and ``weight = weight_orig * weight_mask`` recomputed by a forward pre-hook on every call.
These modules keep exactly that state (same parameter / buffer names, so ``state_dict()`` keys and
``see_weight_rate`` are unchanged) and run the hot arithmetic through the C ABI:

    forward   bf16(weight_orig * weight_mask) in ONE pass (crv_mul_cast_bf16), then the tcgen05 GEMM
    dX        dY . (W (.) M)                           (crv_masked_linear_bwd_dx)
    dW_orig   (dY^T . X) (.) M                         (crv_masked_linear_bwd_ds with the fp32 mask as the multiplier)
    db        column sums of dY

No CPU fallback: on a CPU tensor the ops raise (crvqa.ops._need_cuda).
"""
import torch
import torch.nn.functional as F
from torch import nn

from crvqa import ops


class PrunedLinearFn(torch.autograd.Function):
    """`wm` / `sink`: under the stage-3 engine (hg_transformers._engine_ft.WeightArena) the bf16 operand is the arena's
    (kept current by the optimiser pass) and dW_orig is written into the arena gradient by the GEMM epilogue."""

    @staticmethod
    def forward(ctx, x, weight_orig, mask, bias, wm=None, sink=None):
        shp = x.shape
        x2 = ops.to_bf16(x.reshape(-1, shp[-1]))
        if wm is None:
            wm = ops.mul_cast_bf16(weight_orig, mask)
        y = ops.masked_linear_fwd(x2, wm, None, None, bias, torch.float32)
        ctx.save_for_backward(x2, wm, mask)
        ctx.x_shape = shp
        ctx.has_bias = bias is not None
        ctx.sink = sink
        return y.view(*shp[:-1], wm.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, wm, mask = ctx.saved_tensors
        d = dy.reshape(-1, dy.shape[-1])
        dy2 = ops.to_bf16(d)
        dx = dw = db = None
        if ctx.needs_input_grad[0]:
            dx = ops.masked_linear_bwd_dx(dy2, wm, None, None, torch.float32).view(ctx.x_shape)
        if ctx.needs_input_grad[1]:
            sink = ctx.sink
            if sink is not None and ops._sink_grad(sink) is not None:
                ops.masked_linear_bwd_ds(dy2, x2, mask, out=ops._sink_grad(sink), accumulate=ops.ds_mode(sink))
                ops._sink_done(sink)
            else:
                dw = ops.masked_linear_bwd_ds(dy2, x2, mask)    # (dY^T X) (.) M: the multiplier is the 0/1 mask
        if ctx.has_bias and ctx.needs_input_grad[3]:
            db = d.sum(0)
        return dx, dw, None, db, None, None


class PrunedLinear(nn.Module):
    """nn.Linear under prune.CustomFromMask / prune.l1_unstructured."""

    def __init__(self, weight_orig, weight_mask, bias):
        super().__init__()
        self.weight_orig = weight_orig                       # the module's former `weight` Parameter object
        self.register_buffer("weight_mask", weight_mask.to(dtype=weight_orig.dtype, device=weight_orig.device))
        self.bias = bias
        self.out_features, self.in_features = weight_orig.shape

    @property
    def weight(self):
        return self.weight_orig * self.weight_mask

    def forward(self, x):
        if self.in_features % 8 != 0:
            # box_fc (K = 4) cannot be a TMA operand; 3 K-element rows of torch math on the device
            return F.linear(x, self.weight_orig * self.weight_mask, self.bias)
        arena = getattr(self, "_arena", None)
        wm = sink = None
        if arena is not None and x.is_cuda:
            wm, sink = arena.cached_masked_weight(self), self
            if torch.is_grad_enabled() and getattr(self, "_sync", None) is not None:
                self._sync.note_forward(self)
        return PrunedLinearFn.apply(x, self.weight_orig, self.weight_mask.contiguous(), self.bias, wm, sink)

    def extra_repr(self):
        return f"in_features={self.in_features}, out_features={self.out_features}, bias={self.bias is not None}"


class PrunedEmbedding(nn.Module):
    """nn.Embedding under the same reparametrisation.  Only the gathered rows are masked (the reference
    multiplies the whole 30522 x 768 table on every forward); the gradient of weight_orig is the scatter of
    dOut (.) M[id], which is what autograd derives from the reference graph."""

    def __init__(self, weight_orig, weight_mask, padding_idx):
        super().__init__()
        self.weight_orig = weight_orig
        self.register_buffer("weight_mask", weight_mask.to(dtype=weight_orig.dtype, device=weight_orig.device))
        self.padding_idx = padding_idx
        self.num_embeddings, self.embedding_dim = weight_orig.shape

    @property
    def weight(self):
        return self.weight_orig * self.weight_mask

    def forward(self, ids):
        return F.embedding(ids, self.weight_orig, padding_idx=self.padding_idx) * F.embedding(ids, self.weight_mask)


def _replace(root, dotted, new):
    parent = root
    parts = dotted.split(".")
    for p in parts[:-1]:
        parent = getattr(parent, p)
    setattr(parent, parts[-1], new)


def custom_from_mask(root, dotted, mask):
    """prune.CustomFromMask.apply(module, 'weight', mask) on root.<dotted>: the module is replaced by its pruned
    form, which owns the SAME Parameter objects (so optimisers built before or after see identical tensors)."""
    mod = root
    for p in dotted.split("."):
        mod = getattr(mod, p)
    mask = torch.as_tensor(mask)
    if isinstance(mod, (PrunedLinear, PrunedEmbedding)):     # pruning twice multiplies the masks (PruningContainer)
        mod.weight_mask.mul_(mask.to(mod.weight_mask))
        return mod
    if tuple(mask.shape) != tuple(mod.weight.shape):
        raise ValueError(f"mask shape {tuple(mask.shape)} != weight shape {tuple(mod.weight.shape)} for {dotted}")
    if isinstance(mod, nn.Embedding):
        new = PrunedEmbedding(mod.weight, mask, mod.padding_idx)
    elif isinstance(mod, nn.Linear):
        new = PrunedLinear(mod.weight, mask, mod.bias)
    else:
        raise TypeError(f"cannot prune {type(mod).__name__} at {dotted}")
    _replace(root, dotted, new)
    return new


def l1_unstructured_mask(weight, amount):
    """Mask of prune.l1_unstructured(module, 'weight', amount): exactly k = round(amount * n) entries of smallest
    |w| are zeroed.  The cut is the k-th smallest |w| from the exact batched select (crv_kth_value_batched on
    |w|); everything strictly below it is dropped, and of the entries EQUAL to it the lowest indices are dropped
    until k are gone.  torch.topk leaves the choice among tied magnitudes implementation-defined, so with a tie
    at the cut the dropped entries may differ from torch's inside the tie group; the count never does."""
    n = weight.numel()
    k = round(amount * n)
    if k <= 0:
        return torch.ones_like(weight)
    a = weight.detach().abs().reshape(-1)
    thr = ops.kth_value_batched([a], [k])[0]
    below = a < thr
    tied = a == thr
    need = k - below.sum()
    drop = below | (tied & (tied.cumsum(0) <= need))
    return (~drop).to(weight.dtype).view_as(weight)
