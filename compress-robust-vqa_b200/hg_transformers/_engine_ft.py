"""B200 engine of the stage-3 frozen-mask fine-tune (BASELINE config 4).

The reference fine-tunes EVERY tensor of the pruned network with torch.optim.Adam over ~500 one-tensor parameter groups
(run_vqa_stage3.py:577-598) after reparametrising the masked modules with torch.nn.utils.prune (:227-297).  Here the
whole trainable state lives in flat buffers, as the stage-2 scores do (hg_transformers._engine.ScoreArena):

WeightArena
    P   fp32  every trainable tensor: first the GEMM weights (``weight_orig`` of the pruned Linears, ``weight`` of
              unpruned ones) in execution order, then everything else (biases, LayerNorm, embeddings, box_fc, the
              answer head) in named_parameters order.  The Parameters are views.
    G   fp32  their gradients, same offsets.  dW of a GEMM module is written by the dS-GEMM epilogue
              ((dY^T X) (.) M with the fp32 0/1 mask as the multiplier), bias gradients by deterministic column sums
              of dY, LayerNorm gradients by the LayerNorm backward kernel; autograd accumulates the rest in place.
    M32 fp32  0/1 masks of the GEMM region (the modules' ``weight_mask`` buffers are views; ones for unpruned modules)
    M16 bf16  the same masks as bf16 (read by the optimiser pass)
    Wm  bf16  the GEMM operands bf16(P (.) M): rewritten by the optimiser pass itself (the new weights are in
              registers there), read by every forward / dX GEMM
    exp_avg, exp_avg_sq   fp32, same offsets

One optimiser step = crv_sumsq (global-norm clip) + ONE crv_adamw_segmented launch in torch.optim.Adam mode over P
(clip + Adam + operand refresh + gradient clearing).  The layer kernels of the stage-2 fast path (crvqa.fused) run
unchanged on top; GraphedStep captures the whole step.
"""
import os

import torch
from torch import nn

from crvqa import ops

from ._engine import _ALIGN, execution_order


def _is_pruned(m):
    return hasattr(m, "weight_orig") and hasattr(m, "weight_mask")


class WeightArena:
    cache_on = True

    def __init__(self, model, device=None):
        named = []
        for name, m in model.named_modules():
            if _is_pruned(m) and m.weight_orig.dim() == 2 and hasattr(m, "in_features"):
                if m.in_features % 8 == 0 and m.weight_orig.requires_grad:
                    named.append((name, m))
            elif (isinstance(m, nn.Linear) and "weight" in m._parameters and m.weight.requires_grad
                  and m.in_features % 8 == 0 and ".encoder." in "." + name):
                named.append((name, m))          # FT_randMask leaves r_layers / x_layers unpruned: mask of ones
        if not named:
            raise ValueError("no GEMM modules to place in the arena")
        named = execution_order(named)
        self.names = [n for n, _ in named]
        self.modules = [m for _, m in named]
        gemm_params = [self._wparam(m) for m in self.modules]
        device = device or gemm_params[0].device
        seen = {id(p) for p in gemm_params}
        self.loose = [(n, p) for n, p in model.named_parameters() if p.requires_grad and id(p) not in seen]
        self.offsets, off = [], 0
        for p in gemm_params:
            self.offsets.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.total = off                       # end of the GEMM region (what GradSync buckets)
        self.loose_offsets = []
        for _, p in self.loose:
            self.loose_offsets.append(off)
            off += (p.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.size = off
        f32 = dict(dtype=torch.float32, device=device)
        self.params = torch.zeros(self.size, **f32)
        self.grads = torch.zeros(self.size, **f32)
        self.exp_avg = torch.zeros(self.size, **f32)
        self.exp_avg_sq = torch.zeros(self.size, **f32)
        self.w32 = torch.zeros(self.total, **f32)                                # fp32 masks: the dS multiplier
        self.w16 = torch.zeros(self.total, dtype=torch.bfloat16, device=device)  # bf16 masks
        self.wm = torch.zeros(self.total, dtype=torch.bfloat16, device=device)
        self.scores = self.params                # the name the shared engine code reads (GradSync, GraphedStep)
        self._index, self._pindex = {}, {}
        for i, (m, p) in enumerate(zip(self.modules, gemm_params)):
            o, n = self.offsets[i], p.numel()
            self._adopt(p, o)
            mask = self.w32[o: o + n].view(p.shape)
            if _is_pruned(m):
                mask.copy_(m.weight_mask)
                m.weight_mask = mask             # registered buffer: stays in state_dict, now a view
            else:
                mask.fill_(1.0)
            m._arena, m._arena_grad = self, p.grad
            m._arena_index = i
            m._grad_dirty, m._grad_zero = False, True
            m._wm = self.wm[o: o + n].view(p.shape)
            self._index[id(m)] = i
            self._pindex[id(p)] = ("gemm", i)
        for j, ((_, p), o) in enumerate(zip(self.loose, self.loose_offsets)):
            self._adopt(p, o)
            self._pindex[id(p)] = ("loose", j)
        self._loose_off = {id(p): o for (_, p), o in zip(self.loose, self.loose_offsets)}
        # LayerNorm: when beta directly follows gamma in G, the backward kernel's second stage adds [dgamma | dbeta]
        # straight into the arena (crvqa.fused.DropAddLayerNormFn), no autograd accumulation kernels
        for m in model.modules():
            if isinstance(m, nn.LayerNorm) and m.weight is not None and m.bias is not None:
                ow, ob = self._loose_off.get(id(m.weight)), self._loose_off.get(id(m.bias))
                if ow is not None and ob == ow + m.weight.numel():
                    m.weight._arena_pair = self.grads[ow: ow + 2 * m.weight.numel()]
        self.keep_grads = os.environ.get("CRVQA_KEEP_GRADS", "0") == "1"
        self.shard = None
        self.epoch = 0
        rows = []
        for i, p in enumerate(gemm_params):
            rows += self._rows(self.offsets[i], p.numel(), 1)
        for (_, p), o in zip(self.loose, self.loose_offsets):
            rows += self._rows(o, p.numel(), 0)
        self.step_chunks = torch.tensor(rows, dtype=torch.int32, device=device).contiguous()
        self.refresh_operands()

    @staticmethod
    def _wparam(m):
        return m.weight_orig if _is_pruned(m) else m.weight

    @staticmethod
    def _rows(off, n, flag):
        return [((off + c0) // 8, min(8192, n - c0), 0, flag) for c0 in range(0, n, 8192)]

    def _adopt(self, p, off):
        n = p.numel()
        view = self.params[off: off + n].view(p.shape)
        view.copy_(p.data)
        p.data = view
        p.grad = self.grads[off: off + n].view(p.shape)
        p._arena_grad = p.grad

    # -- protocol shared with ScoreArena (crvqa.fused.ProjectionGroup, GradSync) ---------------------------------
    def owns(self, p):
        return id(p) in self._pindex

    def index_of(self, m):
        return self._index[id(m)]

    def weight_shape(self, m):
        return tuple(self._wparam(m).shape)

    def anchor(self, m):
        return self._wparam(m)

    def _gemm_module(self, m):
        return True

    def group_bias(self, modules):
        """(bias, bias gradient) of adjacent modules as contiguous live views of P / G, or (None, None)."""
        bs = [m.bias for m in modules]
        if any(b is None for b in bs):
            return None, None
        offs = [self._loose_off.get(id(b)) for b in bs]
        if any(o is None for o in offs):
            raise ValueError("bias outside the arena")
        for b, o, nxt in zip(bs, offs, offs[1:]):
            if o + b.numel() != nxt:
                raise ValueError("biases are not adjacent in the arena")
        n = sum(b.numel() for b in bs)
        return self.params[offs[0]: offs[0] + n], self.grads[offs[0]: offs[0] + n]

    def _current(self, m):
        p = self._wparam(m)
        mv = m.weight_mask._version if _is_pruned(m) else 0
        return (p._version, mv, p.data_ptr())

    def cached_masked_weight(self, m):
        """bf16(W (.) M) of module m, rebuilt when someone changed the weight or the mask outside the optimiser pass
        (load_state_dict, a second pruning round)."""
        if getattr(m, "_wm_key", None) != self._current(m):
            i = self._index[id(m)]
            o, n = self.offsets[i], self._wparam(m).numel()
            if _is_pruned(m) and m.weight_mask.data_ptr() != self.w32[o: o + n].data_ptr():
                self.w32[o: o + n].view(m.weight_mask.shape).copy_(m.weight_mask)
                m.weight_mask = self.w32[o: o + n].view(m.weight_mask.shape)
            self.w16[o: o + n].copy_(self.w32[o: o + n])
            ops.mul_cast_bf16(self.params[o: o + n], self.w32[o: o + n], out=self.wm[o: o + n])
            m._wm_key = self._current(m)
        return m._wm

    def module_ready(self, m):
        return id(m) in self._index and self._wparam(m).is_cuda

    def wait_ready(self, m):
        pass

    def refresh_operands(self):
        """M16 and Wm of the whole GEMM region from P and M32 (two launches)."""
        self.w16.copy_(self.w32)
        ops.mul_cast_bf16(self.params[: self.total], self.w32, out=self.wm)
        for m in self.modules:
            m._wm_key = self._current(m)

    def state_views(self, p):
        kind = self._pindex.get(id(p))
        if kind is None:
            return None
        o = self.offsets[kind[1]] if kind[0] == "gemm" else self.loose_offsets[kind[1]]
        n = p.numel()
        return {"exp_avg": self.exp_avg[o: o + n].view(p.shape), "exp_avg_sq": self.exp_avg_sq[o: o + n].view(p.shape)}

    def loose_grads(self):
        """The non-GEMM part of G as one tensor (data parallel: all-reduced as one flat message after backward)."""
        return [self.grads[self.total: self.size]]

    # -- per-step protocol ---------------------------------------------------------------------------------------
    def begin_step(self):
        """zero_grad: nothing to do -- the optimiser pass cleared G -- except to re-attach gradients someone set to
        None and, with CRVQA_KEEP_GRADS=1 (tests that read G after a full step), to clear G here."""
        if self.keep_grads:
            self.grads.zero_()
        for m in self.modules:
            m._grad_dirty = False
            if self.keep_grads:
                m._grad_zero = True
        for m, p in zip(self.modules, (self._wparam(m) for m in self.modules)):
            if p.grad is None:
                p.grad = p._arena_grad
        for _, p in self.loose:
            if p.grad is None:
                p.grad = p._arena_grad

    def finalize_grads(self):
        ops.ds_lane_join()
        for m in self.modules:
            if not m._grad_dirty:
                if not getattr(m, "_grad_zero", False):
                    m._arena_grad.zero_()
                m._grad_dirty = True

    def grad_sumsq_into(self, acc):
        ops.sumsq_into(self.grads, acc)

    def adam_step(self, lr, step, beta1, beta2, eps, weight_decay, clip_sumsq, max_norm, hyper=None):
        """clip + torch.optim.Adam + bf16 operand refresh + gradient clearing over the whole arena: ONE launch."""
        zero = not self.keep_grads
        ops.adamw_segmented(self.params, self.grads, self.exp_avg, self.exp_avg_sq, None, self.step_chunks, None,
                            self.w16, self.wm, lr, step, beta1, beta2, eps, weight_decay, clip_sumsq, max_norm, True,
                            hyper, zero, mode=ops.ADAM_TORCH)
        for m in self.modules:
            if zero:
                m._grad_zero = True
        self.epoch += 1

    def release(self):
        for m in self.modules:
            p = self._wparam(m)
            p.data = p.data.clone()
            p.grad = None
            if _is_pruned(m):
                m.weight_mask = m.weight_mask.clone()
            m._arena = m._arena_grad = None
        for _, p in self.loose:
            p.data = p.data.clone()
            p.grad = None
