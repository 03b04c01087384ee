"""The object that stands where the DeepSpeed engine stands in the reference's mPLUG training loop
(``model, optimizer, _, _ = deepspeed.initialize(...)``, mPLUG/vqa_mplug.py:394-401; used as ``model(...)``,
``model.backward(loss)``, ``model.step()``, ``model.global_steps`` at :171-204).

What DeepSpeed does there (mPLUG/configs/ds_config.json: bf16, ZeRO-2, gradient_clipping 1.0) and what this does:

* bf16 model copy + fp32 master weights  ->  scores stay fp32 Parameters (the master copy the optimiser updates) and
  the masked modules compare them exactly as a bf16 model copy would (``maskers.set_score_dtype``); the frozen
  weights are cast to bf16 once as GEMM operands.  Nothing else is cast: the few unmasked modules run in fp32.
* ZeRO-2 partitioning of gradients / optimiser state  ->  not needed: the whole state of mPLUG-base fits one B200
  (180 GB) many times over; every rank keeps a full replica (same choice as the LXMERT engine).
* gradient all-reduce  ->  one all-reduce(mean) of the flattened trainable gradients (scores + LM head; frozen
  weights have none) when ``torch.distributed`` is initialised with more than one rank.
* gradient_clipping  ->  global-norm clip over the (reduced) trainable gradients before the optimiser step.

Scores only change in ``step()``, thresholds only in ``reset_threshold``; so every masked module keeps its masked bf16
operand ``W (.) M`` between steps (one streaming pass per module and step) and forward / dX run as plain tcgen05 GEMMs
instead of re-deriving the mask inside every GEMM call.  Code that edits scores behind the engine's back must call
``invalidate_masks()``.
"""
import os

import torch
import torch.distributed as dist
from torch._utils import _flatten_dense_tensors, _unflatten_dense_tensors

if __package__:
    from .masking import maskers
else:
    from masking import maskers


class MaskTrainEngine:
    def __init__(self, module, optimizer, lr_scheduler=None, gradient_clipping=1.0, bf16=True, process_group=None,
                 hold_masks=True):
        self.module = module
        self.optimizer = optimizer
        self.lr_scheduler = lr_scheduler
        self.gradient_clipping = gradient_clipping
        self.process_group = process_group
        self.global_steps = 0
        self.last_grad_norm = None
        maskers.set_score_dtype(module, torch.bfloat16 if bf16 else torch.float32)
        # bf16=True is the reference's DeepSpeed configuration (mPLUG/configs/ds_config.json): there the WHOLE forward
        # runs in bf16.  The engine therefore switches the model to bf16 activations between the masked GEMMs (autocast;
        # LayerNorm / softmax / loss stay fp32) unless CRVQA_MPLUG_BF16_ACTIVATIONS=0 asks for fp32 activations.
        # Masks, thresholds and kept counts do not depend on the activation dtype (tests/test_zz_optin_gpu.py).
        if bf16 and hasattr(module, "bf16_activations") and os.environ.get("CRVQA_MPLUG_BF16_ACTIVATIONS", "1") != "0":
            module.bf16_activations = True
        self._masked = [m for m in module.modules() if hasattr(m, "hold_masked_weight")]
        for m in self._masked:
            m.hold_masked_weight(hold_masks)

    # -- the nn.Module face the loop uses -------------------------------------------------------
    def __call__(self, *args, **kwargs):
        return self.module(*args, **kwargs)

    def __getattr__(self, name):                   # named_modules(), named_parameters(), train(), eval(), ...
        return getattr(self.__dict__["module"], name)

    # -- the engine face ------------------------------------------------------------------------
    def backward(self, loss):
        loss.backward()

    def _trainable_grads(self):
        return [p.grad for p in self.module.parameters() if p.requires_grad and p.grad is not None]

    def _world(self):
        if dist.is_available() and dist.is_initialized():
            return dist.get_world_size(self.process_group)
        return 1

    def allreduce_gradients(self):
        """Mean over ranks of every trainable gradient, as ONE collective per dtype over a flat buffer."""
        world = self._world()
        if world == 1:
            return
        by_dtype = {}
        for g in self._trainable_grads():
            by_dtype.setdefault(g.dtype, []).append(g)
        for grads in by_dtype.values():
            flat = _flatten_dense_tensors(grads)
            dist.all_reduce(flat, group=self.process_group)
            flat.div_(world)
            for g, r in zip(grads, _unflatten_dense_tensors(flat, grads)):
                g.copy_(r)

    def step(self):
        self.allreduce_gradients()
        params = [p for p in self.module.parameters() if p.requires_grad and p.grad is not None]
        if self.gradient_clipping and params:
            self.last_grad_norm = torch.nn.utils.clip_grad_norm_(params, self.gradient_clipping)
        self.optimizer.step()
        self.optimizer.zero_grad(set_to_none=True)
        self.invalidate_masks()
        self.global_steps += 1

    def invalidate_masks(self):
        """The scores moved: the next forward of every masked module rebuilds its masked operand."""
        for m in self._masked:
            m.drop_masked_weight()
