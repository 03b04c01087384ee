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
import sys
import types

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
        self._fused = None               # _FusedAdamW (or False), decided at the first step: torch AdamW on CUDA tensors
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
        # the fused layer kernels (dropout + residual + LayerNorm, few-query attention) hash their dropout masks from a
        # device-side (seed, counter): forward and backward of this step have both been queued, so move the counter on
        # -- without this every step would draw the SAME masks
        fused = sys.modules.get("crvqa.fused")       # loaded only if one of those kernels ran
        if fused is not None:
            fused.RngState.advance_all()

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
        if self._fused is None:
            self._fused = _FusedAdamW(self.optimizer, self._masked) if _FusedAdamW.wanted(self.optimizer) else False
        if self._fused and self._fused.step(self.gradient_clipping):
            # clip + AdamW + the refresh of every held masked operand ran as three launches (crv_sumsq_multi,
            # crv_adamw_multi per parameter group); only modules outside that pass have to rebuild their operand
            self.last_grad_norm = self._fused.grad_norm if self.gradient_clipping else None
            self.optimizer.zero_grad(set_to_none=True)
            for m in self._fused.not_refreshed:
                m.drop_masked_weight()
            self.global_steps += 1
            return
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


class _FusedAdamW:
    """Runs the step of a stock ``torch.optim.AdamW`` (what mPLUG/optim/optim_factory.py:60-89 builds and DeepSpeed
    steps for the reference) as multi-tensor launches over the optimiser's OWN tensors: the parameters stay where they
    are, ``optimizer.state`` holds the usual ``step`` / ``exp_avg`` / ``exp_avg_sq`` entries (state_dict round trips and
    a later ``optimizer.step()`` keep working), learning rates are read from the param groups every step (schedulers).

    Per step: crv_sumsq_multi (global gradient norm) -> one crv_adamw_multi per parameter group (clip coefficient folded
    in; the masked bf16 operand W (.) (S_new > T) of every held module is rewritten from the new scores in registers).
    step() returns False -- nothing done, the caller runs the PyTorch path -- whenever something is outside that
    contract (another optimiser class, amsgrad / maximize / capturable, CPU, non-fp32 or strided tensors, sparse
    gradients).  Parameters of one group whose step counts differ are stepped by separate launches."""

    device_types = ("cuda",)          # the kernels have no CPU implementation (tests/test_mplug_cpu.py swaps in an emulation)

    def __init__(self, optimizer, masked_modules):
        self.optimizer = optimizer
        by_param = {}
        for m in masked_modules:
            wm = getattr(m, "weight_mask", None)
            if wm is not None:
                by_param.setdefault(id(wm), []).append(m)
        self.by_param = by_param
        self.masked = list(masked_modules)
        self.plan = None
        self.grad_norm = None
        self.not_refreshed = self.masked
        # a plain optimizer.step() (ours never calls it) may advance the step counters of SOME tensors: re-plan after it
        self._stepped_outside = False
        optimizer.register_step_post_hook(self._note_outside_step)

    def _note_outside_step(self, *_):
        self._stepped_outside = True

    @staticmethod
    def wanted(optimizer):
        if os.environ.get("CRVQA_MPLUG_FUSED", "1") == "0" or type(optimizer) is not torch.optim.AdamW:
            return False
        for g in optimizer.param_groups:
            if g.get("amsgrad") or g.get("maximize") or g.get("capturable") or g.get("differentiable") or g.get("fused"):
                return False
        return any(p.device.type in _FusedAdamW.device_types for g in optimizer.param_groups for p in g["params"])

    # -- plan: tensor lists, address tables, row tables ------------------------------------------------------
    def _collect(self):
        """[(group, [params with a gradient])] or None when a tensor is outside the contract."""
        out, dev = [], None
        for group in self.optimizer.param_groups:
            ps = []
            for p in group["params"]:
                g = p.grad
                if g is None:
                    continue
                if (p.device.type not in self.device_types or p.dtype != torch.float32 or not p.is_contiguous()
                        or g.is_sparse or g.dtype != torch.float32 or not g.is_contiguous() or g.device != p.device
                        or (dev is not None and p.device != dev)):
                    return None
                dev = p.device
                ps.append(p)
            out.append((group, ps))
        return out if dev is not None else None

    def _state_of(self, p):
        st = self.optimizer.state[p]
        if len(st) == 0:                 # exactly what torch.optim.AdamW._init_group creates
            st["step"] = torch.tensor(0.0, dtype=torch.get_default_dtype())
            st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
        return st

    def _build(self, groups):
        from crvqa import ops
        # torch keeps one step counter per parameter (a tensor that starts receiving gradients later is behind the
        # others), the kernel takes one step number per launch: inside each group the parameters are ordered by their
        # count and every run of equal counts becomes one launch -- normally one run per group
        self._stepped_outside = False
        params, spans = [], []
        for gi, (_, ps) in enumerate(groups):
            for p in sorted(ps, key=lambda p: int(self._state_of(p)["step"])):
                t = int(self.optimizer.state[p]["step"])
                if not spans or spans[-1][0] != gi or spans[-1][3] != t:
                    spans.append([gi, len(params), len(params), t])
                params.append(p)
                spans[-1][2] = len(params)
        dev = params[0].device
        states = [self._state_of(p) for p in params]
        for st in states:
            for k in ("exp_avg", "exp_avg_sq"):
                t = st[k]
                if t.dtype != torch.float32 or not t.is_contiguous() or t.device != dev:
                    return None
        mods, thr, w16, wm = [], [], [], []
        for p in params:
            ms = self.by_param.get(id(p), ())
            m = ms[0] if len(ms) == 1 and ms[0].holds_masked_operand() else None
            mods.append(m)
            if m is None:
                thr.append(None), w16.append(None), wm.append(None)
                continue
            t = m._threshold_on(dev)
            w = m._weight_bf16()
            buf = m._wm if getattr(m, "_wm", None) is not None and m._wm.shape == w.shape else torch.empty_like(w)
            thr.append(t), w16.append(w), wm.append(buf)
        plan = types.SimpleNamespace(
            params=params, states=states, mods=mods, thr=thr, w16=w16, wm=wm, dev=dev,
            thr_src=[(m.threshold, m.threshold._version if torch.is_tensor(m.threshold) else None, m.score_dtype)
                     if m is not None else None for m in mods],
            moments=[(st["exp_avg"], st["exp_avg_sq"]) for st in states],
            group_ids=[tuple(id(p) for p in ps) for _, ps in groups], grad_ptrs=None)
        def ptr(t):
            return t.data_ptr() if t is not None else 0

        fixed = [[ptr(p) for p in params], [ptr(st["exp_avg"]) for st in states],
                 [ptr(st["exp_avg_sq"]) for st in states], [ptr(t) for t in w16], [ptr(t) for t in wm],
                 [ptr(t) for t in thr]]
        if any(a % 16 for row in fixed[:5] for a in row):          # the threshold scalars need no vector alignment
            return None
        plan.uniform_steps = len(spans) == sum(1 for _, ps in groups if ps)      # one launch per group
        plan.fixed_key = tuple(fixed[0])
        plan.tables = ops.upload(fixed, torch.int64, dev)          # [6, n]: p, m, v, w16, wm, thr
        plan.g_table = torch.empty(len(params), dtype=torch.int64, device=dev)
        rows = ops.multi_rows([p.numel() for p in params], [1 if m is not None else 0 for m in mods])
        plan.rows = ops.upload(rows, torch.int32, dev).contiguous()
        counts = [0] * len(params)
        for r in rows:
            counts[r[0]] += 1
        first_row = [0]
        for c in counts:
            first_row.append(first_row[-1] + c)
        plan.spans = [(gi, a, b, first_row[a], first_row[b]) for gi, a, b, _ in spans]   # group, params [a, b), rows
        plan.total = torch.zeros(1, dtype=torch.float32, device=dev)
        refreshed = {id(m) for m in mods if m is not None}
        self.not_refreshed = [m for m in self.masked if id(m) not in refreshed]
        return plan

    def _plan_is_current(self, plan, groups):
        if plan is None or self._stepped_outside or plan.group_ids != [tuple(id(p) for p in ps) for _, ps in groups]:
            return False              # (the plan keeps its parameters alive, so an id cannot have been reused)
        for i, p in enumerate(plan.params):
            if plan.fixed_key[i] != p.data_ptr():
                return False
            st = self.optimizer.state[p]              # load_state_dict() swaps in new dicts and tensors
            if st is not plan.states[i] or st.get("exp_avg") is not plan.moments[i][0] or \
                    st.get("exp_avg_sq") is not plan.moments[i][1]:
                return False
            m = plan.mods[i]
            if m is not None:
                t, ver, sd = plan.thr_src[i]
                cur = m.threshold
                if cur is not t or (ver is not None and cur._version != ver) or m.score_dtype is not sd \
                        or not m.holds_masked_operand() or m._w16 is not plan.w16[i]:
                    return False
        return True

    def step(self, max_norm):
        from crvqa import ops
        groups = self._collect()
        if groups is None:
            return False
        if not self._plan_is_current(self.plan, groups):
            self.plan = self._build(groups)
            if self.plan is None:
                return False
        plan = self.plan
        grad_ptrs = [p.grad.data_ptr() for p in plan.params]
        if grad_ptrs != plan.grad_ptrs:                         # fresh gradient tensors usually land where the last were
            if any(a % 16 for a in grad_ptrs):
                return False
            ops.upload(grad_ptrs, torch.int64, plan.dev, out=plan.g_table)
            plan.grad_ptrs = grad_ptrs
        tp, tm, tv, tw, twm, tthr = plan.tables.unbind(0)
        total = None
        if max_norm:
            plan.total.zero_()
            ops.sumsq_multi(plan.g_table, plan.rows, plan.total)
            total = plan.total
            self.grad_norm = plan.total.sqrt().reshape(())
        for gi, a, _, r0, r1 in plan.spans:
            group = groups[gi][0]
            b1, b2 = group["betas"]
            ops.adamw_multi(tp, plan.g_table, tm, tv, tw, twm, tthr, plan.rows[r0:r1], group["lr"],
                            int(plan.states[a]["step"]) + 1, b1, b2, group["eps"], group["weight_decay"], total,
                            max_norm or 1.0)
        torch._foreach_add_([st["step"] for st in plan.states], 1)
        for m, wm, thr in zip(plan.mods, plan.wm, plan.thr):
            if m is not None and m._wm is not wm:
                m.adopt_masked_weight(wm, thr)
        return True
