"""Engine-side fused transformer-layer ops (SURVEY.md section 8 "next" row f3).

When the masked modules of a layer live in a ScoreArena with a valid mask cache (the training engine of
hg_transformers._trainer_core), the layer is evaluated as

    bf16 activations --> [grouped masked GEMM: Q|K|V in one launch] --> SDPA --> [masked GEMM] -->
    [dropout + residual + LayerNorm, one kernel, fp32 residual + bf16 operand copy] -->
    [masked GEMM] --> [GELU bf16] --> [masked GEMM] --> [dropout + residual + LayerNorm]

instead of the reference's fp32 op-by-op chain (hg_transformers/modeling_lxmert.py:770-903).  The math is
the same; every GEMM operand is bf16 as before; the residual stream stays fp32.  Modules outside an arena
(plain drop-in use) keep the generic per-module path of masking._core.MaskedLinear1.
"""
import ctypes
import math
import os

import torch
import torch.nn.functional as F

from . import ops
from ._lib import check, lib

_p, _stream = ops._p, ops._stream


# ----------------------------------------------------------------------------- dropout RNG state
class RngState:
    """(seed, step counter) in device memory: kernels hash (seed, counter, site, element) -> keep bit, so a
    captured CUDA graph draws new masks on every replay once `advance()` is part of the graph."""
    _per_device = {}
    _next_site = [1]

    def __init__(self, device, seed):
        self.state = torch.tensor([seed, 0], dtype=torch.int64, device=device)

    @classmethod
    def get(cls, device):
        key = (device.type, device.index)
        st = cls._per_device.get(key)
        if st is None:
            st = cls._per_device[key] = RngState(device, int(torch.initial_seed()) & 0x7FFFFFFFFFFFFFFF)
        return st

    @classmethod
    def new_site(cls):
        cls._next_site[0] += 1
        return cls._next_site[0]

    @classmethod
    def new_sites(cls, n):
        """First of n consecutive call-site ids."""
        first = cls._next_site[0] + 1
        cls._next_site[0] += n
        return first

    def advance(self):
        check(lib.crv_rng_advance(_p(self.state), _stream()), "crv_rng_advance")

    @classmethod
    def advance_all(cls):
        """Move the counter of every device that has drawn masks (once per training step, after its backward)."""
        for st in cls._per_device.values():
            st.advance()


# ----------------------------------------------------------------------------- grouped masked linear
class ProjectionGroup:
    """Adjacent arena modules sharing one input (query|key|value, or key|value): their masked bf16 weights,
    bf16 weights and score gradients are contiguous [sum N, K] slabs of the arena, so one GEMM serves all."""

    def __init__(self, modules):
        self.modules = list(modules)
        arena = self.modules[0]._arena
        idx = [arena.index_of(m) for m in self.modules]
        shapes = [arena.weight_shape(m) for m in self.modules]
        off0 = arena.offsets[idx[0]]
        off = off0
        for i, shp in zip(idx, shapes):
            if arena.offsets[i] != off or shp[1] != shapes[0][1]:
                raise ValueError("modules are not adjacent in the arena")
            off += shp[0] * shp[1]
        self.arena = arena
        self.K = shapes[0][1]
        self.N = sum(shp[0] for shp in shapes)
        n = self.N * self.K
        self.wm = arena.wm[off0: off0 + n].view(self.N, self.K)
        self.w16 = arena.w16[off0: off0 + n].view(self.N, self.K)
        self.w32 = arena.w32[off0: off0 + n].view(self.N, self.K)      # the dS multiplier: fp32 W (stage 2) / fp32 mask
        self.grad = arena.grads[off0: off0 + n].view(self.N, self.K)
        # stage 2: frozen biases (a private copy); stage 3: live views of the trained biases and of their gradients
        self.bias, self.bias_grad = arena.group_bias(self.modules)
        self.anchor = arena.anchor(self.modules[0])  # keeps the autograd node alive even if x needs no gradient

    def valid(self):
        a = self.arena
        return all(a.cached_masked_weight(m) is not None for m in self.modules)

    def note_forward(self):
        self.arena.wait_ready(self.modules[-1])       # sharded optimiser: this bucket's operand all-gather has landed
        if torch.is_grad_enabled():
            for m in self.modules:
                if getattr(m, "_sync", None) is not None:
                    m._calls_outstanding = getattr(m, "_calls_outstanding", 0) + 1


class GroupLinearFn(torch.autograd.Function):
    """bf16 in -> (bf16 | fp32) out masked linear over a ProjectionGroup, reading the cached W (.) M."""

    @staticmethod
    def forward(ctx, x16, anchor, group, out_dtype):
        shp = x16.shape
        x2 = x16.reshape(-1, shp[-1])
        if not x2.is_contiguous():
            x2 = x2.contiguous()
        y = ops.masked_linear_fwd(x2, group.wm, None, None, group.bias, out_dtype)
        ctx.save_for_backward(x2)
        ctx.group, ctx.x_shape, ctx.need_dx = group, shp, x16.requires_grad
        return y.view(*shp[:-1], group.N)

    @staticmethod
    def backward(ctx, dy):
        (x2,) = ctx.saved_tensors
        g = ctx.group
        dy2 = dy.reshape(-1, g.N)
        dy2 = ops.to_bf16(dy2) if dy2.dtype != torch.bfloat16 else (dy2 if dy2.is_contiguous() else dy2.contiguous())
        lane = ops.ds_lane(dy2.device) if ctx.need_dx else None
        if lane is not None:
            lane.fork()                      # dY is ready here; dS need not wait for dX (see ops._DsLane)
        dx = None
        if ctx.need_dx:
            dx = ops.masked_linear_bwd_dx(dy2, g.wm, None, None, torch.bfloat16).view(ctx.x_shape)
        dirty = ops.ds_mode(*g.modules)
        if lane is not None:
            with torch.cuda.stream(lane.stream):
                ops.masked_linear_bwd_ds(dy2, x2, g.w32, out=g.grad, accumulate=dirty)
                if g.bias_grad is not None:
                    ops.colsum_bf16(dy2, g.bias_grad, accumulate=True)
            lane.hold(dy2, x2)
        else:
            ops.masked_linear_bwd_ds(dy2, x2, g.w32, out=g.grad, accumulate=dirty)
            if g.bias_grad is not None:      # stage 3: db += column sums of dY (the optimiser pass cleared G)
                ops.colsum_bf16(dy2, g.bias_grad, accumulate=True)
        for m in g.modules:
            ops._sink_done(m)
        return dx, None, None, None


def group_linear(group, x16, out_dtype=torch.bfloat16):
    group.note_forward()
    return GroupLinearFn.apply(x16, group.anchor, group, out_dtype)


# ----------------------------------------------------------------------------- grouped launches
def _grouped_on():
    return os.environ.get("CRVQA_GROUPED", "1") != "0"


class MultiLinearFn(torch.autograd.Function):
    """n masked linears over ProjectionGroups as ONE grouped 2-CTA launch (crv_masked_gemm_grouped); their backward is
    one grouped launch list too: dX and dS of every member read the same dY.

    args = x_1 .. x_n (bf16), anchor_1 .. anchor_n (score Parameters: keep the node alive), then the pre-activations
    u_i of the members whose spec carries `u_in` (in member order).
    spec = (group, out_dtype, gelu_out, u_in):
      gelu_out  the forward returns (gelu(y), y): the FFN intermediate with its activation fused into the epilogue
      u_in      this member's INPUT is gelu(u) of the tensor u given in args; its dX epilogue multiplies by gelu'(u),
                i.e. the gradient returned for x is ALREADY the gradient of u.  Private contract of FfnPlan: the
                producer of x (a gelu_out member) passes its incoming gradient through unchanged."""

    @staticmethod
    def forward(ctx, specs, *args):
        n = len(specs)
        xs, extra = args[:n], list(args[2 * n:])
        x2s, outs, problems, ret, u_ins = [], [], [], [], []
        for (group, out_dtype, gelu_out, has_u), x in zip(specs, xs):
            x2 = x.reshape(-1, x.shape[-1])
            x2 = x2 if x2.is_contiguous() else x2.contiguous()
            M = x2.shape[0]
            y = torch.empty((M, group.N), dtype=out_dtype, device=x2.device)
            u = torch.empty((M, group.N), dtype=torch.bfloat16, device=x2.device) if gelu_out else None
            problems.append(ops.gemm_problem(ops.GEMM_FWD, x2, group.wm, y, bias=group.bias, aux=u,
                                             act=ops.ACT_GELU if gelu_out else ops.ACT_NONE))
            x2s.append(x2)
            u_ins.append(extra.pop(0) if has_u else None)
            ret.append(y.view(*x.shape[:-1], group.N))
            if gelu_out:
                ret.append(u.view(*x.shape[:-1], group.N))
                outs.append(ret[-1])
        ops.gemm_grouped(problems)
        ctx.set_materialize_grads(False)     # a member nothing downstream reads gets None, and its GEMMs are skipped
        ctx.specs = specs
        ctx.shapes = [x.shape for x in xs]
        ctx.need_dx = [x.requires_grad for x in xs]
        ctx.n_extra = len(args) - 2 * n
        ctx.save_for_backward(*x2s, *[u for u in u_ins if u is not None])
        ctx.has_u = [u is not None for u in u_ins]
        if outs:
            ctx.mark_non_differentiable(*outs)
        return tuple(ret)

    @staticmethod
    def backward(ctx, *dys):
        specs, n = ctx.specs, len(ctx.specs)
        saved = list(ctx.saved_tensors)
        x2s, us = saved[:n], saved[n:]
        p_dx, p_ds, dxs, seen, k, ui, held, colsums = [], [], [], set(), 0, 0, [], []
        for i, (group, out_dtype, gelu_out, has_u) in enumerate(specs):
            dy = dys[k]
            k += 2 if gelu_out else 1
            if has_u:
                ui += 1
            if dy is None:         # e.g. the vision side of the last cross layer: no gradient, no work
                dxs.append(None)
                continue
            dy2 = dy.reshape(-1, group.N)
            dy2 = ops.to_bf16(dy2) if dy2.dtype != torch.bfloat16 else (dy2 if dy2.is_contiguous() else dy2.contiguous())
            u = us[ui - 1].reshape(-1, group.K) if has_u else None
            dx = None
            if ctx.need_dx[i]:
                dx = torch.empty((dy2.shape[0], group.K), dtype=torch.bfloat16, device=dy2.device)
                p_dx.append(ops.gemm_problem(ops.GEMM_DX, dy2, group.wm, dx, aux=u,
                                             act=ops.ACT_GELU if u is not None else ops.ACT_NONE))
            dxs.append(dx)
            key = group.grad.data_ptr()
            acc = ops.DS_ADD if key in seen else ops.ds_mode(*group.modules)   # second use of a shared module adds
            seen.add(key)
            p_ds.append(ops.gemm_problem(ops.GEMM_DS, dy2, x2s[i], group.grad, w_f32=group.w32, accumulate=acc))
            held += [dy2, x2s[i]]
            if group.bias_grad is not None:
                colsums.append((dy2, group.bias_grad))
        # dX feeds the next layer's backward, dS only the end of the step.  Either everything shares one grouped launch
        # list, or (CRVQA_GROUP_DS_LANE=1) the dS group runs on the dS lane beside the dX chain and the small kernels
        # between the GEMMs (ops._DsLane).
        if not p_ds:
            for group, _, _, _ in specs:
                for m in group.modules:
                    ops._sink_skipped(m)
            return (None,) * (1 + 2 * n + ctx.n_extra)
        lane = (ops.ds_lane(x2s[0].device)
                if (p_dx and os.environ.get("CRVQA_GROUP_DS_LANE", "0") == "1") else None)
        if lane is not None:
            lane.fork()
            ops.gemm_grouped(p_dx)
            with torch.cuda.stream(lane.stream):
                ops.gemm_grouped(p_ds)
            lane.hold(*held)
        else:
            ops.gemm_grouped(p_dx + p_ds)
        for dy2, bg in colsums:              # stage 3: bias gradients (G is cleared by the optimiser pass)
            ops.colsum_bf16(dy2, bg, accumulate=True)
        k = 0
        for group, _, gelu_out, _ in specs:
            live = dys[k] is not None
            k += 2 if gelu_out else 1
            for m in group.modules:
                (ops._sink_done if live else ops._sink_skipped)(m)
        grads = [dx.view(shp) if dx is not None else None for dx, shp in zip(dxs, ctx.shapes)]
        return (None, *grads, *([None] * n), *([None] * ctx.n_extra))


def multi_linear(items):
    """items: (group, x16, out_dtype, gelu_out, u_in).  Returns one entry per item: y, or (gelu(y), y) for gelu_out."""
    specs = [(g, dt, bool(gelu), u is not None) for g, _, dt, gelu, u in items]
    for g, *_ in items:
        g.note_forward()
    args = [x for _, x, _, _, _ in items] + [g.anchor for g, *_ in items] + [u for *_, u in items if u is not None]
    flat = MultiLinearFn.apply(specs, *args)
    out, k = [], 0
    for _, _, _, gelu, _ in items:
        if gelu:
            out.append((flat[k], flat[k + 1]))
            k += 2
        else:
            out.append(flat[k])
            k += 1
    return out


# ----------------------------------------------------------------------------- dropout + residual + LayerNorm
class DropAddLayerNormFn(torch.autograd.Function):
    """(y fp32, y bf16) = LayerNorm(dropout(g) + res).  gamma / beta are frozen in stage 2; when they require a
    gradient (stage-3 fine-tune) the backward kernel also emits per-CTA partial sums of dgamma / dbeta, which a second
    launch adds up in index order."""

    @staticmethod
    def forward(ctx, g, res32, gamma, beta, eps, p, site, rng):
        H = g.shape[-1]
        g2 = g.reshape(-1, H)
        g2 = g2 if g2.is_contiguous() else g2.contiguous()
        M = g2.shape[0]
        r2 = None
        if res32 is not None:
            r2 = res32.reshape(-1, H)
            r2 = r2 if r2.is_contiguous() else r2.contiguous()
        dev = g.device
        y32 = torch.empty((M, H), dtype=torch.float32, device=dev)
        y16 = torch.empty((M, H), dtype=torch.bfloat16, device=dev)
        mean = torch.empty(M, dtype=torch.float32, device=dev)
        rstd = torch.empty(M, dtype=torch.float32, device=dev)
        state = rng.state if (rng is not None and p > 0) else None
        check(lib.crv_ln_fwd(_p(g2), ops.DT_BF16 if g2.dtype == torch.bfloat16 else ops.DT_F32, _p(r2), _p(gamma),
                             _p(beta), float(eps), float(p), _p(state), int(site), _p(y32), _p(y16), _p(mean),
                             _p(rstd), M, H, _stream()), "crv_ln_fwd")
        ctx.save_for_backward(g2, r2, gamma, mean, rstd)
        ctx.p, ctx.site, ctx.state, ctx.shape = float(p), int(site), state, g.shape
        ctx.need_res = res32 is not None and res32.requires_grad
        ctx.need_param = gamma.requires_grad or beta.requires_grad
        return y32.view(g.shape), y16.view(g.shape)

    @staticmethod
    def backward(ctx, dy32, dy16):
        g2, r2, gamma, mean, rstd = ctx.saved_tensors
        M, H = g2.shape
        d32 = dy32.reshape(M, H).contiguous() if dy32 is not None else None
        d16 = dy16.reshape(M, H).contiguous() if dy16 is not None else None
        dg = torch.empty((M, H), dtype=torch.bfloat16, device=g2.device)
        dres = torch.empty((M, H), dtype=torch.float32, device=g2.device) if ctx.need_res else None
        part = None
        if ctx.need_param:
            part = torch.empty((lib.crv_ln_bwd_partials_bytes(M, H) // (8 * H), 2 * H), dtype=torch.float32,
                               device=g2.device)
        check(lib.crv_ln_bwd(_p(d32), _p(d16), _p(g2), ops.DT_BF16 if g2.dtype == torch.bfloat16 else ops.DT_F32,
                             _p(r2), _p(gamma), _p(mean), _p(rstd), ctx.p, _p(ctx.state), ctx.site, _p(dg),
                             ops.DT_BF16, _p(dres), _p(part), M, H, _stream()), "crv_ln_bwd")
        dgamma = dbeta = None
        if part is not None:
            pair = getattr(gamma, "_arena_pair", None)
            if pair is not None and gamma.grad is not None:     # [dgamma | dbeta] added straight into the arena
                ops.partial_reduce(part, pair, accumulate=True)
            else:
                both = torch.empty(2 * H, dtype=torch.float32, device=g2.device)
                ops.partial_reduce(part, both)
                dgamma, dbeta = both[:H], both[H:]
        return (dg.view(ctx.shape), dres.view(ctx.shape) if dres is not None else None, dgamma, dbeta, None, None, None,
                None)


def drop_add_layernorm(g, res32, ln, p, site, training):
    """(y fp32, y bf16) = LayerNorm(dropout(g) + res32): LxmertAttentionOutput / LxmertOutput tail."""
    p = float(p) if training else 0.0
    rng = RngState.get(g.device) if p > 0 else None
    return DropAddLayerNormFn.apply(g, res32, ln.weight, ln.bias, ln.eps, p, site, rng)


class LayerNormBf16Fn(torch.autograd.Function):
    """y (bf16) = LayerNorm(x) for a bf16 x with frozen gamma / beta: crv_ln_fwd / crv_ln_bwd without residual, dropout
    or fp32 copy -- 4 B per element each way.  Under bf16 autocast torch runs the same fp32 arithmetic as three passes
    (cast to fp32, LayerNorm, cast back: 20 B per element) and three more in the backward."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        H = x.shape[-1]
        x2 = x.reshape(-1, H)
        x2 = x2 if x2.is_contiguous() else x2.contiguous()
        M = x2.shape[0]
        y16 = torch.empty((M, H), dtype=torch.bfloat16, device=x.device)
        mean = torch.empty(M, dtype=torch.float32, device=x.device)
        rstd = torch.empty(M, dtype=torch.float32, device=x.device)
        check(lib.crv_ln_fwd(_p(x2), ops.DT_BF16, None, _p(gamma), _p(beta), float(eps), 0.0, None, 0, None, _p(y16),
                             _p(mean), _p(rstd), M, H, _stream()), "crv_ln_fwd")
        ctx.save_for_backward(x2, gamma, mean, rstd)
        ctx.shape = x.shape
        return y16.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, gamma, mean, rstd = ctx.saved_tensors
        M, H = x2.shape
        d = dy.reshape(M, H)
        if d.dtype not in (torch.bfloat16, torch.float32):
            d = d.float()
        d = d if d.is_contiguous() else d.contiguous()
        d32, d16 = (d, None) if d.dtype == torch.float32 else (None, d)
        dx = torch.empty((M, H), dtype=torch.bfloat16, device=x2.device)
        check(lib.crv_ln_bwd(_p(d32), _p(d16), _p(x2), ops.DT_BF16, None, _p(gamma), _p(mean), _p(rstd), 0.0, None, 0,
                             _p(dx), ops.DT_BF16, None, None, M, H, _stream()), "crv_ln_bwd")
        return dx.view(ctx.shape), None, None, None


def layernorm_bf16_usable(x, ln):
    H = x.shape[-1]
    return (x.is_cuda and x.dtype == torch.bfloat16 and H % 128 == 0 and H <= 1024 and x.numel() > 0
            and tuple(ln.normalized_shape) == (H,) and ln.weight is not None and ln.bias is not None
            and ln.weight.dtype == torch.float32 and not ln.weight.requires_grad and not ln.bias.requires_grad)


def layernorm_bf16(x, ln):
    """bf16 LayerNorm(x) of a frozen nn.LayerNorm on a bf16 activation (check layernorm_bf16_usable first)."""
    return LayerNormBf16Fn.apply(x, ln.weight, ln.bias, ln.eps)


# ----------------------------------------------------------------------------- few-query attention (mPLUG text side)
class FewQueryAttentionFn(torch.autograd.Function):
    """ctx = dropout(softmax(Q K^T / 8 + mask)) V on [B, L, heads * 64] bf16 projections (crv_fq_attention_fwd / _bwd):
    Lq <= 16 queries, Lk <= 1024 keys.  The forward saves the probabilities with the dropout decision in their sign."""

    @staticmethod
    def forward(ctx, q, k, v, mask, heads, p, site, rng):
        B, Lq, D = q.shape
        Lk = k.shape[1]
        out = torch.empty((B, Lq, D), dtype=torch.bfloat16, device=q.device)
        probs = torch.empty((B, heads, Lq, Lk), dtype=torch.bfloat16, device=q.device)
        state = rng.state if (rng is not None and p > 0) else None
        p = float(p) if state is not None else 0.0
        sb = sq = 0
        if mask is not None:
            sb = mask.stride(0) if mask.shape[0] > 1 else 0
            sq = mask.stride(2) if mask.shape[2] > 1 else 0
        scale = 1.0 / math.sqrt(D // heads)
        check(lib.crv_fq_attention_fwd(_p(q), _p(k), _p(v), _p(mask), sb, sq, _p(out), _p(probs), B, heads, Lq, Lk,
                                       scale, p, _p(state), int(site), _stream()), "crv_fq_attention_fwd")
        ctx.save_for_backward(q, k, v, probs)
        ctx.heads, ctx.p, ctx.scale = heads, p, scale
        return out

    @staticmethod
    def backward(ctx, dout):
        q, k, v, probs = ctx.saved_tensors
        B, Lq, D = q.shape
        Lk = k.shape[1]
        d = dout if (dout.dtype == torch.bfloat16 and dout.is_contiguous()) else dout.to(torch.bfloat16).contiguous()
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        check(lib.crv_fq_attention_bwd(_p(d), _p(q), _p(k), _p(v), _p(probs), _p(dq), _p(dk), _p(dv), B, ctx.heads, Lq,
                                       Lk, ctx.scale, ctx.p, _stream()), "crv_fq_attention_bwd")
        return dq, dk, dv, None, None, None, None, None


def few_query_attention_usable(q, k, v, mask, heads):
    """q / k / v: the projections' outputs [B, L, heads * 64]; mask: None or additive [B | 1, 1, Lq | 1, Lk]."""
    if not (q.is_cuda and q.dtype == k.dtype == v.dtype == torch.bfloat16 and q.dim() == 3 and k.shape == v.shape
            and q.shape[0] == k.shape[0] and q.shape[2] == k.shape[2] == heads * 64
            and 0 < q.shape[1] <= 16 and 0 < k.shape[1] <= 1024
            and q.is_contiguous() and k.is_contiguous() and v.is_contiguous()):
        return False
    if mask is not None:
        if not (mask.dim() == 4 and mask.shape[1] == 1 and mask.shape[0] in (1, q.shape[0])
                and mask.shape[2] in (1, q.shape[1]) and mask.shape[3] == k.shape[1] and not mask.requires_grad):
            return False
    return True


def few_query_attention(q, k, v, mask, heads, p, site, training):
    """Attention context [B, Lq, heads * 64] (bf16) of BertSelfAttention; check few_query_attention_usable first."""
    if mask is not None and (mask.dtype != torch.float32 or mask.stride(3) != 1):
        mask = mask.float().contiguous()
    p = float(p) if training else 0.0
    rng = RngState.get(q.device) if p > 0 else None
    return FewQueryAttentionFn.apply(q, k, v, mask, heads, p, site, rng)


class LnAvgDropFn(torch.autograd.Function):
    """(y fp32, y bf16) = dropout((LN_a(a) + LN_b(b)) / 2), or dropout(LN_a(a)) when b is None: the entry blocks of LXMERT
    (LxmertVisualFeatureEncoder / LxmertEmbeddings, hg_transformers/modeling_lxmert.py:576-592, 744-770) in one pass each
    way (crv_ln_avg_drop_fwd / _bwd).  LayerNorm parameters are constants here (frozen in stage 2)."""

    @staticmethod
    def forward(ctx, a, b, ga, ba, gb, bb, eps, p, site, rng):
        H = a.shape[-1]
        a2 = a.reshape(-1, H)
        a2 = a2 if a2.is_contiguous() else a2.contiguous()
        b2 = None
        if b is not None:
            b2 = b.reshape(-1, H)
            b2 = b2 if b2.is_contiguous() else b2.contiguous()
        M = a2.shape[0]
        dev = a.device
        y32 = torch.empty((M, H), dtype=torch.float32, device=dev)
        y16 = torch.empty((M, H), dtype=torch.bfloat16, device=dev)
        stats = torch.empty((M, 4), dtype=torch.float32, device=dev)
        state = rng.state if (rng is not None and p > 0) else None
        check(lib.crv_ln_avg_drop_fwd(_p(a2), _p(b2), _p(ga), _p(ba), _p(gb), _p(bb), float(eps), float(p), _p(state),
                                      int(site), _p(y32), _p(y16), _p(stats), M, H, _stream()), "crv_ln_avg_drop_fwd")
        ctx.save_for_backward(a2, b2, ga, gb, stats)
        ctx.p, ctx.site, ctx.state, ctx.shape = float(p), int(site), state, a.shape
        ctx.need = (ctx.needs_input_grad[0], b is not None and ctx.needs_input_grad[1])
        ctx.set_materialize_grads(False)
        return y32.view(a.shape), y16.view(a.shape)

    @staticmethod
    def backward(ctx, dy32, dy16):
        a2, b2, ga, gb, stats = ctx.saved_tensors
        if (dy32 is None and dy16 is None) or not any(ctx.need):
            return (None,) * 10
        M, H = a2.shape
        d32 = dy32.reshape(M, H).contiguous() if dy32 is not None else None
        d16 = dy16.reshape(M, H).contiguous() if dy16 is not None else None
        da = torch.empty((M, H), dtype=torch.float32, device=a2.device) if ctx.need[0] else None
        db = torch.empty((M, H), dtype=torch.float32, device=a2.device) if ctx.need[1] else None
        check(lib.crv_ln_avg_drop_bwd(_p(d32), _p(d16), _p(a2), _p(b2), _p(ga), _p(gb), _p(stats), ctx.p, _p(ctx.state),
                                      ctx.site, _p(da), _p(db), M, H, _stream()), "crv_ln_avg_drop_bwd")
        return (da.view(ctx.shape) if da is not None else None, db.view(ctx.shape) if db is not None else None,
                None, None, None, None, None, None, None, None)


def ln_avg_drop_usable(a, b, ln_a, ln_b):
    """The fused entry block applies on CUDA, fp32 inputs, H % 128 == 0, H <= 1024 and FROZEN LayerNorm parameters (stage
    2); anything else (stage 3 trains them, CPU) keeps the PyTorch ops.  CRVQA_FUSED_ENTRY=0 turns it off."""
    if os.environ.get("CRVQA_FUSED_ENTRY", "1") == "0" or os.environ.get("CRVQA_FUSED", "1") == "0":
        return False
    H = a.shape[-1]
    if not a.is_cuda or a.dtype != torch.float32 or H % 128 or H > 1024:
        return False
    if b is not None and (b.dtype != torch.float32 or b.shape != a.shape):
        return False
    for ln in (ln_a, ln_b):
        if ln is None:
            continue
        if ln.weight is None or ln.bias is None or ln.weight.requires_grad or ln.bias.requires_grad:
            return False
        if tuple(ln.normalized_shape) != (H,):
            return False
    if ln_b is not None and ln_b.eps != ln_a.eps:
        return False
    return True


def ln_avg_drop(a, b, ln_a, ln_b, p, site, training):
    """y fp32 of the entry block, with the bf16 copy the first GEMM reads attached as ``y._crv_bf16`` (an output of the
    same autograd node, so gradients reaching either copy flow back)."""
    p = float(p) if training else 0.0
    rng = RngState.get(a.device) if p > 0 else None
    y32, y16 = LnAvgDropFn.apply(a, b, ln_a.weight, ln_a.bias, ln_b.weight if ln_b is not None else None,
                                 ln_b.bias if ln_b is not None else None, ln_a.eps, p, site, rng)
    y32._crv_bf16 = y16
    return y32


class GeluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, u):
        u = u if u.is_contiguous() else u.contiguous()
        y = torch.empty_like(u)
        check(lib.crv_gelu_fwd(_p(u), _p(y), u.numel(), _stream()), "crv_gelu_fwd")
        ctx.save_for_backward(u)
        return y

    @staticmethod
    def backward(ctx, dy):
        (u,) = ctx.saved_tensors
        dy = dy if dy.is_contiguous() else dy.contiguous()
        du = torch.empty_like(u)
        check(lib.crv_gelu_bwd(_p(u), _p(dy), _p(du), u.numel(), _stream()), "crv_gelu_bwd")
        return du


def gelu_bf16(u):
    return GeluFn.apply(u)


class QuickGeluFn(torch.autograd.Function):
    """x sigmoid(1.702 x) on bf16 (the CLIP vision tower of mPLUG, mPLUG/models/clip/model.py:25-27): one pass each way."""

    @staticmethod
    def forward(ctx, u):
        u = u if u.is_contiguous() else u.contiguous()
        y = torch.empty_like(u)
        check(lib.crv_quick_gelu_fwd(_p(u), _p(y), u.numel(), _stream()), "crv_quick_gelu_fwd")
        ctx.save_for_backward(u)
        return y

    @staticmethod
    def backward(ctx, dy):
        (u,) = ctx.saved_tensors
        dy = dy if dy.is_contiguous() else dy.contiguous()
        du = torch.empty_like(u)
        check(lib.crv_quick_gelu_bwd(_p(u), _p(dy), _p(du), u.numel(), _stream()), "crv_quick_gelu_bwd")
        return du


def quick_gelu_bf16(u):
    return QuickGeluFn.apply(u)


# ----------------------------------------------------------------------------- small-sequence attention
def _off(t, elems):
    return ctypes.c_void_p(t.data_ptr() + 2 * elems)


def _save_probs():
    """Training keeps the (signed) softmax probabilities of the forward for the backward (crv_attention_fwd_p / _bwd_p);
    CRVQA_ATTN_SAVE_P=0 selects the recomputing backward (crv_attention_bwd)."""
    return os.environ.get("CRVQA_ATTN_SAVE_P", "1") != "0"


def _probs_buffer(B, heads, Sq, Sk, device):
    return torch.empty((B * heads, Sq, lib.crv_attention_probs_pitch(Sk)), dtype=torch.bfloat16, device=device)


class SmallAttentionFn(torch.autograd.Function):
    """softmax(Q K^T / sqrt(d) + mask) V with dropout for S <= 64, d = 64 (crv_attention_fwd / _bwd).
    kind 0: srcs = (qkv [B,S,3H],)   kind 1: srcs = (q [B,Sq,H], kv [B,Sk,2H])   kind 2: srcs = (q, k, v)."""

    @staticmethod
    def _views(kind, srcs):
        if kind == 0:
            (qkv,) = srcs
            H = qkv.shape[-1] // 3
            return H, [(qkv, 0), (qkv, H), (qkv, 2 * H)]
        if kind == 1:
            q, kv = srcs
            H = q.shape[-1]
            return H, [(q, 0), (kv, 0), (kv, H)]
        q, k, v = srcs
        return q.shape[-1], [(q, 0), (k, 0), (v, 0)]

    @staticmethod
    def forward(ctx, kind, heads, mask, p, site, rng, *srcs):
        srcs = tuple(s if s.is_contiguous() else s.contiguous() for s in srcs)
        H, views = SmallAttentionFn._views(kind, srcs)
        (qt, qo), (kt, ko), (vt, vo) = views
        B, Sq, Sk = qt.shape[0], qt.shape[1], kt.shape[1]
        out = torch.empty((B, Sq, H), dtype=torch.bfloat16, device=qt.device)
        scale = 1.0 / (H // heads) ** 0.5
        state = rng.state if (rng is not None and p > 0) else None
        m2 = None
        if mask is not None:
            m2 = mask.reshape(B, Sk).float().contiguous()
        probs = None
        if any(ctx.needs_input_grad[6:]) and _save_probs():
            probs = _probs_buffer(B, heads, Sq, Sk, qt.device)
        _attn_fwd_raw(views, m2, out, B, heads, Sq, Sk, scale, p, state, site, probs)
        ctx.save_for_backward(*srcs)
        ctx.probs = probs
        ctx.cfg = (kind, heads, m2, float(p), int(site), state, scale)
        return out

    @staticmethod
    def backward(ctx, dout):
        kind, heads, m2, p, site, state, scale = ctx.cfg
        srcs = ctx.saved_tensors
        H, views = SmallAttentionFn._views(kind, srcs)
        grads = tuple(torch.empty_like(s) for s in srcs)
        _, gviews = SmallAttentionFn._views(kind, grads)
        (qt, qo), (kt, ko), (vt, vo) = views
        (dq, dqo), (dk, dko), (dv, dvo) = gviews
        B, Sq, Sk = qt.shape[0], qt.shape[1], kt.shape[1]
        dout = dout if dout.is_contiguous() else dout.contiguous()
        _attn_bwd_raw(views, m2, dout, gviews, B, heads, Sq, Sk, scale, p, state, site, ctx.probs)
        return (None, None, None, None, None, None) + grads


def _attn_fwd_raw(views, m2, out, B, heads, Sq, Sk, scale, p, state, site, probs=None):
    (qt, qo), (kt, ko), (vt, vo) = views
    if probs is not None:
        check(lib.crv_attention_fwd_p(_off(qt, qo), qt.stride(0), qt.stride(1), _off(kt, ko), kt.stride(0), kt.stride(1),
                                      _off(vt, vo), vt.stride(0), vt.stride(1), _p(m2), _p(out), _p(probs), B, heads, Sq,
                                      Sk, scale, float(p), _p(state), int(site), _stream()), "crv_attention_fwd_p")
        return
    check(lib.crv_attention_fwd(_off(qt, qo), qt.stride(0), qt.stride(1), _off(kt, ko), kt.stride(0), kt.stride(1),
                                _off(vt, vo), vt.stride(0), vt.stride(1), _p(m2), _p(out), B, heads, Sq, Sk,
                                scale, float(p), _p(state), int(site), _stream()), "crv_attention_fwd")


def _attn_bwd_raw(views, m2, dout, gviews, B, heads, Sq, Sk, scale, p, state, site, probs=None):
    (qt, qo), (kt, ko), (vt, vo) = views
    (dq, dqo), (dk, dko), (dv, dvo) = gviews
    if probs is not None:
        check(lib.crv_attention_bwd_p(_off(qt, qo), qt.stride(0), qt.stride(1), _off(kt, ko), kt.stride(0), kt.stride(1),
                                      _off(vt, vo), vt.stride(0), vt.stride(1), _p(probs), _p(dout),
                                      _off(dq, dqo), dq.stride(0), dq.stride(1), _off(dk, dko), dk.stride(0),
                                      dk.stride(1), _off(dv, dvo), dv.stride(0), dv.stride(1), B, heads, Sq, Sk, scale,
                                      float(p), _stream()), "crv_attention_bwd_p")
        return
    check(lib.crv_attention_bwd(_off(qt, qo), qt.stride(0), qt.stride(1), _off(kt, ko), kt.stride(0), kt.stride(1),
                                _off(vt, vo), vt.stride(0), vt.stride(1), _p(m2), _p(dout),
                                _off(dq, dqo), dq.stride(0), dq.stride(1), _off(dk, dko), dk.stride(0), dk.stride(1),
                                _off(dv, dvo), dv.stride(0), dv.stride(1), B, heads, Sq, Sk, scale, p, _p(state),
                                site, _stream()), "crv_attention_bwd")


class CrossPairAttentionFn(torch.autograd.Function):
    """Both directions of LxmertXLayer's cross attention (hg_transformers/modeling_lxmert.py:947-958) on the fused
    projections qkv_l = [q|k|v](lang), qkv_v = [q|k|v](visn) of the SHARED visual_attention module:
        ctx_l = attend(q(lang), k(visn), v(visn), visn mask)      ctx_v = attend(q(visn), k(lang), v(lang), lang mask)
    The two backward calls write disjoint slices that together cover d(qkv_l) and d(qkv_v) exactly once, so no
    zero-fill and no gradient add is needed although each projection feeds both directions."""

    @staticmethod
    def forward(ctx, heads, mask_l, mask_v, p, site_l, site_v, rng, qkv_l, qkv_v):
        qkv_l = qkv_l if qkv_l.is_contiguous() else qkv_l.contiguous()
        qkv_v = qkv_v if qkv_v.is_contiguous() else qkv_v.contiguous()
        H = qkv_l.shape[-1] // 3
        B, Sl, Sv = qkv_l.shape[0], qkv_l.shape[1], qkv_v.shape[1]
        scale = 1.0 / (H // heads) ** 0.5
        state = rng.state if (rng is not None and p > 0) else None
        ml = mask_l.reshape(B, Sl).float().contiguous() if mask_l is not None else None
        mv = mask_v.reshape(B, Sv).float().contiguous() if mask_v is not None else None
        out_l = torch.empty((B, Sl, H), dtype=torch.bfloat16, device=qkv_l.device)
        out_v = torch.empty((B, Sv, H), dtype=torch.bfloat16, device=qkv_l.device)
        pl = pv = None
        if any(ctx.needs_input_grad[7:]) and _save_probs():
            pl, pv = _probs_buffer(B, heads, Sl, Sv, qkv_l.device), _probs_buffer(B, heads, Sv, Sl, qkv_l.device)
        _attn_fwd_raw([(qkv_l, 0), (qkv_v, H), (qkv_v, 2 * H)], mv, out_l, B, heads, Sl, Sv, scale, p, state, site_l, pl)
        _attn_fwd_raw([(qkv_v, 0), (qkv_l, H), (qkv_l, 2 * H)], ml, out_v, B, heads, Sv, Sl, scale, p, state, site_v, pv)
        ctx.save_for_backward(qkv_l, qkv_v)
        ctx.probs = (pl, pv)
        ctx.set_materialize_grads(False)
        ctx.cfg = (heads, ml, mv, float(p), int(site_l), int(site_v), state, scale)
        return out_l, out_v

    @staticmethod
    def backward(ctx, do_l, do_v):
        heads, ml, mv, p, site_l, site_v, state, scale = ctx.cfg
        qkv_l, qkv_v = ctx.saved_tensors
        H = qkv_l.shape[-1] // 3
        B, Sl, Sv = qkv_l.shape[0], qkv_l.shape[1], qkv_v.shape[1]
        if do_l is None and do_v is None:
            return (None,) * 9
        d_l, d_v = torch.empty_like(qkv_l), torch.empty_like(qkv_v)
        pl, pv = ctx.probs
        if do_l is not None:
            do_l = do_l if do_l.is_contiguous() else do_l.contiguous()
            _attn_bwd_raw([(qkv_l, 0), (qkv_v, H), (qkv_v, 2 * H)], mv, do_l, [(d_l, 0), (d_v, H), (d_v, 2 * H)],
                          B, heads, Sl, Sv, scale, p, state, site_l, pl)
        else:                       # this direction feeds nothing: its slices of the two gradients are zero
            d_l[..., :H].zero_()
            d_v[..., H:].zero_()
        if do_v is not None:
            do_v = do_v if do_v.is_contiguous() else do_v.contiguous()
            _attn_bwd_raw([(qkv_v, 0), (qkv_l, H), (qkv_l, 2 * H)], ml, do_v, [(d_v, 0), (d_l, H), (d_l, 2 * H)],
                          B, heads, Sv, Sl, scale, p, state, site_v, pv)
        else:                       # the last cross layer: nothing reads the vision output
            d_v[..., :H].zero_()
            d_l[..., H:].zero_()
        return None, None, None, None, None, None, None, d_l, d_v


def small_attention(kind, heads, mask, p, site, training, *srcs):
    p = float(p) if training else 0.0
    rng = RngState.get(srcs[0].device) if p > 0 else None
    return SmallAttentionFn.apply(kind, heads, mask, p, site, rng, *srcs)


# ----------------------------------------------------------------------------- layer plans
def _arena_ready(*mods):
    for m in mods:
        a = getattr(m, "_arena", None)
        if a is None or not a.cache_on or not a.module_ready(m) or a.cached_masked_weight(m) is None:
            return False
    return True


class AttentionPlan:
    """Fast path of LxmertSelfAttentionLayer / LxmertCrossAttentionLayer (att = .self or .att, out = .output)."""

    def __init__(self, att, out):
        self.att, self.out = att, out
        self.mods = (att.query, att.key, att.value, out.dense)
        self.qkv = self.q = self.kv = self.ao = None
        self.site = RngState.new_site()
        self.site_att = RngState.new_site()

    def ready(self):
        if not _arena_ready(*self.mods):
            return False
        if self.ao is None:
            a = self.att
            try:
                self.qkv = ProjectionGroup([a.query, a.key, a.value])
            except ValueError:
                self.qkv = None
            self.q = ProjectionGroup([a.query])
            try:
                self.kv = ProjectionGroup([a.key, a.value])
            except ValueError:
                self.kv = None
            self.k1, self.v1 = ProjectionGroup([a.key]), ProjectionGroup([a.value])
            self.ao = ProjectionGroup([self.out.dense])
        return True

    def _sdpa(self, q, k, v, mask, training):
        a = self.att
        B, Sq, _ = q.shape
        h, d = a.num_attention_heads, a.attention_head_size
        q = q.view(B, Sq, h, d).transpose(1, 2)
        k = k.view(B, k.shape[1], h, d).transpose(1, 2)
        v = v.view(B, v.shape[1], h, d).transpose(1, 2)
        if mask is not None:
            mask = mask.to(q.dtype)
        p = a.dropout.p if training else 0.0
        ctx = F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=p)
        return ctx.transpose(1, 2).reshape(B, Sq, h * d)

    def _small(self, Sq, Sk):
        # crv_attention_*: one warp per (batch, head), register-resident; CRVQA_ATTN=sdpa selects the library
        a = self.att
        return (os.environ.get("CRVQA_ATTN", "small") == "small" and a.attention_head_size == 64
                and Sq <= 64 and Sk <= 64)

    def _attend(self, kind, mask, training, site, Sq, Sk, *srcs):
        a = self.att
        if self._small(Sq, Sk):
            m = None if mask is None else mask.reshape(mask.shape[0], -1)
            return small_attention(kind, a.num_attention_heads, m, a.dropout.p, site, training, *srcs)
        H = a.head_size
        if kind == 0:
            q, k, v = srcs[0].split(H, dim=-1)
        elif kind == 1:
            q = srcs[0]
            k, v = srcs[1].split(H, dim=-1)
        else:
            q, k, v = srcs
        return self._sdpa(q, k, v, mask, training)

    def self_attention(self, x32, x16, mask, training):
        S = x16.shape[1]
        if _grouped_on() and self.qkv is not None:
            return self_attention_multi([(self, x32, x16, mask)], training)[0]
        if self.qkv is not None:
            ctx = self._attend(0, mask, training, self.site_att, S, S, group_linear(self.qkv, x16))
        else:
            ctx = self._attend(2, mask, training, self.site_att, S, S, group_linear(self.q, x16),
                               group_linear(self.k1, x16), group_linear(self.v1, x16))
        ao = group_linear(self.ao, ctx)
        return drop_add_layernorm(ao, x32, self.out.LayerNorm, self.out.dropout.p, self.site, training)

    def cross_attention(self, x32, x16, c16, ctx_mask, training, site):
        Sq, Sk = x16.shape[1], c16.shape[1]
        q = group_linear(self.q, x16)
        if self.kv is not None:
            ctx = self._attend(1, ctx_mask, training, site + 1, Sq, Sk, q, group_linear(self.kv, c16))
        else:
            ctx = self._attend(2, ctx_mask, training, site + 1, Sq, Sk, q, group_linear(self.k1, c16),
                               group_linear(self.v1, c16))
        ao = group_linear(self.ao, ctx)
        return drop_add_layernorm(ao, x32, self.out.LayerNorm, self.out.dropout.p, site, training)


class FfnPlan:
    """Fast path of LxmertIntermediate + LxmertOutput."""

    def __init__(self, inter, out):
        self.inter, self.out = inter, out
        self.gi = self.go = None
        self.site = RngState.new_site()

    def ready(self):
        if not _arena_ready(self.inter.dense, self.out.dense):
            return False
        if self.gi is None:
            self.gi = ProjectionGroup([self.inter.dense])
            self.go = ProjectionGroup([self.out.dense])
        return True

    def __call__(self, x32, x16, training):
        if _grouped_on():
            return ffn_multi([(self, x32, x16)], training)[0]
        a = gelu_bf16(group_linear(self.gi, x16))
        o = group_linear(self.go, a)
        return drop_add_layernorm(o, x32, self.out.LayerNorm, self.out.dropout.p, self.site, training)


# ----------------------------------------------------------------------------- lockstep helpers (grouped launches)
def self_attention_multi(items, training):
    """items: (AttentionPlan, x32, x16, mask) of INDEPENDENT sub-layers (language / vision stack in lockstep, the two
    self-attention blocks of a cross layer): their QKV projections share one grouped launch, so do their output
    projections.  Returns [(y32, y16)]."""
    if not _grouped_on() or any(pl.qkv is None for pl, *_ in items):
        return [pl.self_attention(x32, x16, mask, training) for pl, x32, x16, mask in items]
    qkvs = multi_linear([(pl.qkv, x16, torch.bfloat16, False, None) for pl, _, x16, _ in items])
    ctxs = []
    for (pl, _, x16, mask), qkv in zip(items, qkvs):
        S = x16.shape[1]
        ctxs.append(pl._attend(0, mask, training, pl.site_att, S, S, qkv))
    aos = multi_linear([(pl.ao, c, torch.bfloat16, False, None) for (pl, *_), c in zip(items, ctxs)])
    return [drop_add_layernorm(ao, x32, pl.out.LayerNorm, pl.out.dropout.p, pl.site, training)
            for (pl, x32, _, _), ao in zip(items, aos)]


def ffn_multi(items, training):
    """items: (FfnPlan, x32, x16).  FF1 of all members in one launch with bias + GELU in the epilogue (pre-activation
    kept for the backward), FF2 likewise with gelu'(u) folded into its dX epilogue."""
    if not _grouped_on():
        return [f(x32, x16, training) for f, x32, x16 in items]
    fuse = all(ops.grouped_2cta_ok(x16.numel() // x16.shape[-1], f.gi.N) and
               ops.grouped_2cta_ok(x16.numel() // x16.shape[-1], f.go.N) for f, _, x16 in items)
    if fuse:
        inter = multi_linear([(f.gi, x16, torch.bfloat16, True, None) for f, _, x16 in items])
        outs = multi_linear([(f.go, a, torch.bfloat16, False, u) for (f, _, _), (a, u) in zip(items, inter)])
    else:
        us = multi_linear([(f.gi, x16, torch.bfloat16, False, None) for f, _, x16 in items])
        outs = multi_linear([(f.go, gelu_bf16(u), torch.bfloat16, False, None) for (f, _, _), u in zip(items, us)])
    return [drop_add_layernorm(o, x32, f.out.LayerNorm, f.out.dropout.p, f.site, training)
            for (f, x32, _), o in zip(items, outs)]


def cross_attention_pair(plan, lang32, lang16, visn32, visn16, lang_mask, visn_mask, training, site_l, site_v):
    """Both directions of a cross layer's shared visual_attention block: one grouped launch for [q|k|v] of both
    modalities, the attention pair, one grouped launch for the two output projections.  Returns
    ((lang_x32, lang_x16), (visn_x32, visn_x16))."""
    a = plan.att
    Sl, Sv = lang16.shape[1], visn16.shape[1]
    if not _grouped_on() or plan.qkv is None or not plan._small(Sl, Sv):
        return (plan.cross_attention(lang32, lang16, visn16, visn_mask, training, site_l),
                plan.cross_attention(visn32, visn16, lang16, lang_mask, training, site_v))
    qkv_l, qkv_v = multi_linear([(plan.qkv, lang16, torch.bfloat16, False, None),
                                 (plan.qkv, visn16, torch.bfloat16, False, None)])
    p = float(a.dropout.p) if training else 0.0
    rng = RngState.get(lang16.device) if p > 0 else None
    ctx_l, ctx_v = CrossPairAttentionFn.apply(a.num_attention_heads, lang_mask, visn_mask, p, site_l + 1, site_v + 1,
                                              rng, qkv_l, qkv_v)
    ao_l, ao_v = multi_linear([(plan.ao, ctx_l, torch.bfloat16, False, None),
                               (plan.ao, ctx_v, torch.bfloat16, False, None)])
    ln, pd = plan.out.LayerNorm, plan.out.dropout.p
    return (drop_add_layernorm(ao_l, lang32, ln, pd, site_l, training),
            drop_add_layernorm(ao_v, visn32, ln, pd, site_v, training))


def cross_attention_lang_only(plan, lang32, lang16, visn16, visn_mask, training, site_l):
    """The language direction of a cross layer alone (q from language, k | v from vision): what is left of the last
    cross layer when nothing reads its vision output."""
    a = plan.att
    Sl, Sv = lang16.shape[1], visn16.shape[1]
    if plan.kv is None or not plan._small(Sl, Sv):
        return plan.cross_attention(lang32, lang16, visn16, visn_mask, training, site_l)
    q, kv = multi_linear([(plan.q, lang16, torch.bfloat16, False, None), (plan.kv, visn16, torch.bfloat16, False, None)])
    m = None if visn_mask is None else visn_mask.reshape(visn_mask.shape[0], -1)
    ctx = small_attention(1, a.num_attention_heads, m, a.dropout.p, site_l + 1, training, q, kv)
    (ao,) = multi_linear([(plan.ao, ctx, torch.bfloat16, False, None)])
    return drop_add_layernorm(ao, lang32, plan.out.LayerNorm, plan.out.dropout.p, site_l, training)
