"""Shared implementation of the three reference masker modules.

``masking/maskers.py``, ``maskers_Robust.py`` and ``maskers_visualBert.py`` of the reference are
620-line near-copies that differ only in (a) which name table ``chain_module_names`` reads, (b)
whether it also returns modality / module / layer dictionaries, and (c) whether ``Masker`` takes an
``hpmodel`` whose ``zerorate_dict`` gives every module its own initial sparsity.  Here the logic
exists once; the three public modules instantiate it.

What changed underneath (B200-native):
  * ``MaskedLinear1.forward`` is ONE fused kernel per GEMM: the TMA-fed prologue binarises the score
    tile against the module threshold in shared memory and masks the bf16 weight tile before
    tcgen05 MMAs accumulate in TMEM (reference: 5 elementwise kernels + ``mul`` + fp32 ``addmm`` per
    call, masking/maskers.py:325-366).  Its backward gives dX through the masked weight and the
    straight-through score gradient dS = (dY^T X) (.) W from the GEMM epilogue.
  * the embedding branch gathers rows and masks only those (reference masks the whole 30522 x 768 table).
  * magnitude initialisation of all score tensors is one batched exact radix select on |W|.
There is no CPU implementation of these ops: without a CUDA device they raise.
"""
import json
import math
import warnings

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from crvqa import ops

# --------------------------------------------------------------------------- name tables
# abbreviation -> attribute path under "<ptl>."  ("{l}" = layer index).  Reference tables:
# masking/maskers.py:13-67, maskers_visualBert.py:24-34.
_ATT = {"K": "attention.self.key", "Q": "attention.self.query", "V": "attention.self.value",
        "AO": "attention.output.dense", "I": "intermediate.dense", "O": "output.dense"}


def _table(spec):
    return {abbr: (lambda ptl, l, _p=path: f"{ptl}.{_p}".replace("{l}", str(l))) for abbr, path in spec.items()}


_bert_spec = {a: "encoder.layer.{l}." + p for a, p in _ATT.items()}
_bert_spec.update({"P": "pooler.dense", "E": "embeddings.word_embeddings"})
_bert_roberta_names = _table(_bert_spec)

_visualbert_names = _table({**_bert_spec, "VP": "embeddings.visual_projection"})

_lx = {"E": "embeddings.word_embeddings", "VV": "encoder.visn_fc.visn_fc", "VB": "encoder.visn_fc.box_fc"}
for _pre, _stack in (("l", "layer"), ("v", "r_layers")):
    for _a, _p in _ATT.items():
        _lx[_pre + _a] = f"encoder.{_stack}.{{l}}.{_p}"
_X = "encoder.x_layers.{l}."
for _a, _p in (("K", "key"), ("Q", "query"), ("V", "value")):
    _lx["vlV" + _a] = _X + "visual_attention.att." + _p
    _lx["vlLa" + _a] = _X + "lang_self_att.self." + _p
    _lx["vlVa" + _a] = _X + "visn_self_att.self." + _p
_lx.update({"vlVAO": _X + "visual_attention.output.dense", "vlLaAO": _X + "lang_self_att.output.dense",
            "vlVaAO": _X + "visn_self_att.output.dense", "vlLi": _X + "lang_inter.dense",
            "vlLo": _X + "lang_output.dense", "vlVi": _X + "visn_inter.dense", "vlVo": _X + "visn_output.dense",
            "P": "pooler.dense"})
_lxmert_names = _table(_lx)

_distilbert_names = {
    a: (lambda _, l, _p=p: f"distilbert.transformer.layer.{l}.{_p}")
    for a, p in {"K": "attention.k_lin", "Q": "attention.q_lin", "V": "attention.v_lin",
                 "AO": "attention.out_lin", "I": "ffn.lin1", "O": "ffn.lin2"}.items()
}
_distilbert_names["P"] = lambda _, l: "pre_classifier"


def lxmert_modality(abbre):
    """abbreviation -> 'Lang' | 'Vis' | 'Fus' | 'P' (reference maskers_Robust.py:79)."""
    if abbre == "P":
        return "P"
    if abbre.startswith("vl"):
        return "Fus"
    if abbre in ("VV", "VB") or abbre.startswith("v"):
        return "Vis"
    return "Lang"  # 'E' and the l* family


def chain_names_plain(table, which_ptl, layer_idices, abbres):
    return {table[a](which_ptl, l) for a in abbres for l in layer_idices}


def chain_names_modal(table, which_ptl, layer_idices, abbres):
    """Robust variant: (names, name_in_modal, name_in_module, name_in_layer)."""
    names, in_modal, in_module, in_layer = set(), {}, {}, {}
    for a in abbres:
        modal = lxmert_modality(a)
        for l in layer_idices:
            n = table[a](which_ptl, l)
            names.add(n)
            in_module[n] = a
            in_modal[n] = modal
            # reference :89-92 compares the MODALITY with ['P','E','VV','VB'], so only 'P' matches
            in_layer[n] = modal if modal in ("P", "E", "VV", "VB") else f"{modal}_{l}"
    return names, in_modal, in_module, in_layer


# --------------------------------------------------------------------------- binarisers
def reshape_mask_for_sp(mask, structured_mask_expanding, name="weight"):
    """Expand a per-head mask to the weight / bias shape (structured pruning helper, maskers.py:292-307)."""
    if structured_mask_expanding is not None:
        _mask = (mask.unsqueeze(1) * structured_mask_expanding).view(-1)
        if name == "weight":
            mask = _mask.unsqueeze(1)
        elif name == "bias":
            mask = _mask.unsqueeze(0).unsqueeze(0)
        else:
            raise NotImplementedError("not supported mask type.")
    return mask


def binarizer_fn1(inputs, threshold):
    """1.0 where inputs > threshold (strict), 0.0 elsewhere -- CUDA kernel (reference maskers.py:325-329)."""
    return ops.binarize(inputs, threshold)


class _Binarizer1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs, threshold):
        return binarizer_fn1(inputs, threshold)

    @staticmethod
    def backward(ctx, gradOutput):
        return gradOutput, None  # straight-through (reference returns a third, ignored None)


def binarizer_fn2(inputs):
    outputs = inputs.clone()
    inputs.data.clamp_(-1, 1)
    outputs.data = (torch.sign(outputs.data) + 1) / 2
    return outputs


class _Binarizer2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs):
        ctx.save_for_backward(inputs)
        return binarizer_fn2(inputs)

    @staticmethod
    def backward(ctx, gradOutput):
        (inputs,) = ctx.saved_tensors
        g = gradOutput.clone()
        g[inputs.ge(1)] = 0
        g[inputs.le(-1)] = 0
        return g


def binarizer_fn3(inputs):
    return torch.bernoulli(torch.sigmoid(inputs))


class _Binarizer3(torch.autograd.Function):
    @staticmethod
    def forward(ctx, inputs):
        return binarizer_fn3(inputs)

    @staticmethod
    def backward(ctx, gradOutput):
        return gradOutput


_scheme_idx_to_fn = {"MaskedLinear1": _Binarizer1, "MaskedLinear2": _Binarizer2, "MaskedLinear3": _Binarizer3}


def _get_nnz_from(tensor):
    return tensor.char().sum().detach().cpu().numpy()


# --------------------------------------------------------------------------- masked modules
class MaskedLinearX(nn.Module):
    """Holds (weight, bias) of the wrapped layer, the real-valued score ``weight_mask`` and the
    binarisation ``threshold`` (reference maskers.py:85-284)."""

    def __init__(self, scheme_idx, weight, bias, mask_biases, **kwargs):
        super().__init__()
        init_scale = kwargs.get("init_scale")
        init_sparsity = kwargs.get("init_sparsity")
        self.name = kwargs.get("name")
        self.padding_idx = kwargs.get("padding_idx")
        self.threshold = kwargs.get("threshold")
        self.threshold_fn = _scheme_idx_to_fn[scheme_idx].apply
        self.mask_biases = mask_biases
        self.weight = weight
        self.bias = bias
        self.structured_mask_expanding = None
        self._controlled_init = kwargs.get("controlled_init", False)
        self._init_sparsity = init_sparsity
        defer = bool(kwargs.get("_defer_magnitude_init", False))

        assert "structured_masking_info" in kwargs
        info = kwargs["structured_masking_info"]
        structured_masking = info["structured_masking"]
        types = info["structured_masking_types"]
        self.force_masking = info["force_masking"]
        self.is_structured_masking = structured_masking is not None and structured_masking != "None"
        self.meet_structured_masking_cond = types is None or any(t in self.name for t in types)

        init_scales = self.get_init_scales(scheme_idx, init_sparsity, init_scale)

        self.structured_masked = False
        if self.is_structured_masking and self.meet_structured_masking_cond:
            self.structured_masked = True
            if structured_masking == "layers":
                tmpl = torch.empty(1).uniform_(*init_scales)
            elif structured_masking == "heads":
                assert "self" in self.name
                conf = info["ptl_config"]
                heads = conf.num_attention_heads
                tmpl = torch.empty(heads).uniform_(*init_scales)
                self.structured_mask_expanding = nn.Parameter(
                    torch.ones(heads, conf.hidden_size // heads), requires_grad=False)
            else:
                raise NotImplementedError(f"structured_masking={structured_masking} not supported yet")
            self.weight_mask = nn.Parameter(tmpl.clone())
            if mask_biases:
                self.bias_mask = nn.Parameter(tmpl.clone())

        self.unstructured_masked = False
        self._pending_magnitude = False
        if not self.structured_masked and (not self.is_structured_masking or self.force_masking in self.name):
            self.unstructured_masked = True
            if self._controlled_init is None:
                self.weight_mask = nn.Parameter(torch.empty_like(self.weight).uniform_(*init_scales))
                if mask_biases:
                    self.bias_mask = nn.Parameter(torch.empty_like(self.bias).uniform_(*init_scales))
            elif defer and self._uses_magnitude(self._controlled_init) and not mask_biases:
                # Masker batches the |W| selects of all modules into one launch (finish_magnitude_init)
                self.weight_mask = nn.Parameter(torch.empty_like(self.weight))
                self._pending_magnitude = True
            else:
                self.weight_mask = self.controlled_init(self.weight, init_sparsity, self.threshold,
                                                        controlled_init_type=self._controlled_init)
                if mask_biases:
                    self.bias_mask = self.controlled_init(self.bias, init_sparsity, self.threshold,
                                                          controlled_init_type=self._controlled_init)

    def _uses_magnitude(self, kind):
        return kind == "magnitude" or (kind == "magnitude_and_uniform" and "bert" in self.name)

    @staticmethod
    def num_zero_elements(weight, init_sparsity):
        return int(weight.nelement() * init_sparsity)  # Python double arithmetic, as the reference

    def controlled_init(self, weight, init_sparsity, threshold, controlled_init_type):
        """Score initialisation (reference maskers.py:199-271)."""
        n = weight.nelement()
        k = self.num_zero_elements(weight, init_sparsity)
        thr = float(threshold)

        def _magnitude():
            w_thr = ops.kth_value_batched([weight.detach()], [k], use_abs=True)
            return ops.magnitude_init(weight, w_thr[0:1], 2.0 * thr, 0.0 * thr)

        def _uniform():
            s = torch.zeros_like(weight.view(-1))
            picked = np.random.choice(np.arange(n), size=k, replace=False)
            keep = torch.ones_like(s)
            keep[picked] = 0
            keep = keep.bool()
            s[keep] = 2.0 * thr
            s[~keep] = 0.0 * thr
            return s.view(*weight.size())

        def _double_uniform():
            s = torch.zeros_like(weight.view(-1))
            picked = np.random.choice(np.arange(n), size=k)
            keep = torch.ones_like(s)
            keep[picked] = 0
            keep = keep.bool()
            above = s.clone().uniform_(1.1 * thr, 1.5 * thr).mul_(keep)
            below = s.clone().uniform_(0.5 * thr, 0.9 * thr).mul_(~keep)
            return (above + below).view(*weight.size())

        if controlled_init_type == "magnitude":
            s = _magnitude()
        elif controlled_init_type == "uniform":
            s = _uniform()
        elif controlled_init_type == "magnitude_and_uniform":
            s = _magnitude() if "bert" in self.name else _uniform()
        elif controlled_init_type == "double_uniform":
            s = _double_uniform()
        else:
            raise NotImplementedError("this controlled init type is not supported.")
        return nn.Parameter(s)

    def get_init_scales(self, scheme_idx, init_sparsity, init_scale):
        if init_scale is None:
            # maskers_Robust.Masker.replace passes no init_scale (reference maskers_Robust.py:599-612) and
            # the reference then dies in `None + tensor`; the scales are unused with a controlled init.
            return (0.0, 0.0)
        if scheme_idx == "MaskedLinear1":
            top = init_scale + float(self.threshold)   # the reference divides a tensor: sparsity 0 gives inf, no error
            return (-init_scale, (top / init_sparsity if init_sparsity else math.copysign(math.inf, top)) - init_scale)
        if scheme_idx == "MaskedLinear2":
            warnings.warn(f"we cannot control the initial sparsity for {scheme_idx}.")
            return (-init_scale, init_scale)
        if scheme_idx == "MaskedLinear3":
            p = 1 - init_sparsity
            i_s = math.log(p / (1 - p))
            return (i_s, i_s)
        return (-init_scale, init_scale)

    def forward(self, x):
        raise NotImplementedError


def global_kth_value(tensors, k, use_abs=False):
    """k-th smallest over the UNION of the tensors (reference: torch.cat([...]).kthvalue(k), global_maskers.py:536-541,
    global_mask_trainer_VQA.py:424-429): one contiguous copy, then a single-segment exact select (the
    sample / filter path reads it once)."""
    flat = torch.cat([ops._stage(t.detach()).reshape(-1) for t in tensors])
    return ops.kth_value_batched([flat], [int(k)], use_abs=use_abs)


def finish_magnitude_init(modules, global_sparsity=None):
    """One batched exact select over |W| of every pending module, then S = 2*thr where |W| > kth else 0.
    With `global_sparsity` (global_maskers.Masker, global_prune=True) ONE magnitude cut is taken over the union of
    all pending weights instead (reference global_maskers.py:219-231,531-541); returns that cut (0-dim tensor)."""
    pend = [m for m in modules if getattr(m, "_pending_magnitude", False)]
    if not pend:
        return None
    cut = None
    if global_sparsity is not None:
        total = sum(m.weight.numel() for m in pend)
        cut = global_kth_value([m.weight for m in pend], int(total * global_sparsity), use_abs=True)
        w_thr = cut.expand(len(pend))
    else:
        ks = [MaskedLinearX.num_zero_elements(m.weight, m._init_sparsity) for m in pend]
        w_thr = ops.kth_value_batched([m.weight.detach() for m in pend], ks, use_abs=True)
    for i, m in enumerate(pend):
        thr = float(m.threshold)
        m.weight_mask.data = ops.magnitude_init(m.weight, w_thr[i:i + 1].contiguous(), 2.0 * thr, 0.0 * thr)
        m._pending_magnitude = False
    return cut[0] if cut is not None else None


class MaskedLinear0(nn.Module):
    """Scheme 0: no mask."""

    def __init__(self, weight, bias, **kwargs):
        super().__init__()
        self.weight = weight
        self.bias = bias

    def forward(self, x):
        return F.linear(x, self.weight, self.bias)


class MaskedLinear1(MaskedLinearX):
    """Scheme 1 (the only one any shipped script uses): threshold binariser with straight-through
    gradient.  forward == F.linear(x, weight * (weight_mask > threshold), bias), computed by the fused
    sm_100a masked GEMM (reference maskers.py:342-366)."""

    def __init__(self, weight, bias, mask_biases, **kwargs):
        super().__init__("MaskedLinear1", weight, bias, mask_biases, **kwargs)
        self._w16 = None
        self._w16_key = None
        self._thr_dev = None
        self._thr_key = None

    # -- cached device-side operands ------------------------------------------------------------
    def _weight_bf16(self):
        w = self.weight
        key = (w.data_ptr(), w._version, w.device)
        if self._w16 is None or self._w16_key != key:
            self._w16 = ops.to_bf16(w.detach())
            self._w16_key = key
        return self._w16

    def _threshold_on(self, device):
        t = self.threshold
        if torch.is_tensor(t) and t.device == device and t.dtype == torch.float32:
            return t
        key = (id(t), t._version if torch.is_tensor(t) else t, device)
        if self._thr_dev is None or self._thr_key != key:
            self._thr_dev = ops.as_thr(t, device)
            self._thr_key = key
        return self._thr_dev

    def get_masks(self):
        thr = self._threshold_on(self.weight_mask.device)
        M_w = self.threshold_fn(self.weight_mask, thr)
        M_w = reshape_mask_for_sp(M_w, self.structured_mask_expanding, name="weight")
        if self.mask_biases:
            M_b = self.threshold_fn(self.bias_mask, thr)
            M_b = reshape_mask_for_sp(M_b, self.structured_mask_expanding, name="bias")
        else:
            M_b = None
        return M_w, M_b

    def forward(self, x):
        if self.structured_masked or self.mask_biases:
            # structured / bias masking is never enabled by the shipped scripts (drivers assert
            # --structured false); keep the reference formula on top of the CUDA binariser.
            M_w, M_b = self.get_masks()
            if "embedding" in self.name:
                return F.embedding(x, self.weight * M_w, padding_idx=self.padding_idx)
            return F.linear(x, self.weight * M_w, self.bias * M_b if M_b is not None else self.bias)
        thr = self._threshold_on(self.weight_mask.device)
        sink = self if getattr(self, "_arena_grad", None) is not None else None
        if sink is not None and torch.is_grad_enabled() and getattr(self, "_sync", None) is not None:
            self._calls_outstanding = getattr(self, "_calls_outstanding", 0) + 1  # one more backward is owed
        if "embedding" in self.name:
            return ops.MaskedEmbeddingFn.apply(x, self.weight_mask, self.weight, thr, self.padding_idx, sink)
        if self.weight.shape[1] % 8 != 0:
            return ops.MaskedLinearSmallKFn.apply(x, self.weight_mask, self.weight, thr, self.bias, sink)
        arena = getattr(self, "_arena", None)
        wm = arena.cached_masked_weight(self) if arena is not None else self._held_masked_weight(thr)
        if arena is not None and wm is not None:
            arena.wait_ready(self)       # sharded optimiser: this module's operand all-gather has landed
        w32 = self.weight if self.weight.dtype == torch.float32 and self.weight.is_contiguous() else None
        return ops.MaskedLinearFn.apply(x, self.weight_mask, self._weight_bf16(), thr, self.bias, sink, wm, w32)

    # -- per-module mask cache for engines without a score arena (mPLUG/engine.py) --------------------------------
    def hold_masked_weight(self, on=True):
        """Opt in: keep W (.) M as a bf16 operand between calls, so forward and dX run as plain GEMMs.  Whoever changes
        the scores (the engine, after its optimiser step) must call drop_masked_weight(); a new threshold object or
        an in-place threshold update is noticed here."""
        self._hold_wm = bool(on)
        self.drop_masked_weight()

    def drop_masked_weight(self):
        self._wm = None
        self._wm_key = None

    def _wm_key_now(self, thr):
        t = self.threshold               # the source object: alive as long as it is current, so its id is not reused
        return (id(t), t._version if torch.is_tensor(t) else t, getattr(self, "score_dtype", None),
                self.weight_mask.data_ptr(), self.weight_mask._version, thr.device)   # optimisers bump _version

    def _held_masked_weight(self, thr):
        if not getattr(self, "_hold_wm", False):
            return None
        key = self._wm_key_now(thr)
        if self._wm is None or self._wm_key != key:
            self._wm = ops.apply_mask_bf16(self._weight_bf16(), self.weight_mask.detach(), thr)
            self._wm_key = key
        return self._wm

    def holds_masked_operand(self):
        """True when forward() takes its operand from _held_masked_weight (the tensor-core path of a held module)."""
        return (getattr(self, "_hold_wm", False) and not (self.structured_masked or self.mask_biases)
                and "embedding" not in self.name and self.weight.shape[1] % 8 == 0
                and getattr(self, "_arena", None) is None)

    def adopt_masked_weight(self, wm, thr):
        """An engine whose optimiser pass wrote W (.) (S_new > thr) into `wm` (crv_adamw_multi) hands it over: valid for
        the current scores / threshold, exactly as if _held_masked_weight had just built it."""
        self._wm = wm
        self._wm_key = self._wm_key_now(thr)


class MaskedLinear2(MaskedLinearX):
    def __init__(self, weight, bias, mask_biases, **kwargs):
        super().__init__("MaskedLinear2", weight, bias, mask_biases, **kwargs)

    def get_masks(self):
        M_w = reshape_mask_for_sp(self.threshold_fn(self.weight_mask), self.structured_mask_expanding, name="weight")
        M_b = None
        if self.mask_biases:
            M_b = reshape_mask_for_sp(self.threshold_fn(self.bias_mask), self.structured_mask_expanding, name="bias")
        return M_w, M_b

    def forward(self, x):
        M_w, M_b = self.get_masks()
        return F.linear(x, self.weight * M_w, self.bias * M_b if M_b is not None else self.bias)


class MaskedLinear3(MaskedLinear2):
    def __init__(self, weight, bias, mask_biases, **kwargs):
        MaskedLinearX.__init__(self, "MaskedLinear3", weight, bias, mask_biases, **kwargs)


_MASKED_CLASSES = {"MaskedLinear0": MaskedLinear0, "MaskedLinear1": MaskedLinear1,
                   "MaskedLinear2": MaskedLinear2, "MaskedLinear3": MaskedLinear3}


# --------------------------------------------------------------------------- the Masker
class MaskerBase(object):
    """Swaps the named nn.Linear / nn.Embedding modules for masked ones (reference maskers.py:478-620)."""

    per_modal = False  # Robust variant: per-module initial sparsity from hpmodel.zerorate_dict

    def _setup(self, masker_scheduler, logger, mask_biases, structured_masking_info, threshold, init_scale,
               which_ptl, controlled_init, hpmodel=None):
        self.hpmodel = hpmodel
        self.masker_scheduler = masker_scheduler
        self.mask_biases = mask_biases
        self.structured_masking_info = structured_masking_info
        self.logger = logger
        self.which_ptl = which_ptl
        self.threshold = torch.tensor(threshold)
        self.init_scale = init_scale
        self.controlled_init = controlled_init
        self.names_tobe_masked = None
        self.name_in_module = None
        self.name_of_masker = None
        self.init_masks = {}

    def patch_modules(self, model, names_tobe_masked, name_of_masker="MaskedLinear1"):
        self.ptl_config = getattr(model, self.which_ptl).config
        masked_linear_cls = _MASKED_CLASSES[name_of_masker]
        self._created = []
        self.replace(model, "", names_tobe_masked, masked_linear_cls)
        if getattr(self, "global_prune", False):
            self.global_threshold = finish_magnitude_init(self._created, self.masker_scheduler.init_sparsity)
        else:
            finish_magnitude_init(self._created)
        self.masked_linear_cls = masked_linear_cls

        self.logger.info("Check the trainable status.")
        for _name, param in model.named_parameters():
            self.logger.info(f"\t {_name} is {'trainable' if param.requires_grad else 'not trainable'}.")

        self.logger.info("Check the masking status.")
        for m_name, m in model.named_modules():
            if m_name not in names_tobe_masked:
                continue
            if isinstance(m, masked_linear_cls):
                param_info = {}
                for _name, param in m.named_parameters():
                    if "mask" in _name:
                        mask = self.eval_binarizer_fn(name_of_masker, param.detach(), self.threshold)
                        param_info[_name] = 1.0 - (_get_nnz_from(mask) / np.prod(param.shape))
                        self.init_masks[f"{m_name}_{_name}"] = mask.cpu()
                self.logger.info(f"\t {m_name} is MASKED -> {json.dumps(param_info)}")
            else:
                self.logger.info(f"\t {m_name} is NOT MASKED")

    @staticmethod
    def eval_binarizer_fn(name_of_masker, param, threshold):
        if name_of_masker == "MaskedLinear1":
            return binarizer_fn1(param, threshold)
        if name_of_masker == "MaskedLinear2":
            return binarizer_fn2(param)
        if name_of_masker == "MaskedLinear3":
            return binarizer_fn3(param)
        raise NotImplementedError(f"incorrect name_of_masker={name_of_masker}.")

    def _init_sparsity_for(self, name):
        if self.per_modal:
            return self.hpmodel.zerorate_dict[self.name_in_module[name]]
        return self.masker_scheduler.init_sparsity

    def replace(self, m, root_name, names_tobe_masked, masked_linear_cls, **_ignored):
        # The reference walks dir(m) (maskers.py:560), which is how nn.Sequential children such as the
        # classifier's '0' / '3' escape the freeze; keep that traversal.
        for attr_str in dir(m):
            try:
                target_attr = getattr(m, attr_str)
            except Exception:
                continue
            if not isinstance(target_attr, nn.Module):
                continue
            name = root_name + "." + attr_str if root_name else attr_str
            frozen_scope = not ("classifier" in name or "lm_head" in name)
            if frozen_scope:
                for pname in ("weight", "bias"):
                    p = getattr(target_attr, pname, None)
                    if isinstance(p, torch.Tensor):
                        p.requires_grad = False
            if type(target_attr) not in (nn.Linear, nn.Embedding):
                continue
            masked = False
            if name in names_tobe_masked:
                kwargs = dict(
                    name=name, weight=target_attr.weight, bias=getattr(target_attr, "bias", None),
                    padding_idx=getattr(target_attr, "padding_idx", None), mask_biases=self.mask_biases,
                    threshold=self.threshold, init_sparsity=self._init_sparsity_for(name),
                    controlled_init=self.controlled_init,
                    structured_masking_info={"ptl_config": self.ptl_config, **self.structured_masking_info},
                    _defer_magnitude_init=True)
                if not self.per_modal:
                    kwargs["init_scale"] = self.init_scale
                masked_linear = masked_linear_cls(**kwargs)
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
