"""LXMERT for VQA: the module tree whose Linear / Embedding call sites stage 2 masks.

Reference: hg_transformers/modeling_lxmert.py (embeddings :729-767, attention :770-827, layers
:830-1037, encoder :1041-1120, pooler :1123-1135, LxmertModel :1316-1448,
LxmertForMultipleChoice :233-360).  Module and parameter names are kept identical so that
``chain_module_names`` finds the same 168 modules, reference checkpoints load, and the same
``torch.manual_seed`` reproduces the reference's random init (construction order and the two
``init_weights`` passes are mirrored).  Everything between the masked GEMMs is plain PyTorch here
("next" row f3 of SURVEY.md section 8); the Linear/Embedding modules are swapped for CUDA-backed
``MaskedLinear1`` objects by ``masking.maskers*.Masker.patch_modules``.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from .classifier import SimpleClassifier
from .configuration_lxmert import LxmertConfig  # noqa: F401  (re-export)


def _act(name):
    if name == "gelu":
        return F.gelu
    if name == "relu":
        return F.relu
    raise KeyError(name)


class LxmertEmbeddings(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=0)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size, padding_idx=0)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size, padding_idx=0)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, input_ids, token_type_ids=None):
        t = input_ids.size(1)
        pos = torch.arange(t, dtype=torch.long, device=input_ids.device).unsqueeze(0).expand_as(input_ids)
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        e = self.word_embeddings(input_ids) + self.position_embeddings(pos) + self.token_type_embeddings(token_type_ids)
        if e.is_cuda:
            from crvqa import fused
            if fused.ln_avg_drop_usable(e, None, self.LayerNorm, None):      # LayerNorm + dropout in one pass
                if not hasattr(self, "_site"):
                    self._site = fused.RngState.new_site()
                return fused.ln_avg_drop(e, None, self.LayerNorm, None, self.dropout.p, self._site, self.training)
        return self.dropout(self.LayerNorm(e))


class LxmertAttention(nn.Module):
    def __init__(self, config, ctx_dim=None):
        super().__init__()
        h = config.hidden_size
        if h % config.num_attention_heads:
            raise ValueError("hidden size must be a multiple of the number of heads")
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = h // config.num_attention_heads
        self.head_size = h
        ctx_dim = h if ctx_dim is None else ctx_dim
        self.query = nn.Linear(h, h)
        self.key = nn.Linear(ctx_dim, h)
        self.value = nn.Linear(ctx_dim, h)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)

    def _heads(self, x):
        b, s, _ = x.shape
        return x.view(b, s, self.num_attention_heads, self.attention_head_size).permute(0, 2, 1, 3)

    def forward(self, hidden_states, context, attention_mask=None):
        q = self._heads(self.query(hidden_states))
        k = self._heads(self.key(context))
        v = self._heads(self.value(context))
        scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(self.attention_head_size)
        if attention_mask is not None:
            scores = scores + attention_mask
        probs = self.dropout(F.softmax(scores, dim=-1))
        ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).contiguous()
        return ctx.view(ctx.size(0), ctx.size(1), self.head_size)


class LxmertAttentionOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        return self.LayerNorm(self.dropout(self.dense(hidden_states)) + input_tensor)


class LxmertCrossAttentionLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.att = LxmertAttention(config)
        self.output = LxmertAttentionOutput(config)

    def forward(self, input_tensor, ctx_tensor, ctx_att_mask=None):
        return self.output(self.att(input_tensor, ctx_tensor, ctx_att_mask), input_tensor)


class LxmertSelfAttentionLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.self = LxmertAttention(config)
        self.output = LxmertAttentionOutput(config)

    def forward(self, input_tensor, attention_mask=None):
        return self.output(self.self(input_tensor, input_tensor, attention_mask), input_tensor)


class LxmertIntermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)
        self.intermediate_act_fn = _act(config.hidden_act)

    def forward(self, hidden_states):
        return self.intermediate_act_fn(self.dense(hidden_states))


class LxmertOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        return self.LayerNorm(self.dropout(self.dense(hidden_states)) + input_tensor)


class LxmertLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.attention = LxmertSelfAttentionLayer(config)
        self.intermediate = LxmertIntermediate(config)
        self.output = LxmertOutput(config)

    def forward(self, hidden_states, attention_mask=None):
        a = self.attention(hidden_states, attention_mask)
        return self.output(self.intermediate(a), a)


class LxmertXLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.visual_attention = LxmertCrossAttentionLayer(config)
        self.lang_self_att = LxmertSelfAttentionLayer(config)
        self.visn_self_att = LxmertSelfAttentionLayer(config)
        self.lang_inter = LxmertIntermediate(config)
        self.lang_output = LxmertOutput(config)
        self.visn_inter = LxmertIntermediate(config)
        self.visn_output = LxmertOutput(config)

    def forward(self, lang, lang_mask, visn, visn_mask):
        # the SAME cross-attention module serves both directions (reference :947-958), so its four
        # masked Linears are invoked twice per forward and their score gradients add.
        lang_x = self.visual_attention(lang, visn, ctx_att_mask=visn_mask)
        visn_x = self.visual_attention(visn, lang, ctx_att_mask=lang_mask)
        lang_s = self.lang_self_att(lang_x, lang_mask)
        visn_s = self.visn_self_att(visn_x, visn_mask)
        lang_o = self.lang_output(self.lang_inter(lang_s), lang_s)
        visn_o = self.visn_output(self.visn_inter(visn_s), visn_s)
        return lang_o, visn_o


class LxmertVisualFeatureEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.visn_fc = nn.Linear(config.visual_feat_dim, config.hidden_size)
        self.visn_layer_norm = nn.LayerNorm(config.hidden_size, eps=1e-12)
        self.box_fc = nn.Linear(config.visual_pos_dim, config.hidden_size)
        self.box_layer_norm = nn.LayerNorm(config.hidden_size, eps=1e-12)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, visual_feats, visual_pos):
        a, b = self.visn_fc(visual_feats), self.box_fc(visual_pos)
        if a.is_cuda:
            from crvqa import fused
            if fused.ln_avg_drop_usable(a, b, self.visn_layer_norm, self.box_layer_norm):
                # both LayerNorms, the average and the dropout in one pass (crv_ln_avg_drop_fwd)
                if not hasattr(self, "_site"):
                    self._site = fused.RngState.new_site()
                return fused.ln_avg_drop(a, b, self.visn_layer_norm, self.box_layer_norm, self.dropout.p, self._site,
                                         self.training)
        x = self.visn_layer_norm(a)
        y = self.box_layer_norm(b)
        return self.dropout((x + y) / 2)


class LxmertEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.visn_fc = LxmertVisualFeatureEncoder(config)
        self.config = config
        self.num_l_layers, self.num_x_layers, self.num_r_layers = config.l_layers, config.x_layers, config.r_layers
        self.layer = nn.ModuleList([LxmertLayer(config) for _ in range(config.l_layers)])
        self.x_layers = nn.ModuleList([LxmertXLayer(config) for _ in range(config.x_layers)])
        self.r_layers = nn.ModuleList([LxmertLayer(config) for _ in range(config.r_layers)])

    def forward(self, lang, lang_mask, visual_feats, visual_pos, visn_mask=None, need_visn_output=True):
        """`lang` is the language embedding output or a callable producing it.  LxmertModel passes a callable so the
        fast path can create the embedding node AFTER the vision stack: autograd runs later-created nodes first,
        which puts the 94 MB word-embedding gradient before the vision stack in the backward pass instead of at
        its very end, where its all-reduce could overlap nothing (_engine.execution_order)."""
        visn = self.visn_fc(visual_feats, visual_pos)
        fast = self._fast_plans() if visn.is_cuda else None
        if fast is not None:
            return self._forward_fast(fast, lang, lang_mask, visn, visn_mask, need_visn_output)
        if callable(lang):
            lang = lang()
        for blk in self.layer:
            lang = blk(lang, lang_mask)
        for blk in self.r_layers:
            visn = blk(visn, visn_mask)
        for blk in self.x_layers:
            lang, visn = blk(lang, lang_mask, visn, visn_mask)
        return lang, visn

    # -- engine fast path (crvqa.fused): same math, bf16 operands produced by fused kernels ------------
    def _fast_plans(self):
        """Layer plans when every masked module of the encoder stack sits in a ScoreArena with a valid mask
        cache (i.e. under the stage-2 training engine); None selects the generic per-module path."""
        import os
        if os.environ.get("CRVQA_FUSED", "1") == "0":
            return None
        from crvqa import fused
        plans = getattr(self, "_plans", None)
        if plans is None:
            def layer_plan(blk):
                return (fused.AttentionPlan(blk.attention.self, blk.attention.output),
                        fused.FfnPlan(blk.intermediate, blk.output))
            plans = {
                "lang": [layer_plan(b) for b in self.layer],
                "visn": [layer_plan(b) for b in self.r_layers],
                "cross": [(fused.AttentionPlan(b.visual_attention.att, b.visual_attention.output),
                           fused.RngState.new_sites(2), fused.RngState.new_sites(2),
                           fused.AttentionPlan(b.lang_self_att.self, b.lang_self_att.output),
                           fused.AttentionPlan(b.visn_self_att.self, b.visn_self_att.output),
                           fused.FfnPlan(b.lang_inter, b.lang_output), fused.FfnPlan(b.visn_inter, b.visn_output))
                          for b in self.x_layers],
            }
            self._plans = plans
        try:
            for att, ffn in plans["lang"] + plans["visn"]:
                if not (att.ready() and ffn.ready()):
                    return None
            for cross, _, _, ls, vs, lf, vf in plans["cross"]:
                if not all(x.ready() for x in (cross, ls, vs, lf, vf)):
                    return None
        except AttributeError:  # modules are not MaskedLinear1 (model not patched)
            return None
        return plans

    def _forward_fast(self, plans, lang32, lang_mask, visn32, visn_mask, need_visn_output=True):
        """Language and vision stacks run in LOCKSTEP (layer i of both, then the remaining language layers): the two
        are independent until the cross layers, so their GEMMs of one phase share a grouped launch
        (crvqa.fused.self_attention_multi / ffn_multi); inside a cross layer the two modalities are grouped the same
        way.  CRVQA_GROUPED=0 issues every GEMM on its own.  With need_visn_output=False (the VQA head reads the
        pooled LANGUAGE output only, reference :256-360) the vision half of the LAST cross layer is dead code -- the
        reference computes it and throws it away -- and is not executed."""
        from crvqa import fused
        tr = self.training
        visn16 = getattr(visn32, "_crv_bf16", None)       # the fused entry block already produced the bf16 copy
        if visn16 is None:
            visn16 = visn32.to(torch.bfloat16)
        if callable(lang32):
            lang32 = lang32()
        lang16 = getattr(lang32, "_crv_bf16", None)
        if lang16 is None:
            lang16 = lang32.to(torch.bfloat16)
        nl, nv = len(plans["lang"]), len(plans["visn"])
        for i in range(max(nl, nv)):
            items = []
            if i < nl:
                items.append((plans["lang"][i], "l"))
            if i < nv:
                items.append((plans["visn"][i], "v"))
            state = {"l": (lang32, lang16, lang_mask), "v": (visn32, visn16, visn_mask)}
            att = fused.self_attention_multi([(pl[0], *state[k]) for pl, k in items], tr)
            ffn = fused.ffn_multi([(pl[1], a32, a16) for (pl, _), (a32, a16) in zip(items, att)], tr)
            for (_, k), (o32, o16) in zip(items, ffn):
                if k == "l":
                    lang32, lang16 = o32, o16
                else:
                    visn32, visn16 = o32, o16
        for li, (cross, site_l, site_v, ls, vs, lf, vf) in enumerate(plans["cross"]):
            if not need_visn_output and li == len(plans["cross"]) - 1 and fused._grouped_on():
                lx32, lx16 = fused.cross_attention_lang_only(cross, lang32, lang16, visn16, visn_mask, tr, site_l)
                ((ls32, ls16),) = fused.self_attention_multi([(ls, lx32, lx16, lang_mask)], tr)
                ((lang32, lang16),) = fused.ffn_multi([(lf, ls32, ls16)], tr)
                visn32 = None
                break
            (lx32, lx16), (vx32, vx16) = fused.cross_attention_pair(cross, lang32, lang16, visn32, visn16, lang_mask,
                                                                    visn_mask, tr, site_l, site_v)
            (ls32, ls16), (vs32, vs16) = fused.self_attention_multi([(ls, lx32, lx16, lang_mask),
                                                                     (vs, vx32, vx16, visn_mask)], tr)
            (lang32, lang16), (visn32, visn16) = fused.ffn_multi([(lf, ls32, ls16), (vf, vs32, vs16)], tr)
        return lang32, visn32


class LxmertPooler(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.activation = nn.Tanh()

    def forward(self, hidden_states):
        return self.activation(self.dense(hidden_states[:, 0]))


class LxmertPreTrainedModel(nn.Module):
    config_class = LxmertConfig
    base_model_prefix = "lxmert"

    def __init__(self, config):
        super().__init__()
        self.config = config

    def _init_weights(self, module):
        # reference :205-219 -- N(0, initializer_range) weights, zero biases, zero padding row
        if isinstance(module, nn.Linear):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
            if module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.Embedding):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
            if module.padding_idx is not None:
                module.weight.data[module.padding_idx].zero_()
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)

    def init_weights(self):
        self.apply(self._init_weights)

    def resize_token_embeddings(self, new_num_tokens=None):
        emb = getattr(self, self.base_model_prefix, self).embeddings.word_embeddings
        if new_num_tokens is None or new_num_tokens == emb.num_embeddings:
            return emb
        raise NotImplementedError("vocabulary resizing is outside the stage-2 path")

    @property
    def dtype(self):
        return next(self.parameters()).dtype


class LxmertModel(LxmertPreTrainedModel):
    def __init__(self, config):
        super().__init__(config)
        self.embeddings = LxmertEmbeddings(config)
        self.encoder = LxmertEncoder(config)
        self.pooler = LxmertPooler(config)
        self.init_weights()

    def forward(self, input_ids=None, visual_feats=None, visual_pos=None, attention_mask=None,
                visual_attention_mask=None, token_type_ids=None, need_visn_output=True, **unused):
        if input_ids is None:
            raise ValueError("You have to specify input_ids")
        if visual_feats is None:
            raise ValueError("`visual_feats` cannot be `None`")
        if visual_pos is None:
            raise ValueError("`visual_pos` cannot be `None`")
        lang_mask = None
        if attention_mask is not None:
            lang_mask = (1.0 - attention_mask[:, None, None, :].to(visual_feats.dtype)) * -10000.0
        visn_mask = None
        if visual_attention_mask is not None:
            visn_mask = (1.0 - visual_attention_mask[:, None, None, :].to(visual_feats.dtype)) * -10000.0
        lang, visn = self.encoder(lambda: self.embeddings(input_ids, token_type_ids), lang_mask, visual_feats,
                                  visual_pos, visn_mask, need_visn_output)
        return lang, visn, self.pooler(lang)


class LxmertForMultipleChoice(LxmertPreTrainedModel):
    """(loss, logits, pooled) = model(ids, feats, pos, labels=target)  -- reference :233-360."""

    def __init__(self, config):
        super().__init__(config)
        self.lxmert = LxmertModel(config)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.classifier = SimpleClassifier(in_dim=config.hidden_size, hid_dim=2 * config.hidden_size,
                                           out_dim=config.ans_num, dropout=0.5, norm="weight", act="ReLU")
        self.init_weights()

    @staticmethod
    def instance_bce_with_logits(logits, labels, reduction="mean"):
        assert logits.dim() == 2
        loss = F.binary_cross_entropy_with_logits(logits, labels, reduction=reduction)
        if reduction == "mean":
            loss = loss * labels.size(1)
        return loss

    def forward(self, input_ids=None, visual_feats=None, visual_pos=None, attention_mask=None,
                visual_attention_mask=None, token_type_ids=None, labels=None, **unused):
        _, _, pooled = self.lxmert(input_ids=input_ids, visual_feats=visual_feats, visual_pos=visual_pos,
                                   attention_mask=attention_mask, visual_attention_mask=visual_attention_mask,
                                   token_type_ids=token_type_ids, need_visn_output=False)
        logits = self.classifier(pooled)
        loss = self.instance_bce_with_logits(logits, labels) if labels is not None else None
        return loss, logits, pooled
