"""VisualBERT for VQA: a 12-layer BERT over [text tokens ; projected region features].

Reference: hg_transformers/modeling_visualbert.py (embeddings :77-204, attention :208-332, layer
:335-402, encoder :405-468, pooler :471-484, VisualBertModel :687-875, VisualBertForMultipleChoice
:1021-1174).  Module names match the reference so ``maskers_visualBert.chain_module_names`` finds the
74 maskable modules (K,Q,V,AO,I,O x 12 layers + pooler + word embeddings); ``visual_projection``
stays dense and frozen, exactly as in the reference's weight-type list.
"""
import math

import torch
import torch.nn.functional as F
from torch import nn

from .classifier import SimpleClassifier
from .configuration_visualbert import VisualBertConfig, visualBERTConfig  # noqa: F401


class VisualBertEmbeddings(nn.Module):
    def __init__(self, config):
        super().__init__()
        h = config.hidden_size
        self.word_embeddings = nn.Embedding(config.vocab_size, h, padding_idx=config.pad_token_id)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, h)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, h)
        self.LayerNorm = nn.LayerNorm(h, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.register_buffer("position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)))
        self.visual_token_type_embeddings = nn.Embedding(config.type_vocab_size, h)
        self.visual_position_embeddings = nn.Embedding(config.max_position_embeddings, h)
        if config.special_visual_initialize:
            self.visual_token_type_embeddings.weight.data = self.token_type_embeddings.weight.data.clone()
            self.visual_position_embeddings.weight.data = self.position_embeddings.weight.data.clone()
        self.visual_projection = nn.Linear(config.visual_embedding_dim, h)

    def forward(self, input_ids, visual_embeds, token_type_ids=None, visual_token_type_ids=None):
        t = input_ids.size(1)
        if token_type_ids is None:
            token_type_ids = torch.zeros_like(input_ids)
        text = (self.word_embeddings(input_ids) + self.token_type_embeddings(token_type_ids)
                + self.position_embeddings(self.position_ids[:, :t]))
        if visual_token_type_ids is None:
            visual_token_type_ids = torch.ones(visual_embeds.shape[:-1], dtype=torch.long, device=input_ids.device)
        vpos = torch.zeros(visual_embeds.shape[:-1], dtype=torch.long, device=input_ids.device)
        vis = (self.visual_projection(visual_embeds) + self.visual_position_embeddings(vpos)
               + self.visual_token_type_embeddings(visual_token_type_ids))
        return self.dropout(self.LayerNorm(torch.cat((text, vis), dim=1)))


class VisualBertSelfAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        h = config.hidden_size
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = h // config.num_attention_heads
        self.all_head_size = self.head_size = h
        self.query, self.key, self.value = nn.Linear(h, h), nn.Linear(h, h), nn.Linear(h, h)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)

    def _heads(self, x):
        b, s, _ = x.shape
        return x.view(b, s, self.num_attention_heads, self.attention_head_size).permute(0, 2, 1, 3)

    def forward(self, hidden_states, attention_mask=None):
        q, k, v = self._heads(self.query(hidden_states)), self._heads(self.key(hidden_states)), self._heads(self.value(hidden_states))
        scores = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(self.attention_head_size)
        if attention_mask is not None:
            scores = scores + attention_mask
        probs = self.dropout(F.softmax(scores, dim=-1))
        ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).contiguous()
        return ctx.view(ctx.size(0), ctx.size(1), self.all_head_size)


class VisualBertSelfOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        return self.LayerNorm(self.dropout(self.dense(hidden_states)) + input_tensor)


class VisualBertAttention(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.self = VisualBertSelfAttention(config)
        self.output = VisualBertSelfOutput(config)

    def forward(self, hidden_states, attention_mask=None):
        return self.output(self.self(hidden_states, attention_mask), hidden_states)


class VisualBertIntermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)

    def forward(self, hidden_states):
        return F.gelu(self.dense(hidden_states))


class VisualBertOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        return self.LayerNorm(self.dropout(self.dense(hidden_states)) + input_tensor)


class VisualBertLayer(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.attention = VisualBertAttention(config)
        self.intermediate = VisualBertIntermediate(config)
        self.output = VisualBertOutput(config)

    def forward(self, hidden_states, attention_mask=None):
        a = self.attention(hidden_states, attention_mask)
        return self.output(self.intermediate(a), a)


class VisualBertEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.layer = nn.ModuleList([VisualBertLayer(config) for _ in range(config.num_hidden_layers)])

    def forward(self, hidden_states, attention_mask=None):
        fast = self._fast_plans() if hidden_states.is_cuda else None
        if fast is not None:
            # engine fast path (crvqa.fused), the same building blocks as LXMERT's single-modality layers: grouped QKV
            # GEMM, small-sequence attention (20 tokens + 36 regions = 56 <= 64), fused dropout+residual+LayerNorm, GELU
            x32, x16 = hidden_states, hidden_states.to(torch.bfloat16)
            for att, ffn in fast:
                a32, a16 = att.self_attention(x32, x16, attention_mask, self.training)
                x32, x16 = ffn(a32, a16, self.training)
            return x32
        for blk in self.layer:
            hidden_states = blk(hidden_states, attention_mask)
        return hidden_states

    def _fast_plans(self):
        """Layer plans when every masked module of the stack sits in a ScoreArena with a valid mask cache (stage-2
        training engine); None selects the generic per-module path (also with CRVQA_FUSED=0)."""
        import os
        if os.environ.get("CRVQA_FUSED", "1") == "0":
            return None
        from crvqa import fused
        plans = getattr(self, "_plans", None)
        if plans is None:
            plans = [(fused.AttentionPlan(b.attention.self, b.attention.output), fused.FfnPlan(b.intermediate, b.output))
                     for b in self.layer]
            self._plans = plans
        try:
            for att, ffn in plans:
                if not (att.ready() and ffn.ready()):
                    return None
        except AttributeError:      # modules are not MaskedLinear1 (model not patched)
            return None
        return plans


class VisualBertPooler(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.activation = nn.Tanh()

    def forward(self, hidden_states):
        return self.activation(self.dense(hidden_states[:, 0]))


class VisualBertPreTrainedModel(nn.Module):
    config_class = VisualBertConfig
    base_model_prefix = "visual_bert"

    def __init__(self, config):
        super().__init__()
        self.config = config

    def _init_weights(self, module):
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=self.config.initializer_range)
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)
        if isinstance(module, nn.Linear) and module.bias is not None:
            module.bias.data.zero_()

    def init_weights(self):
        self.apply(self._init_weights)


class VisualBertModel(VisualBertPreTrainedModel):
    def __init__(self, config, add_pooling_layer=True):
        super().__init__(config)
        self.embeddings = VisualBertEmbeddings(config)
        self.encoder = VisualBertEncoder(config)
        self.pooler = VisualBertPooler(config) if add_pooling_layer else None
        if config.bypass_transformer:
            raise NotImplementedError("bypass_transformer is not used by the VQA path")
        self.init_weights()

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, visual_embeds=None,
                visual_attention_mask=None, visual_token_type_ids=None, **unused):
        if input_ids is None:
            raise ValueError("You have to specify input_ids")
        if visual_embeds is None:
            raise ValueError("`visual_embeds` can not be None when using a VisualBert Model.")
        ext = None
        if attention_mask is not None or visual_attention_mask is not None:
            am = attention_mask if attention_mask is not None else torch.ones_like(input_ids)
            vm = (visual_attention_mask if visual_attention_mask is not None
                  else torch.ones(visual_embeds.shape[:-1], device=input_ids.device))
            both = torch.cat((am.to(visual_embeds.dtype), vm.to(visual_embeds.dtype)), dim=-1)
            ext = (1.0 - both[:, None, None, :]) * -10000.0
        emb = self.embeddings(input_ids, visual_embeds, token_type_ids, visual_token_type_ids)
        seq = self.encoder(emb, ext)
        return seq, (self.pooler(seq) if self.pooler is not None else None)


class VisualBertForMultipleChoice(VisualBertPreTrainedModel):
    """(loss, logits, pooled) = model(input_ids=..., visual_embeds=..., labels=soft_targets); the loss is
    CrossEntropyLoss against the soft label distribution (reference :1021-1174)."""

    def __init__(self, config):
        super().__init__(config)
        self.visual_bert = VisualBertModel(config)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.cls = SimpleClassifier(in_dim=config.hidden_size, hid_dim=2 * config.hidden_size, out_dim=config.ans_num,
                                    dropout=0.5, norm="weight", act="ReLU")
        self.init_weights()

    def forward(self, input_ids=None, attention_mask=None, token_type_ids=None, visual_embeds=None,
                visual_attention_mask=None, visual_token_type_ids=None, labels=None, **unused):
        _, pooled = self.visual_bert(input_ids, attention_mask=attention_mask, token_type_ids=token_type_ids,
                                     visual_embeds=visual_embeds, visual_attention_mask=visual_attention_mask,
                                     visual_token_type_ids=visual_token_type_ids)
        logits = self.cls(self.dropout(pooled))
        loss = F.cross_entropy(logits, labels) if labels is not None else None
        return loss, logits, pooled
