"""The three BERT stacks of mPLUG-VQA on the training path (reference mPLUG/models/modeling_mplug.py): the text
encoder (``BertModel``), the skip-connected fusion encoder (``FusionModel``: :445-519 layer, :600-685 encoder) and the
answer decoder (``BertLMHeadModel``: :1804-1958).  Same module tree and parameter names as the reference, so its
checkpoints load with ``strict=True`` and ``masking.maskers.chain_module_names`` finds every Linear by name.

Everything here is plain torch; the Linear layers named by the masker become sm_100a masked GEMMs when
``Masker.patch_modules`` swaps them.  Attention uses ``F.scaled_dot_product_attention`` (the same softmax(QK^T/sqrt(d)
+ mask)V with dropout on the probabilities).  Not built: generation caches (``past_key_values``), head pruning,
relative position embeddings, the TF checkpoint loader and the other task heads of the reference file.
"""
import json
import os
from types import SimpleNamespace

import torch
import torch.nn.functional as F
from torch import nn


# The reference trains under DeepSpeed bf16: every attention there is a bf16 computation.  On a GPU the attention cores
# (not the projections) therefore run in bf16 through the fused SDPA kernels; fp32 attention at 577 image tokens was
# half of the step (profiles/r01_mplug_profile_fp32_attention.txt).  CRVQA_MPLUG_BF16_ATTENTION=0 keeps fp32.
BF16_ATTENTION = os.environ.get("CRVQA_MPLUG_BF16_ATTENTION", "1") != "0"


class BertConfig:
    """The fields of configs/config_bert*.json that the three stacks read (a plain attribute bag)."""

    _DEFAULTS = dict(vocab_size=30522, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                     intermediate_size=3072, hidden_act="gelu", hidden_dropout_prob=0.1,
                     attention_probs_dropout_prob=0.1, max_position_embeddings=512, type_vocab_size=2,
                     initializer_range=0.02, layer_norm_eps=1e-12, pad_token_id=0, encoder_width=768,
                     add_cross_attention=False, fusion_layers=6, stride_layer=100, text_encoder_layers=6,
                     text_decode_layers=12, tie_word_embeddings=True)

    def __init__(self, **kwargs):
        self.__dict__.update(self._DEFAULTS)
        self.__dict__.update(kwargs)

    @classmethod
    def from_json_file(cls, path):
        with open(path) as f:
            return cls(**json.load(f))

    def to_dict(self):
        return dict(self.__dict__)


def _fused_tail(dense_out, input_tensor, ln, dropout, owner):
    """LayerNorm(dropout(dense_out) + input_tensor) of BertSelfOutput / BertOutput in ONE pass (crv_ln_fwd / crv_ln_bwd,
    the kernels of the LXMERT layer path) when the dense output is bf16 on a GPU -- the engine's bf16-activation mode --
    and the LayerNorm is frozen (mPLUG's masker freezes everything but the scores and the LM head).  Returns None when
    the PyTorch ops have to run (CPU, fp32 activations, trainable LayerNorm, odd widths; CRVQA_MPLUG_FUSED=0)."""
    if (not dense_out.is_cuda or dense_out.dtype != torch.bfloat16 or os.environ.get("CRVQA_MPLUG_FUSED", "1") == "0"
            or ln.weight.requires_grad or ln.bias.requires_grad):
        return None
    H = dense_out.shape[-1]
    if H % 128 or H > 1024 or tuple(ln.normalized_shape) != (H,):
        return None
    from crvqa import fused
    if not hasattr(owner, "_site"):
        owner._site = fused.RngState.new_site()
    res = input_tensor if input_tensor.dtype == torch.float32 else input_tensor.float()
    y32, y16 = fused.drop_add_layernorm(dense_out, res, ln, dropout.p, owner._site, owner.training)
    y32._crv_bf16 = y16          # the next masked GEMM reads this copy instead of casting y32 again
    return y32


def _act(name):
    if callable(name):
        return name
    return {"gelu": F.gelu, "relu": F.relu, "gelu_new": lambda x: F.gelu(x, approximate="tanh")}[name]


class BertEmbeddings(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.word_embeddings = nn.Embedding(config.vocab_size, config.hidden_size, padding_idx=config.pad_token_id)
        self.position_embeddings = nn.Embedding(config.max_position_embeddings, config.hidden_size)
        self.token_type_embeddings = nn.Embedding(config.type_vocab_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)
        self.register_buffer("position_ids", torch.arange(config.max_position_embeddings).expand((1, -1)))

    def forward(self, input_ids):
        T = input_ids.shape[1]
        x = self.word_embeddings(input_ids)
        x = x + self.token_type_embeddings(torch.zeros_like(input_ids))
        x = x + self.position_embeddings(self.position_ids[:, :T])
        return self.dropout(self.LayerNorm(x))


class BertSelfAttention(nn.Module):
    def __init__(self, config, is_cross_attention):
        super().__init__()
        if config.hidden_size % config.num_attention_heads:
            raise ValueError("The hidden size (%d) is not a multiple of the number of attention heads (%d)"
                             % (config.hidden_size, config.num_attention_heads))
        self.num_attention_heads = config.num_attention_heads
        self.attention_head_size = config.hidden_size // config.num_attention_heads
        kv_in = config.encoder_width if is_cross_attention else config.hidden_size
        self.query = nn.Linear(config.hidden_size, config.hidden_size)
        self.key = nn.Linear(kv_in, config.hidden_size)
        self.value = nn.Linear(kv_in, config.hidden_size)
        self.dropout = nn.Dropout(config.attention_probs_dropout_prob)

    def _heads(self, x):
        return x.view(x.shape[0], x.shape[1], self.num_attention_heads, self.attention_head_size).transpose(1, 2)

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None):
        src = hidden_states if encoder_hidden_states is None else encoder_hidden_states
        mask = attention_mask if encoder_hidden_states is None else encoder_attention_mask
        q, k, v = self.query(hidden_states), self.key(src), self.value(src)
        out_dtype = q.dtype
        if q.is_cuda and q.dtype == torch.bfloat16 and os.environ.get("CRVQA_MPLUG_FUSED", "1") != "0":
            # a handful of queries (16 question / 6 answer tokens) against up to ~600 keys: one CTA per (batch, head)
            # reading the projections in place (crv_fq_attention_fwd / _bwd) instead of a 128-query flash tile
            from crvqa import fused
            if fused.few_query_attention_usable(q, k, v, mask, self.num_attention_heads):
                if not hasattr(self, "_site"):
                    self._site = fused.RngState.new_site()
                return fused.few_query_attention(q, k, v, mask, self.num_attention_heads, self.dropout.p, self._site,
                                                 self.training)
        q, k, v = self._heads(q), self._heads(k), self._heads(v)
        if BF16_ATTENTION and q.is_cuda and q.dtype == torch.float32:
            q, k, v = q.to(torch.bfloat16), k.to(torch.bfloat16), v.to(torch.bfloat16)
        if mask is not None:
            mask = mask.to(q.dtype)
        ctx = F.scaled_dot_product_attention(q, k, v, attn_mask=mask,
                                             dropout_p=self.dropout.p if self.training else 0.0)
        return ctx.transpose(1, 2).reshape(hidden_states.shape[0], hidden_states.shape[1], -1).to(out_dtype)


class BertSelfOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        g = self.dense(hidden_states)
        out = _fused_tail(g, input_tensor, self.LayerNorm, self.dropout, self)
        return out if out is not None else self.LayerNorm(self.dropout(g) + input_tensor)


class BertAttention(nn.Module):
    def __init__(self, config, is_cross_attention=False):
        super().__init__()
        self.self = BertSelfAttention(config, is_cross_attention)
        self.output = BertSelfOutput(config)

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None):
        ctx = self.self(hidden_states, attention_mask, encoder_hidden_states, encoder_attention_mask)
        return self.output(ctx, hidden_states)


class BertIntermediate(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.intermediate_size)
        self.intermediate_act_fn = _act(config.hidden_act)

    def forward(self, hidden_states):
        u = self.dense(hidden_states)
        if (u.is_cuda and u.dtype == torch.bfloat16 and self.intermediate_act_fn is F.gelu and u.numel() % 8 == 0
                and os.environ.get("CRVQA_MPLUG_FUSED", "1") != "0"):
            from crvqa import fused
            return fused.gelu_bf16(u)            # erf GELU on bf16, one pass each way (crv_gelu_fwd / crv_gelu_bwd)
        return self.intermediate_act_fn(u)


class BertOutput(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.intermediate_size, config.hidden_size)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)
        self.dropout = nn.Dropout(config.hidden_dropout_prob)

    def forward(self, hidden_states, input_tensor):
        g = self.dense(hidden_states)
        out = _fused_tail(g, input_tensor, self.LayerNorm, self.dropout, self)
        return out if out is not None else self.LayerNorm(self.dropout(g) + input_tensor)


class BertLayer(nn.Module):
    """Self-attention, optional cross-attention (``config.add_cross_attention``), feed-forward."""

    def __init__(self, config, layer_num):
        super().__init__()
        self.attention = BertAttention(config)
        self.has_cross_attention = bool(getattr(config, "add_cross_attention", False))
        if self.has_cross_attention:
            self.crossattention = BertAttention(config, is_cross_attention=True)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None):
        x = self.attention(hidden_states, attention_mask)
        if self.has_cross_attention:
            assert encoder_hidden_states is not None, "encoder_hidden_states must be given for cross-attention layers"
            x = self.crossattention(x, attention_mask, encoder_hidden_states, encoder_attention_mask)
        return self.output(self.intermediate(x), x)


class FusionLayer(nn.Module):
    """One layer of the skip-connected fusion network.  Ordinary layers: text self-attention, then text->image
    cross-attention.  Every ``stride_layer``-th layer (counted from the first fusion layer, the first excluded) is a
    *connected* layer instead: one self-attention over the concatenation [image; text], no cross-attention."""

    def __init__(self, config, layer_num):
        super().__init__()
        self.stride_layer = getattr(config, "stride_layer", 100)
        self.attention = BertAttention(config)
        self.crossattention = BertAttention(config, is_cross_attention=True)
        self.intermediate = BertIntermediate(config)
        self.output = BertOutput(config)

    def forward(self, hidden_states, attention_mask, encoder_hidden_states, encoder_attention_mask, layer_nums):
        if layer_nums == 0 or layer_nums % self.stride_layer != 0:
            x = self.attention(hidden_states, attention_mask)
            x = self.crossattention(x, attention_mask, encoder_hidden_states, encoder_attention_mask)
        else:
            x = self.attention(torch.cat([encoder_hidden_states, hidden_states], 1),
                               torch.cat([encoder_attention_mask, attention_mask], 3))
        return self.output(self.intermediate(x), x)


class BertEncoder(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.layer = nn.ModuleList([BertLayer(config, i) for i in range(config.num_hidden_layers)])

    def forward(self, hidden_states, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None):
        for layer in self.layer:
            hidden_states = layer(hidden_states, attention_mask, encoder_hidden_states, encoder_attention_mask)
        return hidden_states


class FusionEncoder(nn.Module):
    """Runs layers ``start_layer .. num_hidden_layers-1`` only (the first ones exist for checkpoint compatibility and
    never run).  After a connected layer the output is split back and the image half is added to the image states."""

    def __init__(self, config):
        super().__init__()
        self.layer = nn.ModuleList([FusionLayer(config, i) for i in range(config.num_hidden_layers)])
        self.start_layer = max(0, config.num_hidden_layers - config.fusion_layers)

    def forward(self, hidden_states, attention_mask, encoder_hidden_states, encoder_attention_mask):
        image_length, text_length = encoder_hidden_states.shape[1], hidden_states.shape[1]
        for i in range(self.start_layer, len(self.layer)):
            hidden_states = self.layer[i](hidden_states, attention_mask, encoder_hidden_states,
                                          encoder_attention_mask, i - self.start_layer)
            if hidden_states.shape[1] == image_length + text_length:
                image_new, hidden_states = torch.split(hidden_states, (image_length, text_length), 1)
                encoder_hidden_states = encoder_hidden_states + image_new
        return [encoder_hidden_states, hidden_states]


class BertPooler(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)

    def forward(self, hidden_states):
        return torch.tanh(self.dense(hidden_states[:, 0]))


class BertPreTrainedModel(nn.Module):
    """Weight initialisation (:879-888) and the additive attention masks (:1040-1093)."""

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

    @property
    def dtype(self):
        return next(p for p in self.parameters() if p.is_floating_point()).dtype

    def get_extended_attention_mask(self, attention_mask, input_shape, device, is_decoder):
        """[B, T] (1 = attend) -> additive [B, 1, 1|T, T]: 0 where attended, -10000 elsewhere; causal for a decoder."""
        if attention_mask.dim() == 3:
            ext = attention_mask[:, None, :, :]
        elif attention_mask.dim() == 2:
            if is_decoder:
                B, T = input_shape
                ids = torch.arange(T, device=device)
                causal = (ids[None, None, :].repeat(B, T, 1) <= ids[None, :, None]).to(attention_mask.dtype)
                ext = causal[:, None, :, :] * attention_mask[:, None, None, :]
            else:
                ext = attention_mask[:, None, None, :]
        else:
            raise ValueError("Wrong shape for input_ids (shape {}) or attention_mask (shape {})".format(
                input_shape, attention_mask.shape))
        return (1.0 - ext.to(dtype=self.dtype)) * -10000.0

    def invert_attention_mask(self, mask):
        ext = mask[:, None, :, :] if mask.dim() == 3 else mask[:, None, None, :]
        return (1.0 - ext.to(dtype=self.dtype)) * -10000.0


class _BertBase(BertPreTrainedModel):
    encoder_cls = BertEncoder

    def __init__(self, config, add_pooling_layer=True):
        super().__init__(config)
        self.embeddings = BertEmbeddings(config)
        self.encoder = self.encoder_cls(config)
        self.pooler = BertPooler(config) if add_pooling_layer else None
        self.init_weights()

    def get_input_embeddings(self):
        return self.embeddings.word_embeddings

    def _prepare(self, input_ids, attention_mask, encoder_embeds, encoder_hidden_states, encoder_attention_mask,
                 is_decoder):
        if input_ids is not None:
            shape, device = input_ids.shape, input_ids.device
        elif encoder_embeds is not None:
            shape, device = encoder_embeds.shape[:-1], encoder_embeds.device
        else:
            raise ValueError("You have to specify either input_ids or inputs_embeds or encoder_embeds")
        if attention_mask is None:
            attention_mask = torch.ones(tuple(shape), device=device)
        ext = self.get_extended_attention_mask(attention_mask, tuple(shape), device, is_decoder)
        enc_ext = None
        if encoder_hidden_states is not None:
            if encoder_attention_mask is None:
                encoder_attention_mask = torch.ones(encoder_hidden_states.shape[:2], device=device)
            enc_ext = self.invert_attention_mask(encoder_attention_mask)
        x = self.embeddings(input_ids) if encoder_embeds is None else encoder_embeds
        return x, ext, enc_ext


class BertModel(_BertBase):
    def forward(self, input_ids=None, attention_mask=None, encoder_embeds=None, encoder_hidden_states=None,
                encoder_attention_mask=None, return_dict=True, is_decoder=False):
        x, ext, enc_ext = self._prepare(input_ids, attention_mask, encoder_embeds, encoder_hidden_states,
                                        encoder_attention_mask, is_decoder)
        out = self.encoder(x, ext, encoder_hidden_states, enc_ext)
        pooled = self.pooler(out) if self.pooler is not None else None
        return SimpleNamespace(last_hidden_state=out, pooler_output=pooled) if return_dict else (out, pooled)


class FusionModel(_BertBase):
    encoder_cls = FusionEncoder

    def forward(self, input_ids=None, attention_mask=None, encoder_embeds=None, encoder_hidden_states=None,
                encoder_attention_mask=None, return_dict=False, is_decoder=False):
        x, ext, enc_ext = self._prepare(input_ids, attention_mask, encoder_embeds, encoder_hidden_states,
                                        encoder_attention_mask, is_decoder)
        return self.encoder(x, ext, encoder_hidden_states, enc_ext)


class BertPredictionHeadTransform(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.dense = nn.Linear(config.hidden_size, config.hidden_size)
        self.transform_act_fn = _act(config.hidden_act)
        self.LayerNorm = nn.LayerNorm(config.hidden_size, eps=config.layer_norm_eps)

    def forward(self, hidden_states):
        return self.LayerNorm(self.transform_act_fn(self.dense(hidden_states)))


class BertLMPredictionHead(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.transform = BertPredictionHeadTransform(config)
        self.decoder = nn.Linear(config.hidden_size, config.vocab_size, bias=False)
        self.bias = nn.Parameter(torch.zeros(config.vocab_size))
        self.decoder.bias = self.bias           # one Parameter under two names, as in the reference (:826-829)

    def forward(self, hidden_states):
        return self.decoder(self.transform(hidden_states))


class BertOnlyMLMHead(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.predictions = BertLMPredictionHead(config)

    def forward(self, sequence_output):
        return self.predictions(sequence_output)


class BertLMHeadModel(BertPreTrainedModel):
    """Causal answer decoder with cross-attention to [image; question] and the LM head.  The output projection shares
    its weight with the word embeddings (``config.tie_word_embeddings``, the transformers default)."""

    def __init__(self, config):
        super().__init__(config)
        self.bert = BertModel(config, add_pooling_layer=False)
        self.cls = BertOnlyMLMHead(config)
        self.init_weights()
        if getattr(config, "tie_word_embeddings", True):
            self.cls.predictions.decoder.weight = self.bert.embeddings.word_embeddings.weight

    def get_output_embeddings(self):
        return self.cls.predictions.decoder

    def forward(self, input_ids=None, attention_mask=None, encoder_hidden_states=None, encoder_attention_mask=None,
                labels=None, return_dict=True, is_decoder=True, reduction="mean", soft_labels=None, alpha=0,
                return_logits=False):
        hidden = self.bert(input_ids, attention_mask=attention_mask, encoder_hidden_states=encoder_hidden_states,
                           encoder_attention_mask=encoder_attention_mask, is_decoder=is_decoder).last_hidden_state
        scores = self.cls(hidden)
        if return_logits:
            return scores[:, :-1, :].contiguous()
        lm_loss = None
        if labels is not None:
            shifted = scores[:, :-1, :].contiguous()
            labels = labels[:, 1:].contiguous()
            lm_loss = F.cross_entropy(shifted.view(-1, self.config.vocab_size), labels.view(-1), reduction=reduction)
            lm_loss = lm_loss.view(scores.size(0), -1).sum(1)
            if soft_labels is not None:
                # log_softmax over dim=1 (the sequence axis) is what the reference computes (:1920)
                distill = -torch.sum(F.log_softmax(shifted, dim=1) * soft_labels, dim=-1)
                lm_loss = (1 - alpha) * lm_loss + alpha * (distill * (labels != -100)).sum(1)
        return SimpleNamespace(loss=lm_loss, logits=scores)
