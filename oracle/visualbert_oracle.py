"""ORACLE (test infrastructure only -- see oracle/__init__.py): CPU restatement of the VisualBERT stage-2 forward
around the masked call sites (BASELINE config 3).

Functional, parameters keyed like the reference's state_dict.  Follows hg_transformers/modeling_visualbert.py:
:77-215 (embeddings: word + token-type + position for the text, visual_projection + visual token-type (ids = 1) +
visual position (ids = 0) for the regions, concatenated, LayerNorm, dropout), the BERT layer stack (self-attention,
intermediate GELU, output) -- the same sub-module names as LXMERT's single-modality layers, so the layer code of
lxmert_oracle is reused --, the pooler (:505-517), and VisualBertForMultipleChoice (:1021-1174): dropout, the
weight-normed SimpleClassifier `cls`, CrossEntropyLoss against the soft label distribution.
Masked modules (masking/maskers_visualBert.py:33-60): K,Q,V,AO,I,O of every layer, the pooler and the word embeddings;
`visual_projection` stays dense.  Pinned by tests/golden/visualbert_tiny.pt (make_golden_visualbert.py).
"""
import torch
import torch.nn.functional as F

from . import lxmert_oracle as lxo
from . import masked_ops as _ops


def module_names(num_layers):
    """Masked modules in named_modules order."""
    out = ["visual_bert.embeddings.word_embeddings"]
    for l in range(num_layers):
        out += [f"visual_bert.encoder.layer.{l}.{s}" for s in (
            "attention.self.query", "attention.self.key", "attention.self.value", "attention.output.dense",
            "intermediate.dense", "output.dense")]
    return out + ["visual_bert.pooler.dense"]


def classifier(c, pooled):
    def wn(pre):
        v, g = c.P[pre + ".weight_v"], c.P[pre + ".weight_g"]
        return v * (g / v.norm())
    h = F.relu(F.linear(pooled, wn("cls.main.0"), c.P["cls.main.0.bias"]))
    h = c.drop(h, c.p_cls)
    return F.linear(h, wn("cls.main.3"), c.P["cls.main.3.bias"])


def forward(c, ids, feats, num_layers, pad_token_id=1):
    """(logits, pooled) of VisualBertForMultipleChoice; `c` is an lxmert_oracle.Ctx (eps of its LayerNorms must be the
    config's layer_norm_eps = 1e-12, which is lxmert_oracle.LN_EPS)."""
    B, T = ids.shape
    e = "visual_bert.embeddings."
    name = e + "word_embeddings"
    if name in c.S:
        words = _ops.masked_embedding(ids, c.S[name], c.P[name + ".weight"], c.T[name], pad_token_id)
    else:
        words = F.embedding(ids, c.P[name + ".weight"], padding_idx=pad_token_id)
    text = words + c.P[e + "token_type_embeddings.weight"][0].view(1, 1, -1) + c.P[e + "position_embeddings.weight"][:T].unsqueeze(0)
    vis = (F.linear(feats, c.P[e + "visual_projection.weight"], c.P[e + "visual_projection.bias"])
           + c.P[e + "visual_position_embeddings.weight"][0].view(1, 1, -1)
           + c.P[e + "visual_token_type_embeddings.weight"][1].view(1, 1, -1))
    x = c.drop(c.ln(e + "LayerNorm", torch.cat((text, vis), dim=1)), c.p_hidden)
    for l in range(num_layers):
        x = lxo._layer(c, f"visual_bert.encoder.layer.{l}", x)
    pooled = torch.tanh(c.lin("visual_bert.pooler.dense", x[:, 0]))
    return classifier(c, c.drop(pooled, c.p_hidden)), pooled


def soft_cross_entropy(logits, target):
    """nn.CrossEntropyLoss()(logits, target) with class-probability targets: mean_b sum_a -t log_softmax."""
    return -(target * F.log_softmax(logits, dim=-1)).sum(-1).mean()
