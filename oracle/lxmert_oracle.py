"""Oracle restatement of the LXMERT stage-2 forward / training step around the masked call sites.

Functional (no nn.Module): parameters come as a dict keyed like the reference's state_dict
(`lxmert.encoder.layer.0.attention.self.query.weight`, ...), scores as {module_name: tensor},
thresholds as {module_name: float}.  Follows hg_transformers/modeling_lxmert.py:729-767 (embeddings),
:770-827 (attention), :830-1037 (layers), :1041-1120 (encoder order: language, then vision, then
cross), :1123-1135 (pooler), :233-360 (head), hg_transformers/classifier.py:5-22 (weight-normed head).
"""
import math

import torch
import torch.nn.functional as F

from . import adamw as _adamw
from . import losses as _losses
from . import masked_ops as _ops

LN_EPS = 1e-12

LXMERT_WEIGHT_TYPES = ["E", "VV", "VB", "lK", "lQ", "lV", "lAO", "lI", "lO", "vK", "vQ", "vV", "vAO", "vI", "vO",
                       "vlVK", "vlVQ", "vlVV", "vlVAO", "vlLaK", "vlLaQ", "vlLaV", "vlLaAO", "vlVaK", "vlVaQ",
                       "vlVaV", "vlVaAO", "vlLi", "vlLo", "vlVi", "vlVo", "P"]


def module_names(l_layers=9, r_layers=5, x_layers=5):
    """The maskable modules that exist, with modality (masking/maskers_Robust.py:24-57,79) -- in the
    order model.named_modules() visits them (embeddings, visn_fc, layer.*, x_layers.*, r_layers.*, pooler)."""
    out = [("lxmert.embeddings.word_embeddings", "Lang"), ("lxmert.encoder.visn_fc.visn_fc", "Vis"),
           ("lxmert.encoder.visn_fc.box_fc", "Vis")]
    att = ["attention.self.query", "attention.self.key", "attention.self.value", "attention.output.dense",
           "intermediate.dense", "output.dense"]
    for l in range(l_layers):
        out += [(f"lxmert.encoder.layer.{l}.{a}", "Lang") for a in att]
    for l in range(x_layers):
        pre = f"lxmert.encoder.x_layers.{l}."
        out += [(pre + s, "Fus") for s in (
            "visual_attention.att.query", "visual_attention.att.key", "visual_attention.att.value",
            "visual_attention.output.dense",
            "lang_self_att.self.query", "lang_self_att.self.key", "lang_self_att.self.value",
            "lang_self_att.output.dense",
            "visn_self_att.self.query", "visn_self_att.self.key", "visn_self_att.self.value",
            "visn_self_att.output.dense",
            "lang_inter.dense", "lang_output.dense", "visn_inter.dense", "visn_output.dense")]
    for l in range(r_layers):
        out += [(f"lxmert.encoder.r_layers.{l}.{a}", "Vis") for a in att]
    out.append(("lxmert.pooler.dense", "P"))
    return out


class Ctx:
    """Everything a forward needs: parameters, scores, thresholds, switches."""

    def __init__(self, params, scores, thresholds, heads=12, operand="fp32", train=False, p_hidden=0.1,
                 p_attn=0.1, p_cls=0.5):
        self.P, self.S, self.T = params, scores, thresholds
        self.heads, self.operand, self.train = heads, operand, train
        self.p_hidden, self.p_attn, self.p_cls = p_hidden, p_attn, p_cls

    def lin(self, name, x):
        w, b = self.P[name + ".weight"], self.P.get(name + ".bias")
        if name in self.S:
            return _ops.masked_linear(x, self.S[name], w, self.T[name], b, self.operand)
        return F.linear(x, w, b)

    def ln(self, name, x):
        return F.layer_norm(x, (x.shape[-1],), self.P[name + ".weight"], self.P[name + ".bias"], LN_EPS)

    def drop(self, x, p):
        return F.dropout(x, p, training=True) if self.train and p > 0 else x


def _attention(c, pre, hidden, context):
    """LxmertAttention.forward -- modeling_lxmert.py:798-827 (no attention mask: all-ones masks add 0)."""
    B, Sq, H = hidden.shape
    d = H // c.heads

    def heads(t):
        return t.view(t.shape[0], t.shape[1], c.heads, d).permute(0, 2, 1, 3)

    q, k, v = heads(c.lin(pre + ".query", hidden)), heads(c.lin(pre + ".key", context)), heads(c.lin(pre + ".value", context))
    probs = F.softmax(torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(d), dim=-1)
    probs = c.drop(probs, c.p_attn)
    ctx = torch.matmul(probs, v).permute(0, 2, 1, 3).contiguous()
    return ctx.view(B, Sq, H)


def _att_out(c, pre, hidden, residual):
    """LxmertAttentionOutput / LxmertOutput -- modeling_lxmert.py:830-841, 889-900."""
    return c.ln(pre + ".LayerNorm", c.drop(c.lin(pre + ".dense", hidden), c.p_hidden) + residual)


def _self_att(c, pre, x):
    return _att_out(c, pre + ".output", _attention(c, pre + ".self", x, x), x)


def _cross_att(c, pre, x, ctx_in):
    return _att_out(c, pre + ".output", _attention(c, pre + ".att", x, ctx_in), x)


def _ffn(c, inter, out, x):
    return _att_out(c, out, F.gelu(c.lin(inter + ".dense", x)), x)


def _layer(c, pre, x):
    a = _self_att(c, pre + ".attention", x)
    return _ffn(c, pre + ".intermediate", pre + ".output", a)


def _xlayer(c, pre, lang, visn):
    """LxmertXLayer.forward -- modeling_lxmert.py:935-1009: the SAME visual_attention module is applied
    in both directions (its four masked Linears run twice and their score gradients add)."""
    lang_x = _cross_att(c, pre + ".visual_attention", lang, visn)
    visn_x = _cross_att(c, pre + ".visual_attention", visn, lang)
    lang_s = _self_att(c, pre + ".lang_self_att", lang_x)
    visn_s = _self_att(c, pre + ".visn_self_att", visn_x)
    return (_ffn(c, pre + ".lang_inter", pre + ".lang_output", lang_s),
            _ffn(c, pre + ".visn_inter", pre + ".visn_output", visn_s))


def classifier(c, pooled):
    """SimpleClassifier with legacy weight_norm(dim=None): W = g * V / ||V||_F -- classifier.py:5-22."""
    def wn(pre):
        v, g = c.P[pre + ".weight_v"], c.P[pre + ".weight_g"]
        return v * (g / v.norm())
    h = F.relu(F.linear(pooled, wn("classifier.main.0"), c.P["classifier.main.0.bias"]))
    h = c.drop(h, c.p_cls)
    return F.linear(h, wn("classifier.main.3"), c.P["classifier.main.3.bias"])


def forward(c, ids, feats, pos, l_layers=9, r_layers=5, x_layers=5):
    """(logits, pooled) = LxmertForMultipleChoice.forward -- modeling_lxmert.py:256-360."""
    T = ids.shape[1]
    emb_name = "lxmert.embeddings.word_embeddings"
    if emb_name in c.S:
        words = _ops.masked_embedding(ids, c.S[emb_name], c.P[emb_name + ".weight"], c.T[emb_name], 0)
    else:
        words = F.embedding(ids, c.P[emb_name + ".weight"], padding_idx=0)
    # both tables are nn.Embedding(..., padding_idx=0) (modeling_lxmert.py:735-736): row 0 never receives a
    # gradient -- invisible in stage 2 (frozen) but part of the stage-3 fine-tune
    position = F.embedding(torch.arange(T, device=ids.device).unsqueeze(0).expand(ids.shape),
                           c.P["lxmert.embeddings.position_embeddings.weight"], padding_idx=0)
    token_type = F.embedding(torch.zeros_like(ids), c.P["lxmert.embeddings.token_type_embeddings.weight"], padding_idx=0)
    lang = c.drop(c.ln("lxmert.embeddings.LayerNorm", words + position + token_type), c.p_hidden)
    vf = "lxmert.encoder.visn_fc."
    visn = (c.ln(vf + "visn_layer_norm", c.lin(vf + "visn_fc", feats)) + c.ln(vf + "box_layer_norm", c.lin(vf + "box_fc", pos))) / 2
    visn = c.drop(visn, c.p_hidden)
    for l in range(l_layers):
        lang = _layer(c, f"lxmert.encoder.layer.{l}", lang)
    for l in range(r_layers):
        visn = _layer(c, f"lxmert.encoder.r_layers.{l}", visn)
    for l in range(x_layers):
        lang, visn = _xlayer(c, f"lxmert.encoder.x_layers.{l}", lang, visn)
    pooled = torch.tanh(c.lin("lxmert.pooler.dense", lang[:, 0]))
    return classifier(c, pooled), pooled


def compute_loss(kind, logits, pooled, batch, lmh=None, gamma=5.0):
    """Loss dispatch of Trainer._training_step -- mask_trainer_Robust_VQA.py:812-831."""
    if kind == "normal":
        return _losses.bce_loss(logits, batch["target"])
    if kind == "lpf":
        return _losses.lpf_loss(logits, batch["bias"], batch["max_label"], gamma)
    if kind == "lmh":
        return _losses.lmh_loss(pooled, logits, batch["bias"], batch["target"], lmh["lin_w"], lmh["lin_b"],
                                lmh["smooth_param"], w=0.36)
    raise ValueError(kind)


def synthetic_batch(B, A, seed=49, T=20, R=36, feat=2048, vocab=30522):
    """The synthetic stage-2 batch of SURVEY.md section 8(d) (order of draws matters)."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab, (B, T), generator=g)
    feats = torch.randn(B, R, feat, generator=g)
    pos = torch.rand(B, R, 4, generator=g)
    target = (torch.rand(B, A, generator=g) > 0.999).float() * torch.rand(B, A, generator=g)
    bias = torch.rand(B, A, generator=g) * 0.01
    return {"ids": ids, "feats": feats, "pos": pos, "target": target, "bias": bias, "max_label": target.argmax(1)}


def init_scores(params, rates, threshold=1e-2, l_layers=9, r_layers=5, x_layers=5):
    """Masker.patch_modules with controlled_init='magnitude' (maskers_Robust.py:577-612): per-module
    magnitude init at the modality's zero rate; every module starts with the shared threshold."""
    scores, thresholds, modal = {}, {}, {}
    for name, m in module_names(l_layers, r_layers, x_layers):
        s, _ = _ops.magnitude_init(params[name + ".weight"], rates[m], threshold)
        scores[name] = s.requires_grad_(True)
        thresholds[name] = float(torch.tensor(threshold))
        modal[name] = m
    return scores, thresholds, modal


def reset_thresholds(scores, modal, rates):
    """Trainer.reset_threshold -- mask_trainer_Robust_VQA.py:467-482 (returns thresholds and their mean)."""
    thr = {n: float(_ops.reset_threshold(s.detach(), rates[modal[n]])) for n, s in scores.items()}
    mean = float(torch.tensor([thr[n] for n in scores]).mean())
    return thr, mean


def training_step(c, batch, kind, opt_state=None, lmh=None, lr=5e-5, max_grad_norm=1.0, layers=(9, 5, 5)):
    """fwd -> loss -> backward -> clip_grad_norm_ -> AdamW.step -> zero_grad for the trainable set
    (scores + classifier), as Trainer.train / _training_step do (mask_trainer_Robust_VQA.py:640-680, 801-886)."""
    trainable = list(c.S.values()) + [c.P[k] for k in sorted(c.P) if k.startswith("classifier.")
                                      and c.P[k].requires_grad]
    logits, pooled = forward(c, batch["ids"], batch["feats"], batch["pos"], *layers)
    loss = compute_loss(kind, logits, pooled, batch, lmh=lmh)
    grads = torch.autograd.grad(loss, trainable, allow_unused=True)
    grads = [g if g is not None else torch.zeros_like(p) for g, p in zip(grads, trainable)]
    out = {"loss": loss.detach(), "logits": logits.detach(), "pooled": pooled.detach(), "grads": grads,
           "score": _losses.vqa_score(logits.detach(), batch["target"])}
    if opt_state is not None:
        coef, total = _adamw.clip_coef(grads, max_grad_norm)
        out["grad_norm"] = total
        with torch.no_grad():
            for p, g in zip(trainable, grads):
                st = opt_state.setdefault(id(p), _adamw.new_state(p))
                _adamw.adamw_step(p, g * coef, st, lr)
    return out
