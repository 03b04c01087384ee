"""A miniature network with mPLUG's module tree (the names of mPLUG/masking/maskers.py:16-62): CLIP-style visual
resblocks, a BERT text encoder, a fusion encoder and a text decoder with cross-attention, an LM head under
``cls.predictions`` and one momentum twin (``text_encoder_m``).  Test scaffolding only: it lets the reference's mPLUG
masker (golden generator, CPU) and this repo's (GPU tests) patch the same modules and train the same loss.  It is
NOT the mPLUG model (no patch convolution, no stride layers, no beam search)."""
import math
import types
from collections import OrderedDict

import torch
import torch.nn as nn
import torch.nn.functional as F

CFG = dict(hidden=64, heads=4, ffn=128, vocab=96, vis_tokens=16, patch_dim=48,
           vis_layers=2, text_layers=2, fusion_layers=4, fusion_first=2, dec_layers=1)


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class ResBlock(nn.Module):
    def __init__(self, d, heads):
        super().__init__()
        self.attn = nn.MultiheadAttention(d, heads)
        self.ln_1 = nn.LayerNorm(d)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d, d * 2)), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d * 2, d))]))
        self.ln_2 = nn.LayerNorm(d)

    def forward(self, x):                       # x: [tokens, batch, d]
        h = self.ln_1(x)
        x = x + self.attn(h, h, h, need_weights=False)[0]
        return x + self.mlp(self.ln_2(x))


class Visual(nn.Module):
    def __init__(self, c):
        super().__init__()
        d = c["hidden"]
        self.patch = nn.Linear(c["patch_dim"], d, bias=False)
        self.class_embedding = nn.Parameter(torch.randn(d) * d ** -0.5)
        self.positional_embedding = nn.Parameter(torch.randn(c["vis_tokens"] + 1, d) * d ** -0.5)
        self.transformer = nn.Module()
        self.transformer.resblocks = nn.Sequential(*[ResBlock(d, c["heads"]) for _ in range(c["vis_layers"])])

    def forward(self, patches):                 # [B, tokens, patch_dim]
        x = self.patch(patches)
        cls = self.class_embedding.expand(x.shape[0], 1, -1)
        x = torch.cat([cls, x], 1) + self.positional_embedding
        return self.transformer.resblocks(x.permute(1, 0, 2)).permute(1, 0, 2)


class SelfAtt(nn.Module):
    def __init__(self, d, heads):
        super().__init__()
        self.heads = heads
        self.query, self.key, self.value = nn.Linear(d, d), nn.Linear(d, d), nn.Linear(d, d)

    def forward(self, x, ctx=None):
        ctx = x if ctx is None else ctx
        B, T, D = x.shape
        h = self.heads

        def split(t):
            return t.view(B, -1, h, D // h).transpose(1, 2)

        q, k, v = split(self.query(x)), split(self.key(ctx)), split(self.value(ctx))
        p = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(D // h), -1)
        return (p @ v).transpose(1, 2).reshape(B, T, D)


class AddNorm(nn.Module):
    def __init__(self, d_in, d):
        super().__init__()
        self.dense = nn.Linear(d_in, d)
        self.LayerNorm = nn.LayerNorm(d, eps=1e-12)

    def forward(self, h, x):
        return self.LayerNorm(self.dense(h) + x)


class Attention(nn.Module):
    def __init__(self, d, heads):
        super().__init__()
        self.self = SelfAtt(d, heads)
        self.output = AddNorm(d, d)

    def forward(self, x, ctx=None):
        return self.output(self.self(x, ctx), x)


class Intermediate(nn.Module):
    def __init__(self, d, ffn):
        super().__init__()
        self.dense = nn.Linear(d, ffn)

    def forward(self, x):
        return F.gelu(self.dense(x))


class Layer(nn.Module):
    def __init__(self, c, cross):
        super().__init__()
        d = c["hidden"]
        self.attention = Attention(d, c["heads"])
        if cross:
            self.crossattention = Attention(d, c["heads"])
        self.intermediate = Intermediate(d, c["ffn"])
        self.output = AddNorm(c["ffn"], d)

    def forward(self, x, ctx=None):
        x = self.attention(x)
        if ctx is not None:
            x = self.crossattention(x, ctx)
        return self.output(self.intermediate(x), x)


class Encoder(nn.Module):
    def __init__(self, c, n, cross):
        super().__init__()
        self.layer = nn.ModuleList([Layer(c, cross) for _ in range(n)])


class Embeddings(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.word_embeddings = nn.Embedding(c["vocab"], c["hidden"], padding_idx=0)
        self.LayerNorm = nn.LayerNorm(c["hidden"], eps=1e-12)

    def forward(self, ids):
        return self.LayerNorm(self.word_embeddings(ids))


class TextEncoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.config = types.SimpleNamespace(num_attention_heads=c["heads"], hidden_size=c["hidden"])
        self.embeddings = Embeddings(c)
        self.encoder = Encoder(c, c["text_layers"], cross=False)

    def forward(self, ids):
        x = self.embeddings(ids)
        for lyr in self.encoder.layer:
            x = lyr(x)
        return x


class FusionEncoder(nn.Module):
    """Layers 0..fusion_first-1 exist but never run -- like fusion_encoder.encoder.layer.0-5 of mPLUG-VQA, which the
    reference's see_sparsity excludes."""

    def __init__(self, c):
        super().__init__()
        self.first = c["fusion_first"]
        self.encoder = Encoder(c, c["fusion_layers"], cross=True)

    def forward(self, text, image):
        for lyr in self.encoder.layer[self.first:]:
            text = lyr(text, image)
        return text


class Predictions(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.transform = nn.Module()
        self.transform.dense = nn.Linear(c["hidden"], c["hidden"])
        self.transform.LayerNorm = nn.LayerNorm(c["hidden"], eps=1e-12)
        self.decoder = nn.Linear(c["hidden"], c["vocab"])

    def forward(self, x):
        return self.decoder(self.transform.LayerNorm(F.gelu(self.transform.dense(x))))


class TextDecoder(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.bert = nn.Module()
        self.bert.embeddings = Embeddings(c)
        self.bert.encoder = Encoder(c, c["dec_layers"], cross=True)
        self.cls = nn.Module()
        self.cls.predictions = Predictions(c)

    def forward(self, ids, states):
        x = self.bert.embeddings(ids)
        for lyr in self.bert.encoder.layer:
            x = lyr(x, states)
        return self.cls.predictions(x)


class SkeletonMPLUG(nn.Module):
    def __init__(self, c=None):
        super().__init__()
        c = dict(CFG, **(c or {}))
        self.c = c
        self.visual_encoder = nn.Module()
        self.visual_encoder.visual = Visual(c)
        self.text_encoder = TextEncoder(c)
        self.fusion_encoder = FusionEncoder(c)
        self.text_decoder = TextDecoder(c)
        self.text_encoder_m = TextEncoder(c)      # momentum twin: patched by name, never trained
        for p in self.text_encoder_m.parameters():
            p.requires_grad = False

    def forward(self, patches, question_ids, answer_ids, weights):
        image = self.visual_encoder.visual(patches)
        text = self.text_encoder(question_ids)
        fused = self.fusion_encoder(text, image)
        states = torch.cat([image, fused], 1)
        logits = self.text_decoder(answer_ids[:, :-1], states)
        nll = F.cross_entropy(logits.reshape(-1, logits.shape[-1]).float(), answer_ids[:, 1:].reshape(-1),
                              ignore_index=0, reduction="none").view(answer_ids.shape[0], -1).sum(1)
        return (weights * nll).sum() / patches.shape[0]


WEIGHT_TYPES = {
    "visual_encoder": ["I_visual", "O_visual", "AO_visual"],
    "text_encoder": ["K", "Q", "V", "AO", "I", "O", "E"],
    "fusion_encoder": ["SK", "SQ", "SV", "SAO", "CK", "CQ", "CV", "CAO", "I", "O"],
    "text_decoder": ["SK", "SQ", "SV", "SAO", "CK", "CQ", "CV", "CAO", "I", "O"],
}
LAYERS = {"visual_encoder": [0, 1], "text_encoder": [0, 1], "fusion_encoder": [2, 3], "text_decoder": [0]}


def names_to_mask(chain_module_names):
    names = set()
    for tower, abbres in WEIGHT_TYPES.items():
        names.update(chain_module_names(tower, LAYERS[tower], abbres))
    return names


def build(seed=7):
    torch.manual_seed(seed)
    return SkeletonMPLUG()


def batch(seed=11, B=8, c=CFG):
    g = torch.Generator().manual_seed(seed)
    patches = torch.randn(B, c["vis_tokens"], c["patch_dim"], generator=g)
    q = torch.randint(1, c["vocab"], (B, 8), generator=g)
    a = torch.randint(1, c["vocab"], (B, 5), generator=g)
    a[:, -1] = 0                                   # padded tail, ignored by the loss
    w = torch.rand(B, generator=g) + 0.5
    return patches, q, a, w
