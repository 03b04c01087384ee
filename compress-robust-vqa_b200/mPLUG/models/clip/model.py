"""ViT image encoder of CLIP as mPLUG uses it (reference mPLUG/models/clip/model.py:157-249): same module tree and
parameter names (``conv1``, ``class_embedding``, ``positional_embedding``, ``ln_pre``, ``transformer.resblocks.{l}.
{attn,ln_1,mlp.c_fc,mlp.c_proj,ln_2}``, ``ln_post``, ``proj``), so reference checkpoints load with ``strict=True`` and
the masker finds ``mlp.c_fc`` / ``mlp.c_proj`` by name.  Plain torch modules: the masked layers become sm_100a masked
GEMMs when ``Masker.patch_modules`` swaps them."""
import contextlib
import os
from collections import OrderedDict

import torch
from torch import nn

# see modeling_mplug.BF16_ATTENTION: the attention sub-block (in_proj, softmax(QK^T)V, out_proj) runs under bf16 autocast
# on a GPU, as the whole network does in the reference's DeepSpeed-bf16 run
BF16_ATTENTION = os.environ.get("CRVQA_MPLUG_BF16_ATTENTION", "1") != "0"


class LayerNorm(nn.LayerNorm):
    """LayerNorm that returns its input's dtype (half-precision inputs are normalised by torch in fp32 anyway)."""

    def forward(self, x):
        return super().forward(x).type(x.dtype)


class QuickGELU(nn.Module):
    def forward(self, x):
        if (x.is_cuda and x.dtype == torch.bfloat16 and x.requires_grad and x.numel() % 8 == 0
                and os.environ.get("CRVQA_MPLUG_FUSED", "1") != "0"):
            from crvqa import fused
            return fused.quick_gelu_bf16(x)      # one pass each way instead of 3 + 5 elementwise kernels
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model, n_head, attn_mask=None):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head, dropout=0.1)
        self.ln_1 = LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict(c_fc=nn.Linear(d_model, d_model * 4), gelu=QuickGELU(),
                                             c_proj=nn.Linear(d_model * 4, d_model)))
        self.ln_2 = LayerNorm(d_model)
        self.attn_mask = attn_mask

    def attention(self, x, text_mask=None):
        if text_mask is None and self.attn_mask is not None:
            text_mask = self.attn_mask.to(dtype=x.dtype, device=x.device)
        bf16 = (torch.autocast("cuda", dtype=torch.bfloat16) if BF16_ATTENTION and x.is_cuda
                else contextlib.nullcontext())
        with bf16:
            return self.attn(x, x, x, need_weights=False, attn_mask=text_mask)[0].to(x.dtype)

    def forward(self, x, text_mask=None):
        x = x + self.attention(self.ln_1(x), text_mask=text_mask)
        return x + self.mlp(self.ln_2(x))


class Transformer(nn.Module):
    def __init__(self, width, layers, heads, attn_mask=None):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])

    def forward(self, x, text_mask=None, use_checkpoint=False):
        for block in self.resblocks:
            if use_checkpoint:
                x = torch.utils.checkpoint.checkpoint(block, x, text_mask, use_reentrant=False)
            else:
                x = block(x, text_mask=text_mask)
        return x


class VisualTransformer(nn.Module):
    def __init__(self, input_resolution, patch_size, width, layers, heads, output_dim):
        super().__init__()
        self.input_resolution, self.output_dim, self.heads = input_resolution, output_dim, heads
        self.conv1 = nn.Conv2d(3, width, kernel_size=patch_size, stride=patch_size, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((input_resolution // patch_size) ** 2 + 1, width))
        self.ln_pre = LayerNorm(width)
        self.transformer = Transformer(width, layers, heads)
        self.ln_post = LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, output_dim))

    def forward(self, x, skip_last_layer=False, text_embedding=None, text_mask=None, use_checkpoint=False):
        x = self.conv1(x).flatten(2).transpose(1, 2)                           # [B, grid^2, width]
        cls = self.class_embedding.to(x.dtype).expand(x.shape[0], 1, -1)
        x = torch.cat([cls, x], dim=1)
        x = self.ln_pre(x + self.positional_embedding.to(x.dtype)[:x.size(1)])
        x = self.transformer(x.permute(1, 0, 2), use_checkpoint=use_checkpoint).permute(1, 0, 2)
        return self.ln_post(x) if skip_last_layer else x @ self.proj
