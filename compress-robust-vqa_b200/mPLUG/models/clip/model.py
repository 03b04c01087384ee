"""ViT image encoder of CLIP as mPLUG uses it (reference mPLUG/models/clip/model.py:157-249): same module tree and
parameter names (``conv1``, ``class_embedding``, ``positional_embedding``, ``ln_pre``, ``transformer.resblocks.{l}.
{attn,ln_1,mlp.c_fc,mlp.c_proj,ln_2}``, ``ln_post``, ``proj``), so reference checkpoints load with ``strict=True`` and
the masker finds ``mlp.c_fc`` / ``mlp.c_proj`` by name.  Plain torch modules: the masked layers become sm_100a masked
GEMMs when ``Masker.patch_modules`` swaps them."""
import contextlib
import os
from collections import OrderedDict

import torch
import torch.nn.functional as F
from torch import nn

# see modeling_mplug.BF16_ATTENTION: the attention sub-block (in_proj, softmax(QK^T)V, out_proj) runs under bf16 autocast
# on a GPU, as the whole network does in the reference's DeepSpeed-bf16 run
BF16_ATTENTION = os.environ.get("CRVQA_MPLUG_BF16_ATTENTION", "1") != "0"


# CRVQA_MPLUG_VIT_ATTENTION=flash_attn: the tower's attention core through flash_attn's packed-QKV kernel (a library, as
# torch's own scaled_dot_product_attention backends are) instead of torch's SDPA dispatch; anything else: SDPA
_PACKED_ATTENTION = None
if os.environ.get("CRVQA_MPLUG_VIT_ATTENTION", "sdpa") == "flash_attn":
    try:
        from flash_attn import flash_attn_qkvpacked_func as _PACKED_ATTENTION
    except ImportError:
        _PACKED_ATTENTION = None


class LayerNorm(nn.LayerNorm):
    """LayerNorm that returns its input's dtype (half-precision inputs are normalised by torch in fp32 anyway)."""

    def forward(self, x):
        if x.is_cuda and x.dtype == torch.bfloat16 and os.environ.get("CRVQA_MPLUG_FUSED", "1") != "0":
            from crvqa import fused
            if fused.layernorm_bf16_usable(x, self):
                return fused.layernorm_bf16(x, self)     # one bf16 -> bf16 pass (fp32 statistics) instead of three
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

    def attention(self, x, text_mask=None, batch_first=False):
        if text_mask is None and self.attn_mask is not None:
            text_mask = self.attn_mask.to(dtype=x.dtype, device=x.device)
        bf16 = (torch.autocast("cuda", dtype=torch.bfloat16) if BF16_ATTENTION and x.is_cuda
                else contextlib.nullcontext())
        with bf16:
            if batch_first:
                assert text_mask is None
                return self._self_attention_batch_first(x).to(x.dtype)
            return self.attn(x, x, x, need_weights=False, attn_mask=text_mask)[0].to(x.dtype)

    def _self_attention_batch_first(self, x):
        """What ``self.attn(x, x, x)`` computes (torch.nn.functional.multi_head_attention_forward: packed in-projection,
        softmax(Q K^T / sqrt(d)) V with dropout on the probabilities, out-projection read through ``out_proj.weight`` /
        ``.bias`` -- NOT ``out_proj.forward``, so a masked replacement of ``out_proj`` stays unmasked exactly as under
        the reference's nn.MultiheadAttention) on a [batch, tokens, width] input.  The heads are strided views of the
        packed projection and of the attention output: none of the four [tokens x batch x width]-sized layout copies
        per block and direction that the sequence-first module makes (its ``.contiguous()`` after the packed
        projection, the head merge after attention, and their backward copies / zero-filled select gradients)."""
        a = self.attn
        B, L, D = x.shape
        H = a.num_heads
        qkv = F.linear(x, a.in_proj_weight, a.in_proj_bias).view(B, L, 3, H, D // H)
        p_drop = a.dropout if a.training else 0.0
        if _PACKED_ATTENTION is not None and qkv.is_cuda and qkv.dtype in (torch.bfloat16, torch.float16):
            # library kernel on the packed projection itself: its backward writes d(qkv) in place of three tensors + a stack
            ctx = _PACKED_ATTENTION(qkv, dropout_p=p_drop).reshape(B, L, D)
        else:
            q, k, v = (t.transpose(1, 2) for t in qkv.unbind(2))     # [B, H, L, d] views; unbind's backward is one stack
            ctx = F.scaled_dot_product_attention(q, k, v, dropout_p=p_drop).transpose(1, 2).reshape(B, L, D)
        return F.linear(ctx, a.out_proj.weight, a.out_proj.bias)

    def batch_first_ok(self, text_mask=None):
        a = self.attn
        return (text_mask is None and self.attn_mask is None and a._qkv_same_embed_dim and a.bias_k is None
                and a.bias_v is None and not a.add_zero_attn and a.in_proj_bias is not None)

    def forward(self, x, text_mask=None, batch_first=False):
        x = x + self.attention(self.ln_1(x), text_mask=text_mask, batch_first=batch_first)
        return x + self.mlp(self.ln_2(x))


class Transformer(nn.Module):
    def __init__(self, width, layers, heads, attn_mask=None):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads, attn_mask) for _ in range(layers)])

    def forward(self, x, text_mask=None, use_checkpoint=False, batch_first=False):
        """x: [tokens, batch, width] as in the reference, or [batch, tokens, width] with ``batch_first``."""
        for block in self.resblocks:
            if use_checkpoint:
                x = torch.utils.checkpoint.checkpoint(block, x, text_mask, batch_first, use_reentrant=False)
            else:
                x = block(x, text_mask=text_mask, batch_first=batch_first)
        return x

    def batch_first_ok(self, text_mask=None):
        return (os.environ.get("CRVQA_MPLUG_FUSED", "1") != "0"
                and all(block.batch_first_ok(text_mask) for block in self.resblocks))


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
        if self.transformer.batch_first_ok(text_mask):
            # same arithmetic, batch-major: the attention blocks split heads as views (no layout copies)
            x = self.transformer(x, use_checkpoint=use_checkpoint, batch_first=True)
        else:
            x = self.transformer(x.permute(1, 0, 2), use_checkpoint=use_checkpoint).permute(1, 0, 2)
        return self.ln_post(x) if skip_last_layer else x @ self.proj
