"""``initialize_clip`` of the reference (mPLUG/models/visual_transformers.py:40-52) without the checkpoint file: the
reference loads ``ckpts/ViT-B-16.tar`` and resizes its position table to ``image_res``; no checkpoint ships, so the
visual tower is built at ``image_res`` directly with random init (``resize_pos_embed`` is kept for users who load one).
Returns ``(clip_model, None)`` where ``clip_model.visual`` is the tower, as the reference's call sites expect."""
import math

import torch
import torch.nn.functional as F
from torch import nn

from .clip.model import VisualTransformer

_CLIP = {"ViT-B-16": dict(patch_size=16, width=768, layers=12, heads=12, output_dim=512),
         "ViT-L-14": dict(patch_size=14, width=1024, layers=24, heads=16, output_dim=768)}


def resize_pos_embed(posemb, posemb_new):
    """Bilinear resize of the grid part of a [1, 1 + g*g, C] position table to the grid of ``posemb_new``."""
    tok, grid = posemb[:, :1], posemb[0, 1:]
    g_old, g_new = int(math.sqrt(len(grid))), int(math.sqrt(posemb_new.shape[1] - 1))
    grid = grid.reshape(1, g_old, g_old, -1).permute(0, 3, 1, 2)
    grid = F.interpolate(grid.float(), size=(g_new, g_new), mode="bilinear").to(grid.dtype)
    return torch.cat([tok, grid.permute(0, 2, 3, 1).reshape(1, g_new * g_new, -1)], dim=1)


class _ClipShell(nn.Module):
    """Stands for the CLIP model object: only ``.visual`` exists (the text tower is not on the VQA path)."""

    def __init__(self, visual):
        super().__init__()
        self.visual = visual


def initialize_clip(config, num_patches=240):
    spec = dict(_CLIP[config["clip_name"]])
    for key in ("width", "layers", "heads", "output_dim", "patch_size"):      # overridable for small test networks
        if "clip_" + key in config:
            spec[key] = config["clip_" + key]
    return _ClipShell(VisualTransformer(input_resolution=config["image_res"], **spec)), None
