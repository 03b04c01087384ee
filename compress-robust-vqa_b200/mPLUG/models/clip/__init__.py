"""CLIP visual tower used by mPLUG (reference: mPLUG/models/clip/model.py; only ``VisualTransformer`` is on the
VQA path -- the CLIP text tower is never run, the reference's ``see_sparsity`` excludes it by prefix)."""
from .model import LayerNorm, QuickGELU, ResidualAttentionBlock, Transformer, VisualTransformer  # noqa: F401
