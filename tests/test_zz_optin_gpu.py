"""Opt-in modes, run last: mPLUG with bf16 activations (CRVQA_MPLUG_BF16_ACTIVATIONS=1 -- the whole forward under bf16
autocast, which is the reference's DeepSpeed-bf16 arithmetic).  Masks, thresholds and kept counts do not depend on the
activation dtype and stay bit-exact; loss and gradient norms are compared with the fp32 reference run at bf16-level
tolerances (written below)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_mplug_cpu import kept, masked, quiet, thr_record  # noqa: E402
from test_mplug_model_cpu import GOLD, build, run  # noqa: E402

pytestmark = pytest.mark.gpu


def test_mplug_bf16_activation_mode():
    from mPLUG import vqa_mplug
    from mPLUG.masking.mask_config import MaskConfigs
    gold = torch.load(GOLD, weights_only=False)
    G = gold["masked"]
    model = build(gold, "cuda")
    model.bf16_activations = True
    conf = MaskConfigs()
    conf.zero_rate = 0.5
    quiet(vqa_mplug.init_masker, conf, model, layers_to_mask=gold["layers_to_mask"])
    assert thr_record(model) == G["thresholds"] and kept(model) == G["kept"]
    loss, norms = run(model, gold, with_bias=True, device="cuda")
    assert loss == pytest.approx(G["loss"], rel=3e-2)
    assert sorted(norms) == sorted(G["grad_norms"])
    big = max(G["grad_norms"].values())
    for n, want in G["grad_norms"].items():
        if want > 1e-3 * big:                      # tiny gradients are dominated by bf16 rounding of the activations
            assert norms[n] == pytest.approx(want, rel=0.3), n
    # the masked layers really ran with bf16 activations: their outputs are bf16 under this mode
    mod = dict(masked(model))["text_encoder.encoder.layer.0.intermediate.dense"]
    with torch.autocast("cuda", dtype=torch.bfloat16):
        y = mod(torch.randn(2, 7, mod.weight.shape[1], device="cuda"))
    assert y.dtype == torch.bfloat16
    assert mod(torch.randn(2, 7, mod.weight.shape[1], device="cuda")).dtype == torch.float32
