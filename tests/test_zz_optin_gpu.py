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


@pytest.mark.parametrize("through_engine", [False, True])
def test_mplug_training_trajectory_follows_reference(through_engine):
    """through_engine: the same steps driven by MaskTrainEngine (fp32 score comparison as in the golden run), i.e. the
    multi-tensor clip + AdamW launches with the masked operands refreshed by the optimiser pass.
    The drop-in masker + torch AdamW on the GPU against the reference's six-step trajectory (golden 'T'): losses
    within 2e-2, thresholds within one bf16 step, kept counts within 2 % (bf16 MMA operands perturb the scores that
    the order statistics are taken from, so exact equality is not expected after optimiser steps; measured on B200:
    worst module 34 of 2846 kept entries = 1.2 %)."""
    import mplug_skeleton as sk
    from mPLUG.masking import maskers
    from test_mplug_cpu import GOLD as SK_GOLD
    from test_mplug_cpu import _init, fresh
    gold = torch.load(SK_GOLD, weights_only=False)
    T = gold["T"]
    model = fresh(gold).cuda()
    masker = _init(model, zero_rate=0.7, init_sparsity=0.3, final_sparsity_epoch=2)
    assert thr_record(model) == T["init_thresholds"] and kept(model) == T["kept_init"]
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=T["lr"], weight_decay=0.0)
    data = [t.cuda() for t in sk.batch()]
    model.train()
    eng = None
    if through_engine:
        from mPLUG.engine import MaskTrainEngine
        eng = MaskTrainEngine(model, opt, gradient_clipping=1.0, bf16=False)
    for step, want in enumerate(T["steps"]):
        if eng is not None:
            loss = eng(*data)
            eng.backward(loss)
            eng.step()
            assert eng._fused and eng._fused.plan is not None and eng._fused.plan.uniform_steps
        else:
            loss = model(*data)
            opt.zero_grad()
            loss.backward()
            torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.requires_grad and p.grad is not None],
                                           1.0)
            opt.step()
        assert float(loss.detach()) == pytest.approx(want["loss"], rel=2e-2), step
        if "target" in want:
            _, target, _ = masker.masker_scheduler.step(cur_epoch=(step + 1) // 2)
            maskers.reset_threshold(model, target)
            got_thr, got_kept = thr_record(model), kept(model)
            for n, (value, dtype) in want["thresholds"].items():
                assert got_thr[n][1] == dtype and got_thr[n][0] == pytest.approx(value, rel=1e-2, abs=1e-4), (step, n)
                assert abs(got_kept[n] - want["kept"][n]) <= max(2, 0.02 * want["kept"][n]), (step, n)
