"""mPLUG-VQA network on the GPU: the drop-in network masked by the drop-in masker (masked Linear layers on the sm_100a
GEMMs through libcrvqa.so, everything else torch) against the reference network masked by the reference masker
(tests/golden/mplug_model_tiny.pt).  Bit-exact: census, trainable set, thresholds, kept counts before and after a
threshold refresh.  Floating point: loss within 2e-2 relative, gradient norms within 15 % (bf16 MMA operands)."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_mplug_cpu import kept, masked, quiet, thr_record  # noqa: E402
from test_mplug_model_cpu import GOLD, build, run  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


def test_masked_network_matches_reference(gold):
    from mPLUG import vqa_mplug
    from mPLUG.engine import MaskTrainEngine
    from mPLUG.masking import maskers
    from mPLUG.masking.mask_config import MaskConfigs
    G = gold["masked"]
    model = build(gold, "cuda")
    dense, _ = run(model, gold, with_bias=True, device="cuda")
    assert dense == pytest.approx(gold["dense_loss_bias"], rel=2e-3)       # torch modules; attention cores in bf16
    conf = MaskConfigs()
    conf.zero_rate = 0.5
    masker = quiet(vqa_mplug.init_masker, conf, model, layers_to_mask=gold["layers_to_mask"])
    assert [n for n, _ in masked(model)] == G["module_names"]
    assert sorted(n for n, p in model.named_parameters() if p.requires_grad) == G["trainable"]
    assert thr_record(model) == G["thresholds"]
    assert kept(model) == G["kept"]
    loss, norms = run(model, gold, with_bias=True, device="cuda")
    assert loss == pytest.approx(G["loss"], rel=2e-2)
    assert sorted(norms) == sorted(G["grad_norms"])
    for n, want in G["grad_norms"].items():
        assert norms[n] == pytest.approx(want, rel=0.15, abs=1e-7), n
    mean = maskers.reset_threshold(model, 0.7)
    r = G["reset_0.7"]
    assert mean == r["mean"] and thr_record(model) == r["thresholds"] and kept(model) == r["kept"]
    loss, _ = run(model, gold, with_bias=True, device="cuda")
    assert loss == pytest.approx(G["loss_after_reset"], rel=2e-2)

    # a few engine steps on the real network: the loss falls, frozen weights and the momentum twins' masks stay put
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=1e-3, weight_decay=0.0)
    eng = MaskTrainEngine(model, opt, gradient_clipping=1.0, bf16=True)
    from test_mplug_model_cpu import batch
    image, question, answer, k, weights, bias = batch(gold, "cuda")
    first = last = None
    for step in range(8):
        loss = eng(image, question, answer, train=True, alpha=0.4, k=k, weights=weights, bias=bias)
        eng.backward(loss)
        eng.step()
        first = float(loss) if first is None else first
        last = float(loss)
        if eng.global_steps % 4 == 0:
            quiet(vqa_mplug.update_masks, eng, masker, 0)
    assert last < first and eng.global_steps == 8
