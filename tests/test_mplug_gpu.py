"""mPLUG masking path on the GPU: the drop-in ``mPLUG/masking/maskers.py``,
``vqa_mplug.py`` and ``engine.py`` through libcrvqa.so, against the reference's outputs on the miniature mPLUG-shaped
network (tests/golden/mplug_skeleton.pt).  Bit-exact: module census, thresholds (value and dtype), kept counts, masks
in fp32 and in bf16 score mode, threshold means.  Floating point: loss within 2e-2 relative and score gradients within
0.1 norm-wise of the fp32 reference (bf16 MMA operands through ~8 masked GEMMs in sequence)."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mplug_skeleton as sk  # noqa: E402
from oracle import mplug_masking as om  # noqa: E402
from test_mplug_cpu import GOLD, _init, digest, fresh, kept, masked, perturb, quiet, thr_record  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


def _train_state(model):
    for p in model.parameters():
        p.grad = None
    model.train()
    loss = model(*[t.cuda() for t in sk.batch()])
    loss.backward()
    torch.cuda.synchronize()
    return float(loss.detach()), {n: m.weight_mask.grad for n, m in masked(model)}


def _check_grads(grads, want, tol):
    for n, g in grads.items():
        if want[n] is None:
            assert g is None, n
            continue
        ref = want[n].cuda()
        err = float((g - ref).norm() / ref.norm().clamp_min(1e-30))
        assert err < tol, (n, err)


def test_mplug_masker_matches_reference_fp32_and_bf16_modes(gold):
    from mPLUG.masking import maskers
    A = gold["A"]
    model = fresh(gold).cuda()
    masker = _init(model, zero_rate=0.7)
    assert [n for n, _ in masked(model)] == A["module_names"]
    assert sorted(n for n, p in model.named_parameters() if p.requires_grad) == A["trainable"]
    assert thr_record(model) == A["init_thresholds"]
    assert kept(model) == A["kept_init"]
    for n, m in masked(model):
        assert isinstance(m, maskers.MaskedLinear1) and m.weight_mask.is_cuda
        assert np.array_equal(om.packed(masker.init_masks[f"{n}_weight_mask"]), A["init_masks"][f"{n}_weight_mask"])
    assert round(quiet(maskers.see_sparsity, model), 2) == A["start_see_sparsity"]
    assert round(quiet(maskers.save_model_mask, model, is_save=False), 2) == A["start_zero_rate"]

    loss, grads = _train_state(model)
    assert loss == pytest.approx(A["loss"], rel=2e-2)
    _check_grads(grads, A["grads"], 0.1)
    head = float(model.text_decoder.cls.predictions.decoder.weight.grad.norm())
    assert head == pytest.approx(A["head_grad_norm"], rel=5e-2)

    perturb(model, *A["perturb"])
    for r in A["resets"]:
        mean = maskers.reset_threshold(model, r["rate"])
        assert thr_record(model) == r["thresholds"]
        assert kept(model) == r["kept"]
        assert mean == r["mean"]
    for n, m in masked(model):
        assert np.array_equal(om.packed(m.get_masks()[0]), A["after_masks"][n + ".weight"]), n
    before = thr_record(model)
    maskers.reset_threshold(model, 1e-4)
    assert before == thr_record(model)
    loss, grads = _train_state(model)
    assert loss == pytest.approx(A["after_train"]["loss"], rel=2e-2)
    for n, g in grads.items():
        want = A["after_train"]["grad_norms"][n]
        assert (g is None) if want is None else float(g.norm()) == pytest.approx(want, rel=0.1), n

    # bf16 score mode (the reference's DeepSpeed-bf16 model copy): fp32 master scores here, same masks / thresholds
    D = gold["D"]
    assert digest({n: m.weight_mask for n, m in masked(model)}) == D["fp32_scores_sha256"]
    maskers.set_score_dtype(model, torch.bfloat16)
    assert kept(model) == D["kept_before"]
    for r in D["resets"]:
        mean = maskers.reset_threshold(model, r["rate"])
        assert thr_record(model) == r["thresholds"]
        assert mean == r["mean"]
        for n, m in masked(model):
            assert np.array_equal(om.packed(m.get_masks()[0]), r["masks"][n]), n
    # the masked GEMM sees the same mask as get_masks(): forward with bf16-mode thresholds equals W (.) M applied by hand
    name, mod = next((n, m) for n, m in masked(model) if n.endswith("intermediate.dense"))
    x = torch.randn(24, mod.weight.shape[1], device="cuda")
    want = torch.nn.functional.linear(x.bfloat16().float(), mod.weight.bfloat16().float() * mod.get_masks()[0],
                                      mod.bias)
    got = mod(x)
    assert float((got - want).abs().max() / want.abs().max()) < 2e-3, name


def test_mplug_ramp_constant_scores_and_global_variants(gold, tmp_path):
    from mPLUG import vqa_mplug
    from mPLUG.masking import maskers
    B0 = gold["B0"]
    model = fresh(gold).cuda()
    _init(model, zero_rate=0.7, init_sparsity=0.0, final_sparsity_epoch=4)
    assert thr_record(model) == B0["init_thresholds"] and kept(model) == B0["kept_init"]
    loss, _ = _train_state(model)                  # thresholds that are the Python int 0 reach the kernels too
    assert loss == pytest.approx(B0["loss"], rel=2e-2)

    B = gold["B"]
    model = fresh(gold).cuda()
    masker = _init(model, zero_rate=0.7, init_sparsity=0.1, final_sparsity_epoch=4)
    out_dir = str(tmp_path / "masks")
    for r in B["ramp"]:
        mean, target = quiet(vqa_mplug.update_masks, model, masker, r["epoch"], out_dir)
        assert target == r["target"] and mean == r["mean"]
        assert thr_record(model) == r["thresholds"] and kept(model) == r["kept"]
    saved = torch.load(os.path.join(out_dir, "mask.pt"))
    for n, m in masked(model):
        assert not saved[n + ".weight"].is_cuda and torch.equal(saved[n + ".weight"], m.get_masks()[0].cpu())
    cs = B["constant_scores"]
    mod = dict(masked(model))[cs["module"]]
    mod.weight_mask.data.fill_(0.25)
    maskers.reset_threshold(model, 0.5)
    assert float(mod.threshold) == cs["threshold_after"]

    C = gold["C"]
    model = fresh(gold).cuda()
    masker = _init(model, zero_rate=0.6, init_sparsity=0.5, controlled_init="magnitude", global_prune=True)
    assert float(masker.global_threshold) == C["global_weight_threshold"]
    assert kept(model) == C["kept_init"] and thr_record(model) == C["init_thresholds"]
    loss, grads = _train_state(model)
    assert loss == pytest.approx(C["loss"], rel=2e-2)
    for n, g in grads.items():
        want = C["grad_norms"][n]
        assert (g is None) if want is None else float(g.norm()) == pytest.approx(want, rel=0.1), n
    perturb(model, *C["perturb"])
    for r in C["global_resets"]:
        mean = maskers.reset_threshold(model, r["rate"], global_prune=True)
        assert thr_record(model) == r["thresholds"] and kept(model) == r["kept"] and mean == r["mean"]


def test_mplug_engine_trains_scores_and_head_only(gold):
    """The reference loop (forward -> engine.backward -> engine.step, mask update every few steps) on the GPU: the
    loss falls, only scores and the LM head move, and after every update the masks the kernels apply are the
    bf16-mode masks of the oracle for the current scores and thresholds."""
    from mPLUG import vqa_mplug
    from mPLUG.engine import MaskTrainEngine
    model = fresh(gold).cuda()
    masker = _init(model, zero_rate=0.6, init_sparsity=0.2, final_sparsity_epoch=2)
    frozen = {n: p.detach().clone() for n, p in model.named_parameters() if not p.requires_grad}
    head0 = model.text_decoder.cls.predictions.decoder.weight.detach().clone()
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=2e-3, weight_decay=0.0)
    eng = MaskTrainEngine(model, opt, gradient_clipping=1.0, bf16=True)
    batch = tuple(t.cuda() for t in sk.batch())
    losses = []
    for epoch in range(3):
        mean = quiet(vqa_mplug.train_pretokenized, eng, [batch] * 4, epoch, masker=masker, masker_update_step=2)
        losses.append(mean)
        for n, m in masked(model):
            want = om.mask_of(m.weight_mask.detach().cpu(), m.threshold.cpu() if torch.is_tensor(m.threshold)
                              else m.threshold, torch.bfloat16)
            assert torch.equal(m.get_masks()[0].cpu(), want), n
    assert eng.global_steps == 12
    assert losses[-1] < losses[0]
    # the operand every masked GEMM holds between steps is exactly bf16(W) (.) current mask
    loss = eng(*batch)
    held = [m for _, m in masked(model) if getattr(m, "_wm", None) is not None]
    assert len(held) >= 40
    for m in held:
        assert torch.equal(m._wm.float(), m._weight_bf16().float() * m.get_masks()[0])
    # and with the cache off (mask re-derived inside every GEMM call) the same step gives the same loss
    for m in held:
        m.hold_masked_weight(False)
    assert float(eng(*batch)) == pytest.approx(float(loss), rel=1e-3)
    for n, p in model.named_parameters():
        if n in frozen:
            assert torch.equal(p, frozen[n]), n
    assert not torch.equal(model.text_decoder.cls.predictions.decoder.weight, head0)
    assert float(eng.last_grad_norm) > 0


def test_fused_engine_step_matches_clip_plus_torch_adamw(gold):
    """MaskTrainEngine.step() on a GPU runs clip + AdamW + the refresh of every held masked operand as multi-tensor
    launches (crv_sumsq_multi, crv_adamw_multi) over the optimiser's own tensors.  Fed the SAME gradients, it must
    follow torch.nn.utils.clip_grad_norm_ + torch.optim.AdamW.step (what the PyTorch path of the engine runs) to fp32
    rounding over several steps, with two parameter groups (decay / no decay) and a moving learning rate; the operand
    it leaves behind is bit-exactly bf16(W) (.) bf16-mode mask of the NEW scores; optimizer.state stays torch's."""
    from crvqa import lib
    from mPLUG import optim as mplug_optim
    from mPLUG.engine import MaskTrainEngine
    import types
    model = fresh(gold).cuda()
    _init(model, zero_rate=0.6, init_sparsity=0.2, final_sparsity_epoch=2)
    args = types.SimpleNamespace(opt="adamW", lr=2e-3, weight_decay=0.02)
    opt = mplug_optim.create_optimizer(args, model)
    assert len(opt.param_groups) == 2 and opt.param_groups[1]["weight_decay"] == 0.02
    eng = MaskTrainEngine(model, opt, gradient_clipping=0.05, bf16=True)
    trainable = [p for g in opt.param_groups for p in g["params"]]
    twin = [[p.detach().clone().requires_grad_(True) for p in g["params"]] for g in opt.param_groups]
    ref = torch.optim.AdamW([{"params": twin[0], "weight_decay": 0.0}, {"params": twin[1], "weight_decay": 0.02}],
                            lr=2e-3)
    flat_twin = [q for grp in twin for q in grp]
    batch = tuple(t.cuda() for t in sk.batch())
    model.train()
    for it in range(5):
        for g, h in zip(opt.param_groups, ref.param_groups):
            g["lr"] = h["lr"] = 2e-3 * (1.0 - 0.1 * it)
        loss = eng(*batch)
        eng.backward(loss)
        with_grad = 0
        for p, q in zip(trainable, flat_twin):
            q.grad = None if p.grad is None else p.grad.detach().clone()
            with_grad += p.grad is not None
        assert with_grad >= 40
        groups_with_grads = sum(any(p.grad is not None for p in g["params"]) for g in opt.param_groups)
        launches = lib.crv_launch_count()
        eng.step()
        assert eng._fused and eng._fused.plan is not None, "the fused step did not run"
        assert groups_with_grads == 2
        assert lib.crv_launch_count() - launches == 3            # norm + one AdamW launch per parameter group
        want_norm = torch.nn.utils.clip_grad_norm_([q for q in flat_twin if q.grad is not None], 0.05)
        ref.step()
        assert float(eng.last_grad_norm) == pytest.approx(float(want_norm), rel=1e-5)
        assert float(want_norm) > 0.05                           # the clip is active
        for p, q in zip(trainable, flat_twin):
            assert p.grad is None
            # identical maths up to fp32 rounding (torch divides by sqrt(1 - b2^t), the kernel multiplies by its inverse)
            assert torch.allclose(p.detach(), q.detach(), rtol=2e-6, atol=2e-8), float((p - q).abs().max())
            if p in opt.state and len(opt.state[p]):
                assert int(opt.state[p]["step"]) == it + 1
                for k in ("exp_avg", "exp_avg_sq"):              # m + (g - m)(1 - b1) cancels: error relative to the largest
                    a, b = opt.state[p][k], ref.state[q][k]
                    assert float((a - b).abs().max()) <= 2e-6 * float(b.abs().max()), k
        refreshed = [m for m in eng._fused.plan.mods if m is not None]
        assert len(refreshed) >= 40
        for m in refreshed:                                      # straight after the step, before any forward
            thr = m._threshold_on(m.weight_mask.device)
            assert m._wm is not None and m._wm_key == m._wm_key_now(thr)
            assert torch.equal(m._wm.float(), m._weight_bf16().float() * m.get_masks()[0])
    # the state is torch's own: a plain optimizer.step() carries on from it
    loss = eng(*batch)
    eng.backward(loss)
    opt.step()
    assert all(int(st["step"]) == 6 for st in opt.state.values() if len(st))
    sd = opt.state_dict()
    assert len(sd["state"]) >= 40
