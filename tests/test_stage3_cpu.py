"""Stage-3 frozen-mask fine-tune, CPU tier: pins oracle/stage3.py against tests/golden/stage3_full.pt, the
outputs of the reference's own pruning functions and LXMERT (tests/golden/make_golden_stage3.py), and checks
the host logic of the drop-in driver (run_vqa_stage3.py) that needs no kernel."""
import os

import pytest
import torch

from oracle import lxmert_oracle as lxo
from oracle import losses as o_losses
from oracle import stage3 as o3

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD, "stage3_full.pt"), weights_only=False)


@pytest.fixture(scope="module")
def seed49_params():
    from hg_transformers.modeling_lxmert import LxmertConfig, LxmertForMultipleChoice
    torch.manual_seed(49)
    model = LxmertForMultipleChoice(LxmertConfig(ans_num=2274))
    return {k: v.detach().clone() for k, v in model.state_dict().items()}


def _sample(t, n=512):
    flat = t.reshape(-1)
    return flat[:: max(1, flat.numel() // n)][:n]


def test_pruned_linear_backward_is_what_autograd_derives():
    torch.manual_seed(0)
    x = torch.randn(3, 5, 16, requires_grad=True)
    w = torch.randn(8, 16, requires_grad=True)
    m = (torch.rand(8, 16) > 0.6).float()
    b = torch.randn(8, requires_grad=True)
    dy = torch.randn(3, 5, 8)
    y = o3.pruned_linear(x, w, m, b)
    y.backward(dy)
    got = (x.grad.clone(), w.grad.clone(), b.grad.clone())
    x.grad = w.grad = b.grad = None
    ref = torch.nn.functional.linear(x, w * m, b)          # the literal graph of the prune forward pre-hook
    ref.backward(dy)
    torch.testing.assert_close(y, ref, rtol=1e-6, atol=1e-6)
    for g, r in zip(got, (x.grad, w.grad, b.grad)):
        torch.testing.assert_close(g, r, rtol=1e-5, atol=1e-6)
    assert bool((got[1][m == 0] == 0).all())


def test_module_lists_match_reference(gold):
    assert ["lxmert." + n for n in []] == []
    assert sorted(o3.trained_mask_modules()) == sorted(gold["pruned_modules"])
    existing = set(gold["pruned_modules"])
    assert sorted(o3.mag_pruning_modules(existing)) == gold["mag_pruned_modules"]
    import run_vqa_stage3 as s3
    assert s3.trained_mask_module_names() == o3.trained_mask_modules()


def test_masks_zero_rate_and_l1_unstructured(gold, seed49_params):
    P = seed49_params
    masks = {n: o3.magnitude_mask(P[f"lxmert.{n}.weight"], 0.7) for n in o3.trained_mask_modules()}
    assert {n: int(m.sum()) for n, m in masks.items()} == gold["kept"]
    assert o3.see_weight_rate(masks) == pytest.approx(gold["zero_rate_pct"], rel=0, abs=1e-9)
    for n in ("encoder.layer.0.attention.self.query", "encoder.layer.8.output.dense", "pooler.dense"):
        m = o3.l1_unstructured_mask(P[f"lxmert.{n}.weight"], 0.7)
        assert int(m.sum()) == gold["mag_kept"][n]
        assert torch.equal(_sample(m).bool(), gold["mag_mask_sample"][n])
    # the word-embedding table has a zero padding row: its 768 tied magnitudes are all below the cut
    m = o3.l1_unstructured_mask(P["lxmert.embeddings.word_embeddings.weight"], 0.7)
    assert int(m.sum()) == gold["mag_kept"]["embeddings.word_embeddings"]


def test_full_model_forward_backward_and_adam_step_match_reference(gold, seed49_params):
    """B=8, A=2274, eval mode: logits, BCE and LMH losses, the gradient of EVERY trainable tensor (weight_orig
    of the 168 pruned modules, biases, LayerNorms, embeddings, weight-normed head) and one Adam step."""
    P = seed49_params
    names = o3.trained_mask_modules()
    masks = {"lxmert." + n: o3.magnitude_mask(P[f"lxmert.{n}.weight"], 0.7).float() for n in names}
    params = {}
    for k, v in P.items():
        mod = k[: -len(".weight")] if k.endswith(".weight") else None
        key = mod + ".weight_orig" if mod in masks else k
        params[key] = v.clone().requires_grad_(v.dtype.is_floating_point)
    assert sorted(k for k, v in params.items() if v.requires_grad) == gold["trainable"]
    batch = lxo.synthetic_batch(gold["B"], gold["A"])
    lmh = {"lin_w": gold["lmh_lin_w"], "lin_b": gold["lmh_lin_b"], "smooth_param": gold["lmh_smooth_param"]}
    for kind in ("normal", "lmh"):
        for v in params.values():
            v.grad = None
        logits, pooled = o3.forward(params, masks, batch)
        loss = lxo.compute_loss(kind, logits, pooled, batch, lmh=lmh)
        loss.backward()
        torch.testing.assert_close(loss.detach(), gold[f"loss_{kind}"], rtol=2e-6, atol=0)
        stats = gold[f"grad_stats_{kind}"]
        assert sorted(k for k, v in params.items() if v.requires_grad and v.grad is None) == gold[f"nograd_{kind}"]
        for k, st in stats.items():
            g = params[k].grad
            if k.endswith("key.bias"):
                # softmax is invariant to a per-query constant, so d/d(key bias) is identically zero; both sides hold
                # fp32 cancellation noise (1e-7) that has no reason to agree
                assert float(g.double().norm()) < 1e-5 and st["l2"] < 1e-5, k
                continue
            assert abs(float(g.double().norm()) - st["l2"]) <= 2e-4 * st["l2"] + 1e-12, k
            if k.endswith("weight_orig"):
                assert int((g != 0).sum()) == st["nnz"], k                 # zero exactly where the mask is zero
            torch.testing.assert_close(_sample(g), st["sample"], rtol=2e-3,
                                       atol=2e-5 * float(st["sample"].abs().max()) + 1e-12, msg=k)
    torch.testing.assert_close(logits.detach(), gold["logits"], rtol=1e-4, atol=2e-6)
    torch.testing.assert_close(pooled.detach(), gold["pooled"], rtol=1e-4, atol=2e-6)
    q = "lxmert.encoder.layer.0.attention.self.query"
    assert bool((params[q + ".weight_orig"].grad[masks[q] == 0] == 0).all()) and gold["grad_zero_where_masked"]

    # clip_grad_norm_(1.0) + torch.optim.Adam step on the LMH gradients
    grads = [v.grad for v in params.values() if v.requires_grad and v.grad is not None]
    gnorm = float(torch.sqrt(sum(g.double().pow(2).sum() for g in grads)))
    assert gnorm == pytest.approx(gold["grad_norm_lmh"], rel=1e-4)
    coef = min(1.0, 1.0 / (gnorm + 1e-6))
    for k, st in gold["after_step"].items():
        p = params[k].detach().clone()
        g = params[k].grad * coef
        o3.adam_step(p, g, torch.zeros_like(p), torch.zeros_like(p), 1)
        assert abs(float(p.double().norm()) - st["l2"]) <= 1e-6 * st["l2"], k
        torch.testing.assert_close(_sample(p), st["sample"], rtol=1e-5, atol=1e-7, msg=k)


def test_driver_functions_need_no_kernel_for_given_masks(gold):
    """pruning_model_with_mask / see_weight_rate on CPU modules: same state_dict keys, same trainable set and
    the same zero rate as the reference (the forward of the pruned modules is CUDA-only and is not called)."""
    import run_vqa_stage3 as s3
    from hg_transformers.modeling_lxmert import LxmertConfig, LxmertForMultipleChoice
    torch.manual_seed(49)
    model = LxmertForMultipleChoice(LxmertConfig(ans_num=2274))
    sd = model.state_dict()
    mask = {f"lxmert.{n}.weight_mask": o3.magnitude_mask(sd[f"lxmert.{n}.weight"], 0.7) for n in s3.trained_mask_module_names()}
    s3.pruning_model_with_mask(model.lxmert, mask, "lxmert")
    assert s3.see_weight_rate(model, "lxmert") == pytest.approx(gold["zero_rate_pct"], rel=0, abs=1e-9)
    assert sorted(k for k in model.state_dict() if "layer.0.attention.self.query" in k) == gold["state_keys_sample"]
    assert sorted(n for n, p in model.named_parameters() if p.requires_grad) == gold["trainable"]
    assert sorted(n for n, m in model.lxmert.named_modules() if hasattr(m, "weight_orig")) == sorted(gold["pruned_modules"])
    q = model.lxmert.encoder.layer[0].attention.self.query
    torch.testing.assert_close(q.weight, q.weight_orig * q.weight_mask)
    with pytest.raises(RuntimeError):
        q(torch.zeros(2, 768))                                   # no CPU fallback


def test_grouped_adam_is_torch_adam():
    """run_vqa_stage3.init_optimizer's GroupedAdam steps all one-tensor groups with one multi-tensor call: same
    numbers, state and schedule behaviour as torch.optim.Adam built the reference's way."""
    import run_vqa_stage3 as s3
    torch.manual_seed(0)
    ps = [torch.nn.Parameter(torch.randn(s)) for s in [(5, 3), (7,), (2, 2, 2)]]
    qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
    a = torch.optim.Adam([{"params": [p], "name": str(i)} for i, p in enumerate(ps)], lr=1e-3, eps=1e-8)
    b = s3.GroupedAdam([{"params": [p], "name": str(i)} for i, p in enumerate(qs)], lr=1e-3, eps=1e-8)
    sa = torch.optim.lr_scheduler.LambdaLR(a, lambda s: 1 - s / 10)
    sb = torch.optim.lr_scheduler.LambdaLR(b, lambda s: 1 - s / 10)
    for _ in range(4):
        for p, q in zip(ps, qs):
            g = torch.randn_like(p)
            p.grad, q.grad = g.clone(), g.clone()
        a.step(); b.step(); sa.step(); sb.step()
    for p, q in zip(ps, qs):
        assert torch.equal(p, q)
        assert torch.equal(a.state[p]["exp_avg_sq"], b.state[q]["exp_avg_sq"])
    assert [g["name"] for g in b.param_groups] == ["0", "1", "2"]
