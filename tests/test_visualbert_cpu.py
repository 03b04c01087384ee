"""VisualBERT stage-2 path, CPU tier: oracle/visualbert_oracle.py against the reference's outputs
(tests/golden/visualbert_tiny.pt, make_golden_visualbert.py)."""
import os

import torch

from oracle import lxmert_oracle as lxo
from oracle import masked_ops as o_ops
from oracle import visualbert_oracle as vbo

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _batch(cfg, B=8, T=20, R=36, seed=49):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1, cfg["vocab_size"], (B, T), generator=g)
    feats = torch.randn(B, R, cfg["visual_embedding_dim"], generator=g)
    target = (torch.rand(B, cfg["ans_num"], generator=g) > 0.9).float() * torch.rand(B, cfg["ans_num"], generator=g)
    return ids, feats, target


def test_visualbert_oracle_matches_reference():
    g = torch.load(os.path.join(GOLD, "visualbert_tiny.pt"), weights_only=False)
    cfg = g["config"]
    L = cfg["num_hidden_layers"]
    names = vbo.module_names(L)
    assert names == g["module_names"]
    params = {k: v.clone() for k, v in g["state_dict"].items()}
    for k in params:
        if params[k].dtype.is_floating_point:
            params[k].requires_grad_(k.startswith("cls."))
    scores, thr = {}, {}
    for n in names:
        s, _ = o_ops.magnitude_init(params[n + ".weight"], 0.7, 1e-2)
        scores[n] = s.requires_grad_(True)
        thr[n] = 1e-2
    assert {n: int((s > 1e-2).sum()) for n, s in scores.items()} == g["kept_init"]
    ids, feats, target = _batch(cfg)
    c = lxo.Ctx(params, scores, thr, heads=cfg["num_attention_heads"])
    logits, pooled = vbo.forward(c, ids, feats, L)
    loss = vbo.soft_cross_entropy(logits, target)
    loss.backward()
    torch.testing.assert_close(loss.detach(), g["loss"], rtol=1e-6, atol=0)
    torch.testing.assert_close(logits.detach(), g["logits"], rtol=1e-4, atol=2e-6)
    torch.testing.assert_close(pooled.detach(), g["pooled"], rtol=1e-4, atol=2e-6)
    assert g["nograd"] == []
    for n in names:
        st, gr = g["grad_stats"][n], scores[n].grad
        assert abs(float(gr.double().norm()) - st["l2"]) <= 1e-4 * st["l2"] + 1e-12, n
        assert int((gr != 0).sum()) == st["nnz"], n
        flat = gr.reshape(-1)
        torch.testing.assert_close(flat[:: max(1, flat.numel() // 2048)][:2048], st["sample"], rtol=1e-3,
                                   atol=1e-5 * float(st["sample"].abs().max()) + 1e-12)
    for n, st in g["cls_grad_stats"].items():
        assert abs(float(params[n].grad.double().norm()) - st["l2"]) <= 1e-4 * st["l2"] + 1e-12, n
    # threshold refresh on perturbed scores (uniform zero rate 0.7)
    gen = torch.Generator().manual_seed(7)
    after = {n: scores[n].detach() + torch.randn(scores[n].shape, generator=gen) * 5e-3 for n in names}
    for n in names:
        k = max(1, int(after[n].numel() * 0.7))
        t = o_ops.kth_value(after[n], k)
        assert float(t) == float(g["thresholds_after"][n]), n
        assert int((after[n] > float(t)).sum()) == g["kept_after"][n], n


def test_visualbert_driver_init_masker_host_logic(monkeypatch):
    """prune_debias_VQA_visualBERT.init_masker (uniform zero rate, K/Q/V/AO/I/O/P/E) with the oracle as the fake kernel
    backend: same module census, trainable set and initial kept counts as the reference masker produced."""
    import logging

    import pytest
    from crvqa import ops
    from hg_transformers.modeling_visualbert import VisualBertForMultipleChoice, visualBERTConfig
    from prune_debias_VQA_visualBERT import ModelArguments, init_masker

    def kth(tensors, ks, use_abs=False):
        return torch.tensor([float(o_ops.kth_value(t, int(k), use_abs=use_abs)) for t, k in zip(tensors, ks)])

    def mag(weight, w_thr, hi, lo):
        keep = weight.detach().abs() > float(w_thr)
        return torch.where(keep, torch.full_like(weight, hi), torch.full_like(weight, lo))

    monkeypatch.setattr(ops, "kth_value_batched", kth)
    monkeypatch.setattr(ops, "magnitude_init", mag)
    monkeypatch.setattr(ops, "binarize", lambda s, t, want_count=False, as_bool=False: o_ops.binarize(s.detach(), float(t)))
    g = torch.load(os.path.join(GOLD, "visualbert_tiny.pt"), weights_only=False)
    model = VisualBertForMultipleChoice(visualBERTConfig(**g["config"]))
    model.load_state_dict(g["state_dict"], strict=True)
    log = logging.getLogger("t")
    log.setLevel(logging.ERROR)
    margs = ModelArguments(model_type="visual_bert", zero_rate=0.7)
    masker = init_masker(margs, model, log)
    mods = [(n, m) for n, m in model.named_modules() if hasattr(m, "threshold")]
    assert [n for n, _ in mods] == g["module_names"]
    assert sorted(n for n, p in model.named_parameters() if p.requires_grad) == g["trainable"]
    assert {n: int((m.weight_mask > 1e-2).sum()) for n, m in mods} == g["kept_init"]
    assert masker.masker_scheduler.init_sparsity == 0.7 and margs.layers_to_mask_ == list(range(12))
    margs2 = ModelArguments(model_type="visual_bert", mask_classifier=True)
    with pytest.raises(AssertionError):
        init_masker(margs2, VisualBertForMultipleChoice(visualBERTConfig(**g["config"])), log)
