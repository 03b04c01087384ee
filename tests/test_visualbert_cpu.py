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
