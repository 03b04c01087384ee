"""Golden vectors for the VisualBERT stage-2 path (BASELINE config 3 in miniature), from the UNMODIFIED reference
(hg_transformers/modeling_visualbert.py, masking/maskers_visualBert.py, hg_transformers/mask_trainer_visualBERT_VQA.py)
on CPU fp32:

    python tests/golden/make_golden_visualbert.py        # writes tests/golden/visualbert_tiny.pt
    python tests/golden/make_golden_visualbert.py --full # writes tests/golden/visualbert_full.pt (BASELINE config 3 at
                                                         # full size: 12 layers, h = 768, 2048-d regions, A = 3129, 56
                                                         # tokens, batch 32; statistics and samples only -- the
                                                         # seed-49 init is reproduced bit for bit by the drop-in model,
                                                         # pinned by the SHA-256 of the state_dict stored here)

Model: 2 layers, hidden 128 (2 heads of 64), 20 tokens + 36 regions, A = 50; masker: uniform zero rate 0.7 over
K,Q,V,AO,I,O x layers + pooler + word embeddings.  Stored: the reference's random init (state_dict), the batch seed,
kept counts, logits / pooled / CE loss (eval mode), the gradient of every score tensor, thresholds and kept counts
after one reference reset_threshold on perturbed scores.
"""
import importlib
import logging
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

CFG = dict(vocab_size=300, hidden_size=128, num_hidden_layers=2, num_attention_heads=2, intermediate_size=256,
           visual_embedding_dim=64, ans_num=50, max_position_embeddings=64)
ABBRES = ["K", "Q", "V", "AO", "I", "O", "P", "E"]


def batch(B=8, T=20, R=36, seed=49):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1, CFG["vocab_size"], (B, T), generator=g)
    feats = torch.randn(B, R, CFG["visual_embedding_dim"], generator=g)
    target = (torch.rand(B, CFG["ans_num"], generator=g) > 0.9).float() * torch.rand(B, CFG["ans_num"], generator=g)
    return {"ids": ids, "feats": feats, "target": target}


def _sha(sd):
    import hashlib
    h = hashlib.sha256()
    for k, v in sorted(sd.items()):
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def main():
    full = "--full" in sys.argv
    if full:
        CFG.clear()
        CFG.update(visual_embedding_dim=2048, ans_num=3129)
    R = mg.load_reference()
    VB = importlib.import_module("hg_transformers.modeling_visualbert")
    VC = importlib.import_module("hg_transformers.configuration_visualbert")
    VT = importlib.import_module("hg_transformers.mask_trainer_visualBERT_VQA")
    torch.manual_seed(49)
    cfg_obj = VC.visualBERTConfig(**CFG)
    model = VB.VisualBertForMultipleChoice(cfg_obj)
    if full:
        CFG["vocab_size"] = cfg_obj.vocab_size
        out = {"config": dict(visual_embedding_dim=2048, ans_num=3129), "state_sha": _sha(model.state_dict()),
               "vocab_size": cfg_obj.vocab_size, "B": 32}
    else:
        out = {"config": dict(CFG), "state_dict": {k: v.clone() for k, v in model.state_dict().items()}}
    conf = types.SimpleNamespace(
        masking_scheduler_conf_={"lambdas_lr": 0.0, "sparsity_warmup": "automated_gradual_sparsity",
                                 "sparsity_warmup_interval_epoch": 0.1, "init_epoch": 0.0, "final_epoch": 1.0,
                                 "final_sparsity": 0.7},
        logger=logging.getLogger("golden"), num_epochs=20)
    sched = R.sp.MaskerScheduler(conf)
    masker = R.maskers_vb.Masker(masker_scheduler=sched, logger=logging.getLogger("golden"), mask_biases=False,
                                 structured_masking_info={"structured_masking": None, "structured_masking_types": None,
                                                          "force_masking": "bert"},
                                 threshold=1e-2, init_scale=2e-2, which_ptl="visual_bert", controlled_init="magnitude")
    names = R.maskers_vb.chain_module_names("visual_bert", list(range(12)), ABBRES)
    masker.patch_modules(model=model, names_tobe_masked=names, name_of_masker="MaskedLinear1")
    mods = mg.masked_modules(model)
    out["module_names"] = [n for n, _ in mods]
    out["kept_init"] = {n: int((m.weight_mask.detach() > 1e-2).sum()) for n, m in mods}
    out["trainable"] = sorted(n for n, p in model.named_parameters() if p.requires_grad)
    b = batch(B=32) if full else batch()
    model.eval()
    model.zero_grad()
    o = model(input_ids=b["ids"], visual_embeds=b["feats"], labels=b["target"])
    loss, logits, pooled = o[0], o[1], o[2]
    loss.backward()
    out["loss"], out["logits"], out["pooled"] = loss.detach().clone(), logits.detach().clone(), pooled.detach().clone()
    def stat(g):
        flat = g.reshape(-1)
        return {"l2": float(g.double().norm()), "nnz": int((g != 0).sum()),
                "sample": flat[:: max(1, flat.numel() // (512 if full else 2048))][:(512 if full else 2048)].clone()}
    out["grad_stats"] = {n: stat(m.weight_mask.grad.detach()) for n, m in mods if m.weight_mask.grad is not None}
    out["nograd"] = [n for n, m in mods if m.weight_mask.grad is None]
    out["cls_grad_stats"] = {n: stat(p.grad.detach()) for n, p in model.named_parameters()
                             if n.startswith("cls") and p.grad is not None}
    # scores after "training": seeded noise, then the reference trainer's reset_threshold (uniform rate)
    g = torch.Generator().manual_seed(7)
    for n, m in mods:
        m.weight_mask.data.add_(torch.randn(m.weight_mask.shape, generator=g) * 5e-3)
    dummy = types.SimpleNamespace(masker=masker)
    out["mean_threshold"] = float(VT.Trainer.reset_threshold(dummy, model, 0.7))
    out["thresholds_after"] = {n: m.threshold.detach().clone() for n, m in mods}
    out["kept_after"] = {n: int((m.weight_mask.detach() > m.threshold).sum()) for n, m in mods}
    if full:                     # thresholds as Python floats keep the file small
        out["thresholds_after"] = {n: float(t) for n, t in out["thresholds_after"].items()}
    torch.save(out, os.path.join(HERE, "visualbert_full.pt" if full else "visualbert_tiny.pt"))
    print({k: (v if isinstance(v, (int, float, str)) else type(v).__name__) for k, v in out.items()})
    print("loss", float(loss), "modules", len(mods), "nograd", out["nograd"])


if __name__ == "__main__":
    main()
