"""Generate the golden vectors under tests/golden/ by running the UNMODIFIED reference modules
(/root/reference) on seeded synthetic inputs, CPU fp32.

Run in the build container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py            # writes ops.pt, tiny_lxmert.pt, full_lxmert.pt, host.json

The reference has no tests or fixtures of its own, so these files -- outputs of
the reference's own code -- are what pins the oracle (oracle/) and, through it, the CUDA path.
"""
import argparse
import importlib
import json
import logging
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shims  # noqa: E402

WEIGHT_TYPES = ["E", "VV", "VB", "lK", "lQ", "lV", "lAO", "lI", "lO", "vK", "vQ", "vV", "vAO", "vI", "vO",
                "vlVK", "vlVQ", "vlVV", "vlVAO", "vlLaK", "vlLaQ", "vlLaV", "vlLaAO", "vlVaK", "vlVaQ", "vlVaV",
                "vlVaAO", "vlLi", "vlLo", "vlVi", "vlVo", "P"]
RATES = {"Lang": 1 - 0.3, "Vis": 1 - 0.3, "Fus": 1 - 0.3, "P": 0.7}


def load_reference():
    ref_shims.install()
    mods = types.SimpleNamespace()
    mods.maskers = importlib.import_module("masking.maskers")
    mods.maskers_robust = importlib.import_module("masking.maskers_Robust")
    mods.maskers_vb = importlib.import_module("masking.maskers_visualBert")
    mods.sp = importlib.import_module("masking.sparsity_control")
    ref_shims.patch_get_init_scales(mods.maskers_robust)
    mods.lx = importlib.import_module("hg_transformers.modeling_lxmert")
    mods.cfg = importlib.import_module("hg_transformers.configuration_lxmert")
    mods.loss = importlib.import_module("hg_transformers.vqa_debias_loss_functions")
    mods.trainer = importlib.import_module("hg_transformers.mask_trainer_Robust_VQA")
    mods.trainer_base = importlib.import_module("hg_transformers.mask_trainer_VQA")
    mods.metrics = importlib.import_module("hg_transformers.data.metrics")
    mods.optim = importlib.import_module("optimization")
    return mods


class HP:  # HPmodel_modal of prune_debias_VQA.py:369-384 without nn.Module
    def __init__(self, rates):
        self.zerorate_dict = dict(rates)


def make_masker(R, model, zero_rate=0.7):
    conf = types.SimpleNamespace(
        masking_scheduler_conf_={"lambdas_lr": 0.0, "sparsity_warmup": "automated_gradual_sparsity",
                                 "sparsity_warmup_interval_epoch": 0.1, "init_epoch": 0.0, "final_epoch": 1.0,
                                 "final_sparsity": zero_rate},
        logger=logging.getLogger("golden"), num_epochs=20)
    sched = R.sp.MaskerScheduler(conf)
    masker = R.maskers_robust.Masker(
        hpmodel=HP(RATES), masker_scheduler=sched, logger=logging.getLogger("golden"), mask_biases=False,
        structured_masking_info={"structured_masking": None, "structured_masking_types": None, "force_masking": "bert"},
        threshold=1e-2, init_scale=2e-2, which_ptl="lxmert", controlled_init="magnitude")
    names, in_modal, in_module, in_layer = R.maskers_robust.chain_module_names("lxmert", list(range(12)), WEIGHT_TYPES)
    masker.names_tobe_masked = names
    masker.name_in_module = in_modal
    masker.name_of_masker = "MaskedLinear1"
    masker.patch_modules(model=model, names_tobe_masked=names, name_of_masker="MaskedLinear1")
    return masker


def synthetic_batch(B, A, seed=49, T=20, Rg=36, feat=2048, vocab=30522):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(0, vocab, (B, T), generator=g)
    feats = torch.randn(B, Rg, feat, generator=g)
    pos = torch.rand(B, Rg, 4, generator=g)
    target = (torch.rand(B, A, generator=g) > 0.999).float() * torch.rand(B, A, generator=g)
    bias = torch.rand(B, A, generator=g) * 0.01
    return {"ids": ids, "feats": feats, "pos": pos, "target": target, "bias": bias, "max_label": target.argmax(1)}


def masked_modules(model):
    return [(n, m) for n, m in model.named_modules() if hasattr(m, "threshold")]


def run_losses(R, model, batch, lmh_hidden):
    """eval-mode forward; BCE / LPF / LMH losses and their gradients w.r.t. every score tensor."""
    out = {}
    torch.manual_seed(49)
    lmh = R.loss.LearnedMixin(0.36)
    if lmh_hidden != 768:
        lmh.bias_lin = torch.nn.Linear(lmh_hidden, 1)
    out["lmh_lin_w"] = lmh.bias_lin.weight.detach().clone()
    out["lmh_lin_b"] = lmh.bias_lin.bias.detach().clone()
    out["lmh_smooth_param"] = lmh.smooth_param.detach().clone()
    mods = masked_modules(model)
    for kind in ("normal", "lpf", "lmh"):
        model.zero_grad()
        loss, logits, pooled = model(batch["ids"], batch["feats"], batch["pos"], labels=batch["target"])
        if kind == "lpf":
            loss = R.trainer.LPF_loss(logits, batch["bias"], batch["max_label"], "cpu", 5)
        elif kind == "lmh":
            loss = lmh(pooled, logits, batch["bias"], batch["target"], "cpu")
        loss.backward()
        out[f"loss_{kind}"] = loss.detach().clone()
        # the vision-side modules of the LAST cross layer feed nothing the pooler reads: grad is None
        out[f"grads_{kind}"] = {n: (m.weight_mask.grad.detach().clone() if m.weight_mask.grad is not None
                                    else torch.zeros_like(m.weight_mask)) for n, m in mods}
        out[f"nograd_{kind}"] = [n for n, m in mods if m.weight_mask.grad is None]
        out[f"cls_grads_{kind}"] = {n: p.grad.detach().clone() for n, p in model.named_parameters()
                                    if n.startswith("classifier") and p.grad is not None}
    out["logits"], out["pooled"] = logits.detach().clone(), pooled.detach().clone()
    out["score"] = R.metrics.compute_score_with_logits("vqa", logits.detach(), batch["target"])["acc"].clone()
    return out


def gen_ops(R):
    """Op-level goldens: binariser, MaskedLinear1 fwd/bwd (linear + embedding), kthvalue with ties,
    magnitude init, losses, AdamW."""
    g = torch.Generator().manual_seed(7)
    out = {}
    s = torch.rand(37, 53, generator=g) * 0.02
    s[torch.rand(37, 53, generator=g) < 0.4] = 0.0
    s[0, 0] = 1e-2
    out["bin_in"], out["bin_thr"] = s, torch.tensor(1e-2)
    out["bin_out"] = R.maskers.binarizer_fn1(s, out["bin_thr"])

    info = {"structured_masking": None, "structured_masking_types": None, "force_masking": "bert", "ptl_config": None}
    lin = torch.nn.Linear(24, 40)
    lin.weight.data = torch.randn(40, 24, generator=g) * 0.02
    lin.bias.data = torch.randn(40, generator=g) * 0.1
    ml = R.maskers.MaskedLinear1(weight=lin.weight, bias=lin.bias, mask_biases=False, name="x.dense",
                                 threshold=torch.tensor(1e-2), init_sparsity=0.7, init_scale=2e-2,
                                 controlled_init="magnitude", structured_masking_info=info)
    out["ml_weight"], out["ml_bias"] = lin.weight.detach().clone(), lin.bias.detach().clone()
    out["ml_scores_init"] = ml.weight_mask.detach().clone()
    # move scores off the {0, 0.02} lattice so the mask is non-trivial
    ml.weight_mask.data.add_(torch.randn(40, 24, generator=g) * 0.01)
    out["ml_scores"] = ml.weight_mask.detach().clone()
    x = torch.randn(3, 5, 24, generator=g, requires_grad=True)
    y = ml(x)
    dy = torch.randn(3, 5, 40, generator=g)
    y.backward(dy)
    out["ml_x"], out["ml_y"], out["ml_dy"] = x.detach().clone(), y.detach().clone(), dy
    out["ml_dx"], out["ml_ds"] = x.grad.clone(), ml.weight_mask.grad.clone()

    emb = torch.nn.Embedding(50, 16, padding_idx=0)
    emb.weight.data = torch.randn(50, 16, generator=g) * 0.02
    me = R.maskers.MaskedLinear1(weight=emb.weight, bias=None, mask_biases=False, padding_idx=0,
                                 name="emb.word_embeddings", threshold=torch.tensor(1e-2), init_sparsity=0.7,
                                 init_scale=2e-2, controlled_init="magnitude", structured_masking_info=info)
    ids = torch.randint(0, 50, (4, 7), generator=g)
    ids[0, :3] = 0
    ids[1, 2] = ids[1, 3]
    e = me(ids)
    de = torch.randn(4, 7, 16, generator=g)
    e.backward(de)
    out["emb_weight"], out["emb_scores"], out["emb_ids"] = emb.weight.detach().clone(), me.weight_mask.detach().clone(), ids
    out["emb_out"], out["emb_dout"], out["emb_ds"] = e.detach().clone(), de, me.weight_mask.grad.clone()

    # kthvalue: continuous, tie-heavy (the {0, 0.02} lattice), with +-0 and negatives
    cases = []
    a = torch.randn(10007, generator=g)
    b = torch.where(torch.rand(8192, generator=g) < 0.7, torch.zeros(8192), torch.full((8192,), 0.02))
    c = torch.cat([torch.zeros(100), -torch.zeros(100), torch.randn(300, generator=g) * 1e-3])
    d = torch.rand(3072, generator=g) * 0.02
    for t in (a, b, c, d):
        for rate in (0.7, 0.3, 1e-5, 0.99999):
            k = int(t.numel() * rate)
            k = 1 if k == 0 else k
            cases.append({"x": t, "k": k, "v": torch.kthvalue(t.view(-1), k).values.clone(),
                          "v_abs": torch.kthvalue(t.abs().view(-1), k).values.clone()})
    out["kth_cases"] = cases

    # losses on random inputs
    B, A = 6, 97
    logits = (torch.randn(B, A, generator=g) * 2).requires_grad_(True)
    labels = (torch.rand(B, A, generator=g) > 0.95).float() * torch.rand(B, A, generator=g)
    bias = torch.rand(B, A, generator=g) * 0.3
    pooled = torch.randn(B, 768, generator=g).requires_grad_(True)
    max_label = labels.argmax(1)
    torch.manual_seed(3)
    lmh = R.loss.LearnedMixin(0.36)
    rec = {"logits": logits.detach().clone(), "labels": labels, "bias": bias, "pooled": pooled.detach().clone(),
           "max_label": max_label, "lin_w": lmh.bias_lin.weight.detach().clone(),
           "lin_b": lmh.bias_lin.bias.detach().clone(), "smooth_param": lmh.smooth_param.detach().clone()}
    l = F.binary_cross_entropy_with_logits(logits, labels, reduction="mean") * labels.size(1)
    rec["bce"], rec["bce_dlogits"] = l.detach().clone(), torch.autograd.grad(l, logits)[0]
    l = R.trainer_base.LPF_loss(logits, bias, max_label, "cpu", 5)
    rec["lpf"], rec["lpf_dlogits"] = l.detach().clone(), torch.autograd.grad(l, logits)[0]
    l = lmh(pooled, logits, bias, labels, "cpu")
    gl, gp = torch.autograd.grad(l, [logits, pooled])
    rec["lmh"], rec["lmh_dlogits"], rec["lmh_dpooled"] = l.detach().clone(), gl, gp
    rec["score"] = R.metrics.compute_score_with_logits("vqa", logits.detach(), labels)["acc"].clone()
    out["loss"] = rec

    # three steps of clip_grad_norm_ + the reference AdamW on two tensors
    p1 = torch.nn.Parameter(torch.randn(11, 13, generator=g) * 0.02)
    p2 = torch.nn.Parameter(torch.randn(29, generator=g) * 0.02)
    opt = R.optim.AdamW([{"params": [p1]}, {"params": [p2]}], lr=5e-5, eps=1e-8)
    trace = {"p0": [p1.detach().clone(), p2.detach().clone()], "grads": [], "p": [], "sum": []}
    for step in range(3):
        g1 = torch.randn(11, 13, generator=g) * (10.0 if step == 0 else 1e-3)
        g2 = torch.randn(29, generator=g) * (10.0 if step == 0 else 1e-3)
        trace["grads"].append([g1.clone(), g2.clone()])
        p1.grad, p2.grad = g1, g2
        torch.nn.utils.clip_grad_norm_([p1, p2], 1.0)
        opt.step()
        trace["p"].append([p1.detach().clone(), p2.detach().clone()])
        trace["sum"].append([opt.state[p1]["sum"].clone(), opt.state[p2]["sum"].clone()])
    out["adamw"] = trace
    return out


def gen_host(R):
    """Host-logic goldens (JSON): name sets / dictionaries and sparsity schedules."""
    out = {}
    names = R.maskers.chain_module_names("lxmert", list(range(12)), WEIGHT_TYPES)
    rn, modal, module, layer = R.maskers_robust.chain_module_names("lxmert", list(range(12)), WEIGHT_TYPES)
    vb = R.maskers_vb.chain_module_names("visual_bert", list(range(12)), ["K", "Q", "V", "AO", "I", "O", "P", "E"])
    out["chain_lxmert"] = sorted(names)
    out["chain_robust"] = {"names": sorted(rn), "modal": modal, "module": module, "layer": layer}
    out["chain_visualbert"] = sorted(vb)
    f = R.sp.automated_gradual_sparsity(0.1, 0.7, 0.1, 2, 16)
    out["ags"] = [f(e, 0.0) for e in range(0, 20)]
    f = R.sp.stepwise_sparsity(0.1, 0.7, 2, 2, 16, 0.2)
    cur, seq = 0.1, []
    for e in range(0, 20):
        cur = f(e, cur)
        seq.append(cur)
    out["stepwise"] = seq
    conf = types.SimpleNamespace(
        masking_scheduler_conf_={"init_sparsity": 0.2, "final_sparsity": 0.7, "sparsity_warmup_interval_epoch": 1,
                                 "init_epoch": 1, "final_epoch": 8},
        logger=logging.getLogger("golden"), num_epochs=10)
    sch = R.sp.MaskerScheduler(conf)
    out["scheduler_steps"] = [list(sch.step(e)) for e in range(0, 10)]
    out["scheduler_is_skip"] = sch.is_skip
    return out


def gen_model(R, tiny):
    torch.manual_seed(49)
    if tiny:
        cfg = R.cfg.LxmertConfig(vocab_size=200, hidden_size=64, ans_num=50, num_attention_heads=4,
                                 intermediate_size=128, l_layers=2, x_layers=1, r_layers=1, visual_feat_dim=32,
                                 visual_pos_dim=4, max_position_embeddings=16)
        batch = synthetic_batch(4, 50, seed=49, T=6, Rg=5, feat=32, vocab=200)
    else:
        cfg = R.cfg.LxmertConfig(ans_num=2274)
        batch = synthetic_batch(32, 2274)
    model = R.lx.LxmertForMultipleChoice(cfg)
    out = {"config": {k: getattr(cfg, k) for k in ("vocab_size", "hidden_size", "ans_num", "num_attention_heads",
                                                   "intermediate_size", "l_layers", "x_layers", "r_layers",
                                                   "visual_feat_dim", "visual_pos_dim", "max_position_embeddings")}}
    if tiny:
        out["state_dict"] = {k: v.clone() for k, v in model.state_dict().items()}
        out["batch"] = batch
    masker = make_masker(R, model)
    model.eval()
    mods = masked_modules(model)
    out["module_names"] = [n for n, _ in mods]
    out["modal"] = {n: masker.name_in_module[n] for n, _ in mods}
    out["kept_init"] = {n: int((m.weight_mask.detach() > 1e-2).sum()) for n, m in mods}
    out["trainable"] = sorted(n for n, p in model.named_parameters() if p.requires_grad)
    res = run_losses(R, model, batch, cfg.hidden_size)

    # one optimiser step with the LMH gradients (last computed), then the per-modality threshold refresh
    params = [p for _, p in model.named_parameters() if p.requires_grad]
    opt = R.optim.AdamW([{"params": [p]} for p in params], lr=5e-5, eps=1e-8)
    gnorm = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()
    dummy = types.SimpleNamespace(masker=masker)
    mean_thr = R.trainer.Trainer.reset_threshold(dummy, model, 0.7)
    out["grad_norm_lmh"] = gnorm.detach().clone()
    out["mean_threshold"] = mean_thr
    out["thresholds_after"] = {n: m.threshold.detach().clone() for n, m in mods}
    out["kept_after"] = {n: int((m.weight_mask.detach() > m.threshold).sum()) for n, m in mods}
    if tiny:
        out["scores_init"] = {n: None for n, _ in mods}  # filled below from init masks
        out["scores_after"] = {n: m.weight_mask.detach().clone() for n, m in mods}
        out.update(res)
    else:
        # full model: keep the small tensors plus per-module statistics and a strided sample of dS
        for k in ("logits", "pooled", "score", "loss_normal", "loss_lpf", "loss_lmh", "lmh_lin_w", "lmh_lin_b",
                  "lmh_smooth_param", "nograd_lmh"):
            out[k] = res[k]
        for kind in ("normal", "lpf", "lmh"):
            gs = res[f"grads_{kind}"]
            out[f"grad_stats_{kind}"] = {n: {"l2": float(g.double().norm()), "abs_mean": float(g.abs().mean()),
                                             "nnz": int((g != 0).sum()),
                                             "sample": g.reshape(-1)[:: max(1, g.numel() // 512)][:512].clone()}
                                         for n, g in gs.items()}
            out[f"cls_grads_{kind}"] = {n: {"l2": float(g.double().norm())} for n, g in res[f"cls_grads_{kind}"].items()}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-full", action="store_true")
    args = ap.parse_args()
    logging.basicConfig(level=logging.WARNING)
    R = load_reference()
    torch.save(gen_ops(R), os.path.join(HERE, "ops.pt"))
    with open(os.path.join(HERE, "host.json"), "w") as f:
        json.dump(gen_host(R), f)
    tiny = gen_model(R, tiny=True)
    tiny.pop("scores_init")
    torch.save(tiny, os.path.join(HERE, "tiny_lxmert.pt"))
    if not args.skip_full:
        torch.save(gen_model(R, tiny=False), os.path.join(HERE, "full_lxmert.pt"))
    print("golden vectors written to", HERE)


if __name__ == "__main__":
    main()
