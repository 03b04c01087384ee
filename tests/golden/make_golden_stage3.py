"""Golden vectors for the stage-3 frozen-mask fine-tune, produced by running the
UNMODIFIED reference code on CPU fp32 in the build container:

    python tests/golden/make_golden_stage3.py        # writes tests/golden/stage3_full.pt

What runs: the reference's own LXMERT (hg_transformers/modeling_lxmert.py, seed-49 init, full 9/5/5 config,
A=2274), and the reference's own `pruning_model_with_mask`, `mag_pruning` and `see_weight_rate`
(run_vqa_stage3.py:75-300).  run_vqa_stage3.py cannot be imported as a module here (its top level pulls the
datasets and the auto-model zoo), so the three FunctionDefs are cut out of its source with `ast` and executed
unchanged in a namespace that holds only `torch` and `torch.nn.utils.prune`.

The trained mask of stage 2 is replaced by the magnitude mask of the seed-49 weights at zero rate 0.7
(`|W| > kthvalue(|W|, int(0.7 n))`), which the GPU side can rebuild exactly (the model init is bit-identical and
the select is exact), so no 207 M-element mask has to be stored.
"""
import ast
import os
import sys
import types

import torch
import torch.nn.utils.prune as prune

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

REF = "/root/reference"
WANTED = ("see_weight_rate", "mag_pruning", "pruning_model_with_mask")


def reference_stage3_functions():
    src = open(os.path.join(REF, "run_vqa_stage3.py")).read()
    tree = ast.parse(src)
    ns = {"torch": torch, "prune": prune, "print": lambda *a, **k: None}
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name in WANTED:
            code = compile(ast.Module(body=[node], type_ignores=[]), "run_vqa_stage3.py", "exec")
            exec(code, ns)
    return types.SimpleNamespace(**{k: ns[k] for k in WANTED})


def pruned_module_names(model):
    """Modules that carry a prune reparametrisation, in named_modules order, without the 'lxmert.' prefix."""
    return [n for n, m in model.named_modules() if hasattr(m, "weight_orig")]


def sample(t, n=512):
    flat = t.reshape(-1)
    return flat[:: max(1, flat.numel() // n)][:n].clone()


def main():
    R = mg.load_reference()
    S3 = reference_stage3_functions()
    B, A = 8, 2274
    out = {"B": B, "A": A, "zero_rate": 0.7}

    # ---- FT_trainedMask: CustomFromMask with a given mask dict ------------------------------------
    torch.manual_seed(49)
    model = R.lx.LxmertForMultipleChoice(R.cfg.LxmertConfig(ans_num=A))
    bert = model.lxmert
    mask = {}
    names = []
    for n, m in bert.named_modules():
        if isinstance(m, (torch.nn.Linear, torch.nn.Embedding)) and hasattr(m, "weight"):
            names.append(n)
    # the modules pruning_model_with_mask touches are exactly the stage-2 masked set: build masks for all Linear /
    # word-embedding modules it names (keys: '<model_type>.<module>.weight_mask', as save_model_mask writes them)
    wanted = []
    for ii in range(9):
        for s in ("attention.self.query", "attention.self.key", "attention.self.value", "attention.output.dense",
                  "intermediate.dense", "output.dense"):
            wanted.append(f"encoder.layer.{ii}.{s}")
    for ii in range(5):
        for s in ("attention.self.query", "attention.self.key", "attention.self.value", "attention.output.dense",
                  "intermediate.dense", "output.dense"):
            wanted.append(f"encoder.r_layers.{ii}.{s}")
    for ii in range(5):
        for s in ("visual_attention.att.query", "visual_attention.att.key", "visual_attention.att.value",
                  "visual_attention.output.dense", "lang_self_att.self.query", "lang_self_att.self.key",
                  "lang_self_att.self.value", "lang_self_att.output.dense", "visn_self_att.self.query",
                  "visn_self_att.self.key", "visn_self_att.self.value", "visn_self_att.output.dense",
                  "lang_inter.dense", "lang_output.dense", "visn_inter.dense", "visn_output.dense"):
            wanted.append(f"encoder.x_layers.{ii}.{s}")
    wanted += ["pooler.dense", "embeddings.word_embeddings", "encoder.visn_fc.visn_fc", "encoder.visn_fc.box_fc"]
    mods = dict(bert.named_modules())
    kept = {}
    for n in wanted:
        w = mods[n].weight.detach()
        k = max(1, int(w.numel() * 0.7))
        thr = torch.kthvalue(w.abs().reshape(-1), k).values
        m = (w.abs() > thr)
        mask[f"lxmert.{n}.weight_mask"] = m
        kept[n] = int(m.sum())
    out["mask_rule"] = "mask = |W| > kthvalue(|W|.view(-1), max(1, int(0.7 * numel)))  on the seed-49 weights"
    out["kept"] = kept
    S3.pruning_model_with_mask(bert, mask, "lxmert")
    out["pruned_modules"] = pruned_module_names(bert)
    out["zero_rate_pct"] = float(S3.see_weight_rate(model, "lxmert"))
    out["state_keys_sample"] = sorted(k for k in model.state_dict().keys() if "layer.0.attention.self.query" in k)
    out["trainable"] = sorted(n for n, p in model.named_parameters() if p.requires_grad)

    batch = mg.synthetic_batch(B, A)
    model.eval()
    torch.manual_seed(49)
    lmh = R.loss.LearnedMixin(0.36)
    out["lmh_lin_w"] = lmh.bias_lin.weight.detach().clone()
    out["lmh_lin_b"] = lmh.bias_lin.bias.detach().clone()
    out["lmh_smooth_param"] = lmh.smooth_param.detach().clone()
    for kind in ("normal", "lmh"):
        model.zero_grad()
        loss, logits, pooled = model(batch["ids"], batch["feats"], batch["pos"], labels=batch["target"])
        if kind == "lmh":
            loss = lmh(pooled, logits, batch["bias"], batch["target"], "cpu")
        loss.backward()
        out[f"loss_{kind}"] = loss.detach().clone()
        stats = {}
        for n, p in model.named_parameters():
            if p.grad is None:
                continue
            g = p.grad.detach()
            stats[n] = {"l2": float(g.double().norm()), "nnz": int((g != 0).sum()), "sample": sample(g)}
        out[f"grad_stats_{kind}"] = stats
        out[f"nograd_{kind}"] = sorted(n for n, p in model.named_parameters() if p.requires_grad and p.grad is None)
    out["logits"], out["pooled"] = logits.detach().clone(), pooled.detach().clone()
    # a masked weight gets no gradient where the mask is zero: record it for one module as a known answer
    q = mods["encoder.layer.0.attention.self.query"]
    out["grad_zero_where_masked"] = bool((q.weight_orig.grad[q.weight_mask == 0] == 0).all())

    # one Adam step (run_vqa_stage3.init_optimizer: torch.optim.Adam, lr 5e-5, eps 1e-8) on the LMH gradients
    opt = torch.optim.Adam([{"params": [p]} for p in model.parameters() if p.requires_grad], lr=5e-5, betas=(0.9, 0.999),
                           eps=1e-8)
    gnorm = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
    opt.step()
    out["grad_norm_lmh"] = float(gnorm)
    out["after_step"] = {n: {"l2": float(p.detach().double().norm()), "sample": sample(p.detach())}
                         for n, p in model.named_parameters()
                         if n in ("lxmert.encoder.layer.0.attention.self.query.weight_orig",
                                  "lxmert.encoder.x_layers.4.lang_output.dense.weight_orig",
                                  "lxmert.encoder.layer.0.attention.output.LayerNorm.weight",
                                  "lxmert.pooler.dense.bias")}

    # ---- FT_randMask: the reference's mag_pruning (prune.l1_unstructured at amount px) ----------------
    torch.manual_seed(49)
    model2 = R.lx.LxmertForMultipleChoice(R.cfg.LxmertConfig(ans_num=A))
    S3.mag_pruning(model2.lxmert, 0.7)
    pm = {n: m for n, m in model2.lxmert.named_modules() if hasattr(m, "weight_mask")}
    out["mag_pruned_modules"] = sorted(pm)
    out["mag_kept"] = {n: int(m.weight_mask.sum()) for n, m in pm.items()}
    out["mag_mask_sample"] = {n: sample(pm[n].weight_mask).bool() for n in
                              ("encoder.layer.0.attention.self.query", "encoder.layer.8.output.dense",
                               "embeddings.word_embeddings", "pooler.dense")}
    torch.save(out, os.path.join(HERE, "stage3_full.pt"))
    print("wrote stage3_full.pt:", {k: (v if isinstance(v, (int, float, str, bool)) else type(v).__name__)
                                    for k, v in out.items()})


if __name__ == "__main__":
    main()
