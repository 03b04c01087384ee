"""Parity of the path that is BENCHMARKED -- score arena + mask cache + 2-CTA tcgen05 GEMMs + fused LayerNorm / GELU /
attention kernels + the whole-step CUDA graph -- against outputs of the unmodified reference
(hg_transformers/modeling_lxmert.py:1060-1120 + masking/maskers.py:337-366), at BASELINE config 1
(tests/golden/full_lxmert.pt) and config 2 (tests/golden/config2_lxmert.pt, the bench workload), plus

  * the tolerance story as evidence: with every MMA operand carried as hi + lo bf16 halves (CRVQA_OPERAND=split,
    same kernels, ~16 mantissa bits) the end-to-end gap to the fp32 reference falls below the north_star's 2e-3,
    so what the bf16 product path shows is operand rounding, not a kernel defect;
  * mask agreement with the reference TRAJECTORY (IoU as the reference's compare_mask.py:31-43) after 1 and 6
    optimiser steps (tests/golden/trajectory_small.pt).

Measured numbers (B200, this tree) are quoted next to each tolerance."""
import hashlib
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
REPORT = {}


def _sha(model):
    h = hashlib.sha256()
    for k, v in sorted(model.state_dict().items()):
        if "weight_mask" in k:
            continue
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def _strided(t, n=512):
    f = t.reshape(-1)
    return f[:: max(1, f.numel() // n)][:n]


def _mods(model):
    return [(n, m) for n, m in model.named_modules() if hasattr(m, "threshold")]


def _grad_report(mods, stats, nograd):
    """worst relative gap of per-module gradient L2 norms, worst norm-wise gap over the stored 512-element samples,
    and the pooled (all modules) sample gap."""
    worst_l2, worst_s, num, den = (0.0, ""), (0.0, ""), 0.0, 0.0
    for n, m in mods:
        st = stats[n]
        g = m.weight_mask.grad
        if n in nograd or st["l2"] == 0.0:
            assert g is None or float(g.abs().max()) == 0.0, n
            continue
        l2 = float(g.double().norm())
        worst_l2 = max(worst_l2, (abs(l2 - st["l2"]) / st["l2"], n))
        got = _strided(g.detach()).cpu().double()
        ref = st["sample"].double()
        d2, r2 = float((got - ref).pow(2).sum()), float(ref.pow(2).sum())
        if r2 > 0:
            worst_s = max(worst_s, ((d2 / r2) ** 0.5, n))
        num += d2 / max(st["l2"] ** 2, 1e-300)     # every module weighted equally
        den += r2 / max(st["l2"] ** 2, 1e-300)
    return worst_l2, worst_s, (num / den) ** 0.5


def _build(ans, cfg=None, arena=False, dropout_off=False, seed=49):
    from hg_transformers._engine import ScoreArena, execution_order, masked_modules_of
    from prune_debias_VQA import build_stage2
    cfg = dict(cfg or {})
    if dropout_off:
        cfg.update(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0)
    model, masker, margs = build_stage2(ans, device=torch.device("cuda"), seed=seed, config_kwargs=cfg)
    model.eval()
    ar = None
    if arena:
        ar = ScoreArena(execution_order(masked_modules_of(model)))
        ar.enable_mask_cache()
        assert model.lxmert.encoder._fast_plans() is not None
    return model, masker, margs, ar


def _fwd_bwd(model, b, kind, g=None, arena=None):
    from crvqa import ops
    if arena is not None:
        arena.begin_step()
    else:
        model.zero_grad()
    _, logits, pooled = model(b["ids"], b["feats"], b["pos"], labels=b["target"])
    if kind == "normal":
        loss, score = ops.vqa_loss_bce(logits, b["target"])
    elif kind == "lpf":
        loss, score = ops.vqa_loss_lpf(logits, b["bias"], b["max_label"], 5.0, b["target"])
    else:
        from hg_transformers.vqa_debias_loss_functions import LearnedMixin
        lm = LearnedMixin(0.36).cuda()
        lm.bias_lin.weight.data.copy_(g["lmh_lin_w"])
        lm.bias_lin.bias.data.copy_(g["lmh_lin_b"])
        loss = lm(pooled, logits, b["bias"], b["target"], "cuda")
        score = lm.last_score
    loss.backward()
    if arena is not None:
        arena.finalize_grads()
    return loss.detach(), logits.detach(), pooled.detach(), score


# ---------------------------------------------------------------------------------------------------------------
# config 1 through the fast path
# ---------------------------------------------------------------------------------------------------------------
def test_fast_path_config1_vs_reference_golden():
    """full_lxmert.pt (B=32, A=2274) through _forward_fast + masked_gemm2_kernel.  Tolerances = measured bf16
    floor of this path x ~2 (see test_split_operands_close_the_gap for why it is the floor):
    loss 1e-3 (LMH 2e-3), logits 1e-2 of max|logit|, gradient L2 norms 3e-2, sampled gradients 0.15 norm-wise."""
    from oracle import lxmert_oracle as lxo
    g = torch.load(os.path.join(GOLD, "full_lxmert.pt"), weights_only=False)
    model, _, _, arena = _build(2274, arena=True)
    mods = _mods(model)
    assert [n for n, _ in mods] == g["module_names"]
    b = {k: v.cuda() for k, v in lxo.synthetic_batch(32, 2274).items()}
    from crvqa import lib
    c0 = lib.crv_launch_count()
    for kind in ("normal", "lpf", "lmh"):
        loss, logits, pooled, score = _fwd_bwd(model, b, kind, g, arena)
        ref = float(g[f"loss_{kind}"])
        lerr = abs(float(loss) - ref) / abs(ref)
        err = float((logits.cpu() - g["logits"]).abs().max() / g["logits"].abs().max())
        perr = float((pooled.cpu() - g["pooled"]).abs().max() / g["pooled"].abs().max())
        wl2, ws, pooled_s = _grad_report(mods, g[f"grad_stats_{kind}"], g["nograd_lmh"])
        REPORT[f"config1 fast {kind}"] = dict(loss=lerr, logits=err, pooled=perr, grad_l2=wl2, grad_sample=ws,
                                              grad_sample_all=pooled_s)
        print(f"[config1 fast {kind}] loss {lerr:.2e} logits {err:.2e} pooled {perr:.2e} grad-L2 {wl2} "
              f"sample {ws} all-modules sample {pooled_s:.3e}")
        assert lerr < (2e-3 if kind == "lmh" else 1e-3)
        assert err < 1e-2
        assert wl2[0] < 3e-2, wl2
        assert ws[0] < 0.15, ws
        assert pooled_s < 0.1
    assert float(score) == float(g["score"])
    assert lib.crv_launch_count() - c0 > 600           # the fused kernels ran (not the torch modules)


# ---------------------------------------------------------------------------------------------------------------
# the tolerance story: raise operand precision, same kernels, gap collapses below 2e-3
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("cached", [False, True])
def test_split_operands_close_the_gap(cached, monkeypatch):
    """CRVQA_OPERAND=split carries every MMA operand as hi + lo bf16 halves (three launches of the SAME kernels per
    GEMM).  cached=False: the in-kernel mask-transform 1-CTA kernels; cached=True: plain 2-CTA kernels on the arena's
    materialised W (.) M.  Against the fp32 reference golden the north_star bar holds end to end: logits, losses and
    score gradients within 2e-3 relative.  The same comparison with bf16 operands (product path) gives 5e-3 / 3-9e-2
    -- i.e. that gap is operand rounding through 19 layers, not the kernels."""
    from hg_transformers._engine import ScoreArena, execution_order, masked_modules_of
    from oracle import lxmert_oracle as lxo
    monkeypatch.setenv("CRVQA_OPERAND", "split")
    monkeypatch.setenv("CRVQA_FUSED", "0")             # per-module path: fp32 activations between the GEMMs
    # the attention cores and the answer head of the per-module path are torch matmuls: strict fp32 for this test
    # (a Trainer constructed earlier in the process switches TF32 on for the answer head)
    monkeypatch.setattr(torch.backends.cuda.matmul, "allow_tf32", False)
    monkeypatch.setattr(torch.backends.cudnn, "allow_tf32", False)
    g = torch.load(os.path.join(GOLD, "full_lxmert.pt"), weights_only=False)
    model, _, _, _ = _build(2274)
    arena = None
    if cached:
        arena = ScoreArena(execution_order(masked_modules_of(model)))
        arena.enable_mask_cache()
        assert model.lxmert.encoder._fast_plans() is None
    mods = _mods(model)
    b = {k: v.cuda() for k, v in lxo.synthetic_batch(32, 2274).items()}
    for kind in ("lpf", "normal"):
        loss, logits, pooled, _ = _fwd_bwd(model, b, kind, g, arena)
        ref = float(g[f"loss_{kind}"])
        lerr = abs(float(loss) - ref) / abs(ref)
        err = float((logits.cpu() - g["logits"]).abs().max() / g["logits"].abs().max())
        wl2, ws, pooled_s = _grad_report(mods, g[f"grad_stats_{kind}"], g["nograd_lmh"])
        REPORT[f"config1 split cached={cached} {kind}"] = dict(loss=lerr, logits=err, grad_l2=wl2, grad_sample=ws,
                                                               grad_sample_all=pooled_s)
        print(f"[config1 split-operand cached={cached} {kind}] loss {lerr:.2e} logits {err:.2e} grad-L2 {wl2} "
              f"sample {ws} all-modules sample {pooled_s:.3e}")
        # measured (B200): LPF loss 1.3e-7, logits 6.4e-5, gradient L2 norms <= 1.9e-4, sampled gradients 8.0e-4 over
        # all modules (worst single module 1.7e-3) -- against 1.1e-5 / 5.2e-3 / 1.2e-2 / 5.8e-2 (9e-2) on the bf16 path
        assert lerr < 2e-3 and err < 2e-3
        assert wl2[0] < 2e-3, wl2
        if kind == "lpf":
            assert pooled_s < 2e-3, pooled_s
            assert ws[0] < 2e-3, ws
        else:
            # BCE back-propagates a dense, nearly uniform dlogits (sigmoid(x) / B over 2274 answers): the sum over
            # answers cancels to ~1e-3 of its terms, so fp32 summation order alone (ours vs the CPU reference's)
            # shows at 5e-3 in the sampled entries; norms and logits still meet the bar
            assert pooled_s < 1e-2, pooled_s
            assert ws[0] < 5e-2, ws


# ---------------------------------------------------------------------------------------------------------------
# config 2 (the bench workload) through the fast path, eager and as one replayed CUDA graph
# ---------------------------------------------------------------------------------------------------------------
def _config2_trainer(tmp_path, graph):
    from hg_transformers.data.data_collator import TrimCollator
    from hg_transformers.data.metrics import vqa_compute_metrics
    from hg_transformers.mask_trainer_Robust_VQA import Trainer
    from hg_transformers.optimization import get_constant_schedule
    from hg_transformers.training_args import TrainingArguments
    from optimization import AdamW
    from prune_debias_VQA import build_stage2
    os.environ["CRVQA_CUDA_GRAPH"] = "1" if graph else "0"
    os.environ["CRVQA_KEEP_GRADS"] = "1"     # this test reads the gradient arena AFTER the step (default: cleared)
    targs = TrainingArguments(output_dir=str(tmp_path), per_gpu_train_batch_size=256, logging_steps=1000, seed=49,
                              Masker_type="lpf", training_type="Masker", save_steps=0, dataloader_num_workers=0)
    model, masker, margs = build_stage2(3129, device=targs.device, seed=49,
                                        config_kwargs=dict(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0))
    model.classifier.main[2].p = 0.0
    params = [{"params": [p], "name": n} for n, p in model.named_parameters() if p.requires_grad]
    opt = AdamW(params, lr=5e-5, eps=1e-8)
    sched = get_constant_schedule(opt)
    trainer = Trainer(model=model, args=targs, model_args=margs, data_collator=TrimCollator(), train_dataset=None,
                      compute_metrics=vqa_compute_metrics, optimizers=(opt, sched), masker=masker)
    trainer._setup_engine(opt)
    trainer.global_step = 0
    trainer._zero_grad(opt)
    return trainer, model, masker, opt, sched


def _snapshot_state(arena, opt, model):
    arena._ensure_state()
    snap = {"scores": arena.scores.clone(), "thr": arena.thr_vec.clone(),
            "loose": [p.detach().clone() for p in model.parameters() if p.requires_grad and not arena.owns(p)]}
    return snap


def _restore_state(snap, arena, opt, model):
    arena.scores.copy_(snap["scores"])
    arena.set_thresholds(snap["thr"])
    for buf in (arena.exp_avg, arena.exp_avg_sq, arena.sum, arena.grads):
        buf.zero_()
    loose = [p for p in model.parameters() if p.requires_grad and not arena.owns(p)]
    for p, q in zip(loose, snap["loose"]):
        p.data.copy_(q)
        st = opt.state[p]
        for k in ("exp_avg", "exp_avg_sq", "sum"):
            if k in st:
                st[k].zero_()
    for grp in opt.param_groups:
        for p in grp["params"]:
            opt.state[p]["step"] = 0
    arena.refresh_masked()
    arena.begin_step()


@pytest.mark.parametrize("graph", [False, True])
def test_fast_path_config2_step_vs_reference_golden(tmp_path, graph):
    """B=256, A=3129, LPF, dropout off: ONE full training step of the benchmarked engine (eager, and replayed from the
    captured whole-step CUDA graph) against the reference's step: loss, logits (through the loss and the score), score
    gradients (statistics + samples), the clip norm, then AdamW + reset_threshold: kept counts and sampled scores."""
    from oracle import lxmert_oracle as lxo
    g = torch.load(os.path.join(GOLD, "config2_lxmert.pt"), weights_only=False)
    trainer, model, masker, opt, sched = _config2_trainer(tmp_path, graph)
    try:
        assert _sha(model) == g["state_sha"]                 # same seed-49 random init as the reference, bit for bit
        arena = trainer.arena
        mods = _mods(model)
        host = lxo.synthetic_batch(256, 3129)
        order = ["ids", "feats", "pos", "target", None, None, "bias", "max_label"]
        inputs = [host[k].cuda() if k else torch.arange(256).cuda() for k in order]
        snap = _snapshot_state(arena, opt, model)
        grads_seen = {}
        if graph:
            gs = trainer._make_graphed_step(model, opt, sched)
            assert gs is not None
            while gs.graph is None:                          # eager warm-up steps + capture (+ first replay)
                gs.step(inputs)
            _restore_state(snap, arena, opt, model)
            # gradients are consumed inside the graph (zero_grad protocol does not clear the arena buffer)
            loss, score = gs.step(inputs)
        else:
            _restore_state(snap, arena, opt, model)
            loss, score = trainer._device_step(model, inputs, opt)
        torch.cuda.synchronize()
        ref = float(g["loss_lpf"])
        lerr = abs(float(loss) - ref) / abs(ref)
        wl2, ws, pooled_s = _grad_report(mods, g["grad_stats_lpf"], g["nograd_lpf"])
        print(f"[config2 graph={graph}] loss {float(loss):.6f} vs {ref:.6f} ({lerr:.2e}) grad-L2 {wl2} sample {ws} "
              f"all-modules sample {pooled_s:.3e}")
        REPORT[f"config2 graph={graph}"] = dict(loss=lerr, grad_l2=wl2, grad_sample=ws, grad_sample_all=pooled_s)
        assert lerr < 1e-3
        assert float(score) == float(g["score"])
        assert wl2[0] < 3e-2, wl2
        assert ws[0] < 0.15, ws
        # after the optimiser step: thresholds + masks as the reference's reset_threshold / save_model_mask see them
        trainer.reset_threshold(model, 0.7)
        kept_diff = thr_rel = 0.0
        agree = total = 0
        for n, m in mods:
            kept = int((m.weight_mask.detach() > m.threshold).sum())
            kept_diff = max(kept_diff, abs(kept - g["kept_after"][n]) / max(1, g["kept_after"][n]))
            t_ref = float(g["thresholds_after"][n])
            thr_rel = max(thr_rel, abs(float(m.threshold) - t_ref) / max(abs(t_ref), 1e-12))
            s_ref = g["scores_after_sample"][n]
            s_got = _strided(m.weight_mask.detach()).cpu()
            m_ref, m_got = s_ref > t_ref, s_got > float(m.threshold)
            agree += int((m_ref == m_got).sum())
            total += m_ref.numel()
        print(f"[config2 graph={graph}] after 1 step: worst kept-count gap {kept_diff:.2e}, worst threshold gap "
              f"{thr_rel:.2e}, sampled mask agreement {agree}/{total}")
        REPORT[f"config2 graph={graph} after-step"] = dict(kept=kept_diff, thr=thr_rel, mask_agree=agree / total)
        assert kept_diff < 1e-3
        assert agree / total > 0.999
    finally:
        os.environ.pop("CRVQA_CUDA_GRAPH", None)
        os.environ.pop("CRVQA_KEEP_GRADS", None)


# ---------------------------------------------------------------------------------------------------------------
# mask agreement with the reference trajectory
# ---------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("graph", [False, True])
def test_mask_trajectory_iou_vs_reference(tmp_path, graph):
    """Six optimiser steps (lr 2e-3, LPF) of the engine against the reference loop's own trajectory: per-step losses,
    and after steps 1 and 6 the refreshed thresholds, kept counts and the masks themselves -- IoU per module as
    compare_mask.py:31-43, averaged.  Scores move by +-lr per step (Adam), so a mask bit differs only where the
    gradient SIGN of an element differs between the bf16 path and the fp32 reference at some step."""
    from hg_transformers.data.data_collator import TrimCollator
    from hg_transformers.data.metrics import vqa_compute_metrics
    from hg_transformers.mask_trainer_Robust_VQA import Trainer
    from hg_transformers.optimization import get_constant_schedule
    from hg_transformers.training_args import TrainingArguments
    from optimization import AdamW
    from oracle import lxmert_oracle as lxo
    from prune_debias_VQA import build_stage2
    t = torch.load(os.path.join(GOLD, "trajectory_small.pt"), weights_only=False)
    os.environ["CRVQA_CUDA_GRAPH"] = "1" if graph else "0"
    try:
        cfg = {k: v for k, v in t["config"].items() if k != "ans_num"}
        bt = t["batch"]
        targs = TrainingArguments(output_dir=str(tmp_path), per_gpu_train_batch_size=bt["B"], logging_steps=1000,
                                  seed=49, Masker_type="lpf", training_type="Masker", save_steps=0,
                                  dataloader_num_workers=0)
        model, masker, margs = build_stage2(bt["A"], device=targs.device, seed=49, config_kwargs=cfg)
        assert _sha(model) == t["state_sha"]
        model.classifier.main[2].p = 0.0
        params = [{"params": [p], "name": n} for n, p in model.named_parameters() if p.requires_grad]
        opt = AdamW(params, lr=t["lr"], eps=1e-8)
        sched = get_constant_schedule(opt)
        trainer = Trainer(model=model, args=targs, model_args=margs, data_collator=TrimCollator(), train_dataset=None,
                          compute_metrics=vqa_compute_metrics, optimizers=(opt, sched), masker=masker)
        trainer._setup_engine(opt)
        trainer.global_step = 0
        trainer._zero_grad(opt)
        assert model.lxmert.encoder._fast_plans() is not None
        gs = trainer._make_graphed_step(model, opt, sched) if graph else None
        if gs is not None:
            gs.warmup_steps = 0          # capture on the first call: the trajectory has no spare steps
        mods = _mods(model)
        assert [n for n, _ in mods] == t["module_names"]
        order = ["ids", "feats", "pos", "target", None, None, "bias", "max_label"]
        losses = []
        for step in range(1, t["steps"] + 1):
            host = lxo.synthetic_batch(bt["B"], bt["A"], seed=bt["seed"] + step, T=bt["T"], R=bt["Rg"],
                                       feat=bt["feat"], vocab=bt["vocab"])
            inputs = [host[k].cuda() if k else torch.arange(bt["B"]).cuda() for k in order]
            if gs is not None:
                loss, _ = gs.step(inputs)
            else:
                loss, _ = trainer._device_step(model, inputs, opt)
                sched.step()
            losses.append(float(loss))
            if step in t["snapshots"]:
                snap = t["snapshots"][step]
                mean_thr = trainer.reset_threshold(model, 0.7)
                ious, kept_gap, thr_gap = [], 0.0, 0.0
                for n, m in mods:
                    mask = (m.weight_mask.detach() > m.threshold).reshape(-1).cpu().numpy()
                    ref = np.unpackbits(snap["mask_bits"][n].numpy())[: mask.size].astype(bool)
                    ious.append((mask & ref).sum() / max(1, (mask | ref).sum()))
                    kept_gap = max(kept_gap, abs(int(mask.sum()) - snap["kept"][n]) / max(1, snap["kept"][n]))
                    tr = float(snap["thresholds"][n])
                    thr_gap = max(thr_gap, abs(float(m.threshold) - tr) / max(abs(tr), 1e-12))
                iou = float(np.mean(ious))
                print(f"[trajectory graph={graph}] step {step}: mask IoU mean {iou:.5f} min {min(ious):.5f}, worst "
                      f"kept-count gap {kept_gap:.2e}, worst threshold gap {thr_gap:.2e}, mean threshold "
                      f"{mean_thr:.6e} vs {snap['mean_threshold']:.6e}")
                REPORT[f"trajectory graph={graph} step {step}"] = dict(iou=iou, iou_min=float(min(ious)), kept=kept_gap,
                                                                       thr=thr_gap)
                assert iou > (0.999 if step == 1 else 0.98), iou
                assert kept_gap < 2e-3
        rel = [abs(a - b) / abs(b) for a, b in zip(losses, t["losses"])]
        print(f"[trajectory graph={graph}] losses {losses} vs reference {t['losses']}")
        REPORT[f"trajectory graph={graph} losses"] = rel
        assert rel[0] < 2e-3
        assert max(rel) < 0.1      # the last steps sit at loss ~0.2-0.6 where one flipped mask bit shows
    finally:
        os.environ.pop("CRVQA_CUDA_GRAPH", None)


def test_zz_write_parity_report():
    """Collect the measured gaps of this module into gpurun_out/parity_r2.json (copied to profiles/ by hand)."""
    import json
    out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out, exist_ok=True)
    with open(os.path.join(out, "parity_r2.json"), "w") as f:
        json.dump(REPORT, f, indent=1, default=str)
