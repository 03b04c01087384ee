"""Round-2 golden vectors, again outputs of the UNMODIFIED reference modules (/root/reference) on CPU fp32:

    python tests/golden/make_golden_r2.py [--skip-config2]

  config2_lxmert.pt    BASELINE config 2 (the benchmarked workload): LXMERT 9/5/5 h=768, batch 256, A=3129, LPF loss,
                       seed 49, dropout off -- loss, logits / pooled (first 32 rows + per-row statistics of all 256),
                       per-module score-gradient statistics and 512-element strided samples, the clip norm, then one
                       reference AdamW step + reset_threshold: thresholds, kept counts, a strided sample of the scores.
  trajectory_small.pt  a small LXMERT (h=256, 2L/1R/2X) trained for 6 steps exactly as the reference loop does
                       (forward, LPF_loss, backward, clip_grad_norm_, root optimization.AdamW, zero_grad) with
                       reset_threshold + mask export after step 1 and step 6: per-step losses, thresholds, full masks
                       (bit-packed) and kept counts.  lr = 2e-3 so that scores actually cross the threshold within
                       six steps (at the recipe's 5e-5 a mask needs ~200 steps to change at all).

The reference's loop: hg_transformers/mask_trainer_Robust_VQA.py:640-680 (clip, step, zero_grad), :801-886
(_training_step), :467-482 (reset_threshold), :930-949 (save_model_mask = S > thr).
"""
import argparse
import hashlib
import logging
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402

SMALL_CFG = dict(vocab_size=1000, hidden_size=256, num_attention_heads=4, intermediate_size=512, l_layers=2,
                 x_layers=2, r_layers=1, visual_feat_dim=128, visual_pos_dim=4, max_position_embeddings=32)
SMALL_BATCH = dict(B=16, A=96, seed=5, T=10, Rg=8, feat=128, vocab=1000)
TRAJ_STEPS, TRAJ_LR, TRAJ_SNAP = 6, 2e-3, (1, 6)


def state_sha(model):
    h = hashlib.sha256()
    for k, v in sorted(model.state_dict().items()):
        if "weight_mask" in k:
            continue
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def strided(t, n=512):
    f = t.reshape(-1)
    return f[:: max(1, f.numel() // n)][:n].clone()


def gen_config2(R):
    torch.manual_seed(49)
    cfg = R.cfg.LxmertConfig(ans_num=3129)
    model = R.lx.LxmertForMultipleChoice(cfg)
    out = {"state_sha": state_sha(model), "batch": {"B": 256, "A": 3129, "seed": 49}}
    batch = mg.synthetic_batch(256, 3129)
    masker = mg.make_masker(R, model)
    model.eval()
    mods = mg.masked_modules(model)
    model.zero_grad()
    _, logits, pooled = model(batch["ids"], batch["feats"], batch["pos"], labels=batch["target"])
    loss = R.trainer.LPF_loss(logits, batch["bias"], batch["max_label"], "cpu", 5)
    loss.backward()
    out["loss_lpf"] = loss.detach().clone()
    out["logits_head"] = logits.detach()[:32].clone()
    out["pooled_head"] = pooled.detach()[:32].clone()
    out["logits_row_l2"] = logits.detach().double().norm(dim=1).float()
    out["logits_absmax"] = logits.detach().abs().max().clone()
    out["logits_argmax"] = logits.detach().argmax(1)
    out["pooled_row_l2"] = pooled.detach().double().norm(dim=1).float()
    out["score"] = R.metrics.compute_score_with_logits("vqa", logits.detach(), batch["target"])["acc"].clone()
    stats = {}
    for n, m in mods:
        g = m.weight_mask.grad
        if g is None:
            stats[n] = {"l2": 0.0, "abs_mean": 0.0, "nnz": 0, "sample": torch.zeros(512)}
            continue
        g = g.detach()
        stats[n] = {"l2": float(g.double().norm()), "abs_mean": float(g.abs().mean()), "nnz": int((g != 0).sum()),
                    "sample": strided(g)}
    out["grad_stats_lpf"] = stats
    out["nograd_lpf"] = [n for n, m in mods if m.weight_mask.grad is None]
    out["cls_grads_lpf"] = {n: {"l2": float(p.grad.double().norm())} for n, p in model.named_parameters()
                            if n.startswith("classifier") and p.grad is not None}
    params = [p for _, p in model.named_parameters() if p.requires_grad]
    opt = R.optim.AdamW([{"params": [p]} for p in params], lr=5e-5, eps=1e-8)
    out["grad_norm_lpf"] = torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0).detach().clone()
    opt.step()
    dummy = types.SimpleNamespace(masker=masker)
    out["mean_threshold"] = R.trainer.Trainer.reset_threshold(dummy, model, 0.7)
    out["thresholds_after"] = {n: m.threshold.detach().clone() for n, m in mods}
    out["kept_after"] = {n: int((m.weight_mask.detach() > m.threshold).sum()) for n, m in mods}
    out["scores_after_sample"] = {n: strided(m.weight_mask.detach()) for n, m in mods}
    return out


def gen_trajectory(R):
    torch.manual_seed(49)
    b = SMALL_BATCH
    cfg = R.cfg.LxmertConfig(ans_num=b["A"], hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0, **SMALL_CFG)
    model = R.lx.LxmertForMultipleChoice(cfg)
    out = {"config": dict(SMALL_CFG, ans_num=b["A"], hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0),
           "batch": dict(b), "state_sha": state_sha(model), "lr": TRAJ_LR, "steps": TRAJ_STEPS}
    masker = mg.make_masker(R, model)
    model.eval()                     # dropout off everywhere (the classifier's 0.5 included)
    mods = mg.masked_modules(model)
    out["module_names"] = [n for n, _ in mods]
    params = [p for _, p in model.named_parameters() if p.requires_grad]
    opt = R.optim.AdamW([{"params": [p]} for p in params], lr=TRAJ_LR, eps=1e-8)
    dummy = types.SimpleNamespace(masker=masker)
    losses, norms, snaps = [], [], {}
    for step in range(1, TRAJ_STEPS + 1):
        batch = mg.synthetic_batch(b["B"], b["A"], seed=b["seed"] + step, T=b["T"], Rg=b["Rg"], feat=b["feat"],
                                   vocab=b["vocab"])
        _, logits, _ = model(batch["ids"], batch["feats"], batch["pos"], labels=batch["target"])
        loss = R.trainer.LPF_loss(logits, batch["bias"], batch["max_label"], "cpu", 5)
        loss.backward()
        norms.append(float(torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)))
        opt.step()
        model.zero_grad()
        losses.append(float(loss))
        if step in TRAJ_SNAP:
            mean_thr = R.trainer.Trainer.reset_threshold(dummy, model, 0.7)
            snap = {"mean_threshold": mean_thr, "thresholds": {}, "kept": {}, "mask_bits": {}, "score_sample": {}}
            for n, m in mods:
                mask = (m.weight_mask.detach() > m.threshold)
                snap["thresholds"][n] = m.threshold.detach().clone()
                snap["kept"][n] = int(mask.sum())
                snap["mask_bits"][n] = torch.from_numpy(np.packbits(mask.reshape(-1).numpy()))
                snap["score_sample"][n] = strided(m.weight_mask.detach())
            snaps[step] = snap
    out["losses"], out["grad_norms"], out["snapshots"] = losses, norms, snaps
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--skip-config2", action="store_true")
    args = ap.parse_args()
    logging.basicConfig(level=logging.WARNING)
    R = mg.load_reference()
    torch.save(gen_trajectory(R), os.path.join(HERE, "trajectory_small.pt"))
    if not args.skip_config2:
        torch.save(gen_config2(R), os.path.join(HERE, "config2_lxmert.pt"))
    print("round-2 golden vectors written to", HERE)


if __name__ == "__main__":
    main()
