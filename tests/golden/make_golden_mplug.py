"""Golden vectors for the mPLUG masking path, from the UNMODIFIED reference modules
mPLUG/masking/{maskers,sparsity_control}.py run on the CPU over the miniature mPLUG-shaped network of
tests/mplug_skeleton.py:

    python tests/golden/make_golden_mplug.py        # writes tests/golden/mplug_skeleton.pt

Variants: (A) the shipped MaskConfigs (magnitude_soft; initial sparsity = zero rate); (B) init_sparsity 0 with the
automated-gradual ramp driving reset_threshold ("keep the old threshold" branch; B0: rank 0 -> threshold 0); (C) magnitude
init with one global cut and the global reset_threshold; (D) variant A's network cast to bf16 -- the DeepSpeed-bf16
state the reference trains in -- for the bf16 score / threshold comparisons.
"""
import contextlib
import copy
import importlib
import io
import logging
import os
import re
import sys
import types

import hashlib

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import mplug_skeleton as sk  # noqa: E402

REF = os.environ.get("CRVQA_REFERENCE_ROOT", "/root/reference")
SCHED = "lambdas_lr=0,sparsity_warmup=automated_gradual_sparsity,sparsity_warmup_interval_epoch=0.1,init_epoch=0,final_epoch=1"


def load_reference():
    pkg = types.ModuleType("ref_mplug_masking")
    pkg.__path__ = [os.path.join(REF, "mPLUG", "masking")]
    sys.modules["ref_mplug_masking"] = pkg
    return (importlib.import_module("ref_mplug_masking.maskers"),
            importlib.import_module("ref_mplug_masking.sparsity_control"))


def dict_parser(values):   # the `k=v,...` option strings of mPLUG/param_parser.py (floats where they parse)
    out = {}
    for kv in values.split(","):
        k, v = kv.split("=")
        try:
            out[k] = float(v)
        except ValueError:
            out[k] = v
    return out


def make_masker(M, SP, *, zero_rate, init_sparsity, final_epoch, controlled_init, global_prune):
    """The argument wiring of vqa_mplug.init_masker (:59-128), which itself cannot be imported (deepspeed)."""
    c = dict_parser(SCHED)
    c["final_sparsity"] = zero_rate
    c["final_epoch"] = final_epoch
    if init_sparsity is not None:
        c["init_sparsity"] = init_sparsity
    conf = types.SimpleNamespace(masking_scheduler_conf_=c, logger=logging.getLogger("golden"))
    sched = SP.MaskerScheduler(conf)
    masker = M.Masker(masker_scheduler=sched, logger=logging.getLogger("golden"), mask_biases=False,
                      structured_masking_info={"structured_masking": None, "structured_masking_types": None,
                                               "force_masking": "bert"},
                      threshold=1e-2, init_scale=2e-2, controlled_init=controlled_init, train_classifier=True,
                      global_prune=global_prune)
    return masker, sched


def masked(model):
    return [(n, m) for n, m in model.named_modules() if hasattr(m, "threshold")]


def pack(mask):
    return np.packbits((mask.detach().cpu().float() != 0).numpy().reshape(-1))


def digest(tensors):
    """SHA-256 over a name-sorted dict of tensors (fp32 bytes): the skeleton's weights and the perturbed scores are
    regenerated from their seeds in the tests and checked against this instead of being stored."""
    h = hashlib.sha256()
    for k in sorted(tensors):
        h.update(k.encode())
        h.update(tensors[k].detach().cpu().float().contiguous().numpy().tobytes())
    return h.hexdigest()


def thr_record(model):
    rec = {}
    for n, m in masked(model):
        t = m.threshold
        rec[n] = (float(t), str(t.dtype).replace("torch.", "") if torch.is_tensor(t) else type(t).__name__)
    return rec


def quiet(fn, *a, **k):
    buf = io.StringIO()
    with contextlib.redirect_stdout(buf):
        r = fn(*a, **k)
    return r, buf.getvalue()


def train_state(model, out):
    """loss + score gradients of one batch (fp32 reference arithmetic)."""
    model.train()
    for p in model.parameters():
        p.grad = None
    loss = model(*sk.batch())
    loss.backward()
    out["loss"] = float(loss.detach())
    out["grads"] = {n: (m.weight_mask.grad.clone() if m.weight_mask.grad is not None else None)
                    for n, m in masked(model)}
    out["head_grad_norm"] = float(model.text_decoder.cls.predictions.decoder.weight.grad.norm())


def variant(M, SP, *, zero_rate, init_sparsity, final_epoch, controlled_init, global_prune):
    model = sk.build()
    masker, sched = make_masker(M, SP, zero_rate=zero_rate, init_sparsity=init_sparsity, final_epoch=final_epoch,
                                controlled_init=controlled_init, global_prune=global_prune)
    names = sk.names_to_mask(M.chain_module_names)
    quiet(masker.patch_modules, model=model, names_tobe_masked=names, name_of_masker="MaskedLinear1")
    out = {"names_tobe_masked": sorted(names), "module_names": [n for n, _ in masked(model)],
           "trainable": sorted(n for n, p in model.named_parameters() if p.requires_grad),
           "init_sparsity": float(sched.init_sparsity), "init_thresholds": thr_record(model),
           "init_masks": {k: pack(v) for k, v in masker.init_masks.items()},
           "kept_init": {n: int(m.get_masks()[0].float().sum()) for n, m in masked(model)}}
    if global_prune:
        out["global_weight_threshold"] = float(masker.global_threshold)
    train_state(model, out)
    return model, masker, sched, out


def perturb(model, seed, scale):
    g = torch.Generator().manual_seed(seed)
    for _, m in masked(model):
        m.weight_mask.data.add_(torch.randn(m.weight_mask.shape, generator=g) * scale)


def reports(M, model, out, key):
    _, txt = quiet(M.see_sparsity, model)
    out[key + "_see_sparsity"] = float(re.search(r"Sparsity of entire model = ([0-9.]+)", txt).group(1))
    _, txt = quiet(M.save_model_mask, model, is_save=False)
    out[key + "_zero_rate"] = float(re.search(r"Zero rate of entire model = ([0-9.]+)", txt).group(1))
    out[key + "_masks"] = {n + ".weight": pack(m.get_masks()[0]) for n, m in masked(model)}
    out[key + "_kept"] = {n: int(m.get_masks()[0].float().sum()) for n, m in masked(model)}


def main():
    M, SP = load_reference()
    gold = {"skeleton_seed": 7, "batch_seed": 11, "state_dict_sha256": digest(sk.build().state_dict())}

    # (A) shipped configuration: magnitude_soft, init sparsity = zero rate
    model, masker, sched, A = variant(M, SP, zero_rate=0.7, init_sparsity=None, final_epoch=1,
                                      controlled_init="magnitude_soft", global_prune=False)
    reports(M, model, A, "start")
    perturb(model, 3, 2e-3)
    A["perturb"] = (3, 2e-3)
    A["resets"] = []
    for rate in (0.7, 0.35, 0.9):
        mean = M.reset_threshold(model, rate)
        A["resets"].append({"rate": rate, "mean": mean, "thresholds": thr_record(model),
                            "kept": {n: int(m.get_masks()[0].float().sum()) for n, m in masked(model)}})
    reports(M, model, A, "after")
    before = thr_record(model)
    M.reset_threshold(model, 1e-4)                  # rank int(n * 1e-4) == 0 for every module: nothing moves
    A["tiny_rate_moves_nothing"] = before == thr_record(model)
    train_state(model, A_after := {})
    A_after["grad_norms"] = {n: (float(g.norm()) if g is not None else None) for n, g in A_after.pop("grads").items()}
    A["after_train"] = A_after
    gold["A"] = A

    # (D) the same network in the DeepSpeed-bf16 state: parameters (weights AND scores) bf16, thresholds as they are
    model16 = copy.deepcopy(model)
    D = {"fp32_scores_sha256": digest({n: m.weight_mask for n, m in masked(model)}),
         "fp32_thresholds": thr_record(model)}
    model16.bfloat16()
    D["kept_before"] = {n: int(m.get_masks()[0].float().sum()) for n, m in masked(model16)}
    D["resets"] = []
    for rate in (0.5, 0.8):
        mean = M.reset_threshold(model16, rate)
        D["resets"].append({"rate": rate, "mean": mean, "thresholds": thr_record(model16),
                            "masks": {n: pack(m.get_masks()[0]) for n, m in masked(model16)}})
    gold["D"] = D

    # (B0) init_sparsity 0: rank 0 -> every threshold is the Python int 0 and every non-zero weight is kept.  (The
    # reference cannot refresh thresholds from this state at target 0: torch.tensor([0, 0, ...]).mean() raises.)
    model, masker, sched, B0 = variant(M, SP, zero_rate=0.7, init_sparsity=0.0, final_epoch=4,
                                       controlled_init="magnitude_soft", global_prune=False)
    try:
        M.reset_threshold(model, 0.0)
        B0["reset_at_zero"] = "ok"
    except RuntimeError as e:
        B0["reset_at_zero"] = "RuntimeError: " + str(e)[:60]
    B0.pop("grads")
    gold["B0"] = B0

    # (B) ramp 0.1 -> 0.7 over 4 epochs: scheduler.step drives the refresh
    model, masker, sched, B = variant(M, SP, zero_rate=0.7, init_sparsity=0.1, final_epoch=4,
                                      controlled_init="magnitude_soft", global_prune=False)
    B.pop("grads")
    B["ramp"] = []
    for epoch in range(6):
        _, target, changed = masker.masker_scheduler.step(cur_epoch=epoch)
        mean = M.reset_threshold(model, target)
        B["ramp"].append({"epoch": epoch, "target": target, "changed": changed, "mean": mean,
                          "thresholds": thr_record(model),
                          "kept": {n: int(m.get_masks()[0].float().sum()) for n, m in masked(model)}})
    # all scores of one module equal: the k-th value is not below the maximum, the old threshold stays
    name0, mod0 = masked(model)[3]
    mod0.weight_mask.data.fill_(0.25)
    before = float(mod0.threshold)
    M.reset_threshold(model, 0.5)
    B["constant_scores"] = {"module": name0, "threshold_before": before, "threshold_after": float(mod0.threshold)}
    gold["B"] = B

    # (C) magnitude init with ONE global |W| cut, then the global threshold refresh
    model, masker, sched, C = variant(M, SP, zero_rate=0.6, init_sparsity=0.5, final_epoch=1,
                                      controlled_init="magnitude", global_prune=True)
    C["grad_norms"] = {n: (float(g.norm()) if g is not None else None) for n, g in C.pop("grads").items()}
    perturb(model, 5, 4e-3)
    C["perturb"] = (5, 4e-3)
    C["global_resets"] = []
    for rate in (0.6, 0.2):
        mean = M.reset_threshold(model, rate, global_prune=True)
        C["global_resets"].append({"rate": rate, "mean": mean, "thresholds": thr_record(model),
                                   "kept": {n: int(m.get_masks()[0].float().sum()) for n, m in masked(model)}})
    gold["C"] = C

    # (T) a short training trajectory: AdamW on scores + LM head, mask update (scheduler.step -> reset_threshold) every
    # two steps, bf16-free (fp32 scores): losses, thresholds and kept counts along the way
    model, masker, sched, T = variant(M, SP, zero_rate=0.7, init_sparsity=0.3, final_epoch=2,
                                      controlled_init="magnitude_soft", global_prune=False)
    T.pop("grads")
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=2e-3, weight_decay=0.0)
    T["lr"], T["steps"] = 2e-3, []
    data = sk.batch()
    model.train()
    for step in range(6):
        loss = model(*data)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.requires_grad and p.grad is not None], 1.0)
        opt.step()
        rec = {"loss": float(loss.detach())}
        if (step + 1) % 2 == 0:
            _, target, _ = masker.masker_scheduler.step(cur_epoch=(step + 1) // 2)
            rec["target"] = target
            rec["mean"] = M.reset_threshold(model, target)
            rec["thresholds"] = thr_record(model)
            rec["kept"] = {n: int(m.get_masks()[0].float().sum()) for n, m in masked(model)}
        T["steps"].append(rec)
    gold["T"] = T

    # host-only vectors: chain_module_names per tower
    gold["chain"] = {t: sorted(M.chain_module_names(t, list(range(3)), ab)) for t, ab in {
        "visual_encoder": ["AO_visual", "I_visual", "O_visual", "AO", "I", "O", "E"],
        "text_encoder": ["K", "Q", "V", "AO", "I", "O", "E"],
        "fusion_encoder": ["SK", "SQ", "SV", "SAO", "CK", "CQ", "CV", "CAO", "I", "O", "E"],
        "text_decoder": ["SK", "SQ", "SV", "SAO", "CK", "CQ", "CV", "CAO", "I", "O", "E"]}.items()}
    path = os.path.join(HERE, "mplug_skeleton.pt")
    torch.save(gold, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")
    print("A loss", A["loss"], "see", A["start_see_sparsity"], "zero", A["start_zero_rate"], "resets",
          [(r["rate"], r["mean"]) for r in A["resets"]])
    print("B0", B0["reset_at_zero"], "B ramp", [(r["epoch"], round(r["target"], 4), r["mean"]) for r in B["ramp"]], B["constant_scores"])
    print("C", C["global_weight_threshold"], [(r["rate"], r["mean"]) for r in C["global_resets"]])
    print("D", [(r["rate"], r["mean"]) for r in D["resets"]])
    print("T", [round(r["loss"], 4) for r in T["steps"]], [r.get("target") for r in T["steps"]])


if __name__ == "__main__":
    main()
