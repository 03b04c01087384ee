"""Golden vectors for the global-threshold variant, from the UNMODIFIED reference
(masking/global_maskers.py, hg_transformers/global_mask_trainer_VQA.py) on the tiny LXMERT of make_golden.py:

    python tests/golden/make_golden_global.py        # writes tests/golden/global_tiny.pt
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


def main():
    R = mg.load_reference()
    G = importlib.import_module("masking.global_maskers")
    T = importlib.import_module("hg_transformers.global_mask_trainer_VQA")
    torch.manual_seed(49)
    cfg = R.cfg.LxmertConfig(vocab_size=200, hidden_size=64, ans_num=50, num_attention_heads=4, intermediate_size=128,
                             l_layers=2, x_layers=1, r_layers=1, visual_feat_dim=32, visual_pos_dim=4,
                             max_position_embeddings=16)
    model = R.lx.LxmertForMultipleChoice(cfg)
    out = {"state_dict_is": "tests/golden/tiny_lxmert.pt[state_dict] (same config, seed 49)"}
    conf = types.SimpleNamespace(
        masking_scheduler_conf_={"lambdas_lr": 0.0, "sparsity_warmup": "automated_gradual_sparsity",
                                 "sparsity_warmup_interval_epoch": 0.1, "init_epoch": 0.0, "final_epoch": 1.0,
                                 "final_sparsity": 0.7},
        logger=logging.getLogger("golden"), num_epochs=20)
    sched = R.sp.MaskerScheduler(conf)
    masker = G.Masker(masker_scheduler=sched, logger=logging.getLogger("golden"), mask_biases=False,
                      structured_masking_info={"structured_masking": None, "structured_masking_types": None,
                                               "force_masking": "bert"},
                      threshold=1e-2, init_scale=2e-2, which_ptl="lxmert", controlled_init="magnitude", global_prune=True)
    names = G.chain_module_names("lxmert", list(range(12)), mg.WEIGHT_TYPES)
    masker.patch_modules(model=model, names_tobe_masked=names, name_of_masker="MaskedLinear1")
    mods = mg.masked_modules(model)
    out["init_sparsity"] = float(sched.init_sparsity)
    out["module_names"] = [n for n, _ in mods]
    out["global_weight_threshold"] = masker.global_threshold.detach().clone()
    out["kept_init"] = {n: int((m.weight_mask.detach() > 1e-2).sum()) for n, m in mods}
    # scores after "training": add seeded noise, then the union reset_threshold of the global trainer
    g = torch.Generator().manual_seed(7)
    noise = {n: torch.randn(m.weight_mask.shape, generator=g) * 5e-3 for n, m in mods}
    for n, m in mods:
        m.weight_mask.data.add_(noise[n])
    out["noise_seed"] = 7
    dummy = types.SimpleNamespace(model_args=types.SimpleNamespace(global_prune=True))
    for rate in (0.7, 0.35):
        mean_thr = T.Trainer.reset_threshold(dummy, model, rate)
        out[f"union_threshold_{rate}"] = float(mean_thr)
        out[f"kept_after_{rate}"] = {n: int((m.weight_mask.detach() > m.threshold).sum()) for n, m in mods}
    torch.save(out, os.path.join(HERE, "global_tiny.pt"))
    print({k: v for k, v in out.items() if not isinstance(v, dict)})


if __name__ == "__main__":
    main()
