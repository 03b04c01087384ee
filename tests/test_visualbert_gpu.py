"""VisualBERT stage-2 path on the GPU against the reference's outputs (tests/golden/visualbert_tiny.pt): state-dict
compatibility, bit-exact initial masks and thresholds, logits / loss / score gradients of the generic per-module path
and of the fused fast path (bf16 operands: tolerances written at the asserts)."""
import logging
import os
import types

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _batch(cfg, B=8, T=20, R=36, seed=49):
    g = torch.Generator().manual_seed(seed)
    ids = torch.randint(1, cfg["vocab_size"], (B, T), generator=g)
    feats = torch.randn(B, R, cfg["visual_embedding_dim"], generator=g)
    target = (torch.rand(B, cfg["ans_num"], generator=g) > 0.9).float() * torch.rand(B, cfg["ans_num"], generator=g)
    return ids.cuda(), feats.cuda(), target.cuda()


def _build(g):
    from hg_transformers.modeling_visualbert import VisualBertForMultipleChoice, visualBERTConfig
    from masking import maskers_visualBert as mk
    from masking import sparsity_control as spc
    model = VisualBertForMultipleChoice(visualBERTConfig(**g["config"]))
    model.load_state_dict(g["state_dict"], strict=True)          # same parameter / buffer names as the reference
    model.cuda()
    log = logging.getLogger("vbg"); log.setLevel(logging.ERROR)
    conf = types.SimpleNamespace(masking_scheduler_conf_={"final_sparsity": 0.7, "sparsity_warmup_interval_epoch": 0.1,
                                                          "lambdas_lr": 0.0, "init_epoch": 0, "final_epoch": 1},
                                 logger=log, num_epochs=20)
    masker = mk.Masker(masker_scheduler=spc.MaskerScheduler(conf), logger=log, mask_biases=False,
                       structured_masking_info={"structured_masking": None, "structured_masking_types": None,
                                                "force_masking": "bert"},
                       threshold=1e-2, init_scale=2e-2, which_ptl="visual_bert", controlled_init="magnitude")
    masker.patch_modules(model, mk.chain_module_names("visual_bert", list(range(12)), ["K", "Q", "V", "AO", "I", "O", "P", "E"]),
                         "MaskedLinear1")
    return model, masker


def _check(g, mods, logits, loss, grads, tol_logit, tol_norm, tol_elem):
    """Against the fp32 reference.  Logits / loss: 2e-3-class agreement scaled by depth.  Score gradients: the L2
    norm per module within tol_norm; element-wise they sit on the bf16-operand noise floor (a
    1e-6 input perturbation already moves bf16-operand gradients by 2 % inside the CPU oracle; query / key gradients,
    which pass through the softmax, are the most sensitive), hence the looser tol_elem."""
    scale = float(g["logits"].abs().max())
    assert float((logits.cpu() - g["logits"]).abs().max()) < tol_logit * scale
    assert float(loss) == pytest.approx(float(g["loss"]), rel=2e-3)
    for n, _ in mods:
        st = g["grad_stats"][n]
        gr = grads[n].cpu()
        flat = gr.reshape(-1)
        samp = flat[:: max(1, flat.numel() // 2048)][:2048]
        rel = float((samp - st["sample"]).norm() / (st["sample"].norm() + 1e-30))
        assert rel < tol_elem, (n, rel)
        assert abs(float(gr.double().norm()) - st["l2"]) <= tol_norm * st["l2"], n


def test_visualbert_against_reference_generic_and_fast_path():
    from hg_transformers._engine import ScoreArena, masked_modules_of
    g = torch.load(os.path.join(GOLD, "visualbert_tiny.pt"), weights_only=False)
    model, masker = _build(g)
    mods = masked_modules_of(model)
    assert [n for n, _ in mods] == g["module_names"]
    assert {n: int((m.weight_mask.detach() > 1e-2).sum()) for n, m in mods} == g["kept_init"]      # bit-exact masks
    assert sorted(n for n, p in model.named_parameters() if p.requires_grad) == g["trainable"]
    ids, feats, target = _batch(g["config"])
    model.eval()

    def run():
        loss, logits, _ = model(input_ids=ids, visual_embeds=feats, labels=target)[:3]
        loss.backward()
        return logits.detach(), float(loss.detach())

    # generic per-module path (MaskedLinear1 -> masked GEMM with the in-kernel mask transform), 2 layers: bf16 operands
    model.zero_grad()
    logits, loss = run()
    _check(g, mods, logits, loss, {n: m.weight_mask.grad for n, m in mods}, 1e-2, 3e-2, 8e-2)
    # fused fast path (arena, mask cache, grouped QKV, small-sequence attention, fused LN / GELU, bf16 activations)
    arena = ScoreArena(mods)
    arena.enable_mask_cache()
    assert model.visual_bert.encoder._fast_plans() is not None
    arena.begin_step()
    logits, loss = run()
    arena.finalize_grads()
    _check(g, mods, logits, loss, {n: m.weight_mask.grad for n, m in mods}, 2e-2, 6e-2, 1.2e-1)


def test_visualbert_reset_threshold_bit_exact():
    from hg_transformers import mask_trainer_visualBERT_VQA as vt
    from hg_transformers._engine import masked_modules_of
    g = torch.load(os.path.join(GOLD, "visualbert_tiny.pt"), weights_only=False)
    model, masker = _build(g)
    mods = masked_modules_of(model)
    gen = torch.Generator().manual_seed(7)
    for n, m in mods:
        m.weight_mask.data.add_((torch.randn(m.weight_mask.shape, generator=gen) * 5e-3).cuda())
    tr = vt.Trainer.__new__(vt.Trainer)
    tr.masker = masker
    mean_thr = tr.reset_threshold(model, 0.7)
    assert mean_thr == pytest.approx(g["mean_threshold"], rel=1e-6)
    for n, m in mods:
        assert float(m.threshold) == float(g["thresholds_after"][n]), n
        assert int((m.weight_mask.detach() > m.threshold).sum()) == g["kept_after"][n], n


def test_visualbert_full_size_fast_path_against_reference():
    """BASELINE config 3 at FULL size (12 layers, h = 768, 2048-d regions, A = 3129, 20 + 36 tokens; batch 32, eval
    mode) against tests/golden/visualbert_full.pt, written by the unmodified reference
    (hg_transformers/modeling_visualbert.py:1037-1174, mask_trainer_visualBERT_VQA.py:815-830): the seed-49 init is
    reproduced bit for bit (state-dict SHA-256), the 74 magnitude-initialised masks are bit-exact, and the FUSED fast
    path (arena, mask cache, grouped 2-CTA GEMMs, fused LayerNorm / GELU / attention, bf16 activations) agrees with the
    fp32 reference within the tolerances written below (bf16 operands through 12 layers; measured values in the
    asserts' messages).  Thresholds after a seeded score perturbation: the exact order statistics, bit for bit."""
    import hashlib
    from hg_transformers import mask_trainer_visualBERT_VQA as vt
    from hg_transformers._engine import ScoreArena, masked_modules_of
    from hg_transformers.modeling_visualbert import VisualBertForMultipleChoice, visualBERTConfig
    from masking import maskers_visualBert as mk
    from masking import sparsity_control as spc
    g = torch.load(os.path.join(GOLD, "visualbert_full.pt"), weights_only=False)
    torch.manual_seed(49)
    model = VisualBertForMultipleChoice(visualBERTConfig(**g["config"]))
    h = hashlib.sha256()
    for k, v in sorted(model.state_dict().items()):
        h.update(k.encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    assert h.hexdigest() == g["state_sha"]
    model.cuda()
    log = logging.getLogger("vbg"); log.setLevel(logging.ERROR)
    conf = types.SimpleNamespace(masking_scheduler_conf_={"final_sparsity": 0.7, "sparsity_warmup_interval_epoch": 0.1,
                                                          "lambdas_lr": 0.0, "init_epoch": 0, "final_epoch": 1},
                                 logger=log, num_epochs=20)
    masker = mk.Masker(masker_scheduler=spc.MaskerScheduler(conf), logger=log, mask_biases=False,
                       structured_masking_info={"structured_masking": None, "structured_masking_types": None,
                                                "force_masking": "bert"},
                       threshold=1e-2, init_scale=2e-2, which_ptl="visual_bert", controlled_init="magnitude")
    masker.patch_modules(model, mk.chain_module_names("visual_bert", list(range(12)),
                                                      ["K", "Q", "V", "AO", "I", "O", "P", "E"]), "MaskedLinear1")
    mods = masked_modules_of(model)
    assert [n for n, _ in mods] == g["module_names"] and len(mods) == 74
    assert {n: int((m.weight_mask.detach() > 1e-2).sum()) for n, m in mods} == g["kept_init"]
    assert sorted(n for n, p in model.named_parameters() if p.requires_grad) == g["trainable"]
    ids, feats, target = _batch(dict(vocab_size=g["vocab_size"], visual_embedding_dim=2048, ans_num=3129), B=g["B"])
    model.eval()
    arena = ScoreArena(mods)
    arena.enable_mask_cache()
    assert model.visual_bert.encoder._fast_plans() is not None
    arena.begin_step()
    loss, logits, pooled = model(input_ids=ids, visual_embeds=feats, labels=target)[:3]
    loss.backward()
    arena.finalize_grads()
    scale = float(g["logits"].abs().max())
    gap = float((logits.detach().cpu() - g["logits"]).abs().max()) / scale
    assert gap < 1e-2, gap                                   # LXMERT's 19 layers measure 5e-3
    assert float(loss.detach()) == pytest.approx(float(g["loss"]), rel=2e-3)
    worst_l2, worst_s = 0.0, 0.0
    for n, m in mods:
        st = g["grad_stats"][n]
        gr = m.weight_mask.grad.detach()
        l2 = float(gr.double().norm())
        worst_l2 = max(worst_l2, abs(l2 - st["l2"]) / st["l2"])
        flat = gr.reshape(-1)
        samp = flat[:: max(1, flat.numel() // 512)][:512].cpu()
        worst_s = max(worst_s, float((samp - st["sample"]).norm() / (st["sample"].norm() + 1e-30)))
    assert worst_l2 < 3e-2, worst_l2                         # per-module gradient L2 norms
    assert worst_s < 1.5e-1, worst_s                         # 512-element samples: the bf16-operand noise floor
    # thresholds: seeded perturbation of the scores, then the visualBERT trainer's reset_threshold -- bit-exact
    gen = torch.Generator().manual_seed(7)
    for n, m in mods:
        m.weight_mask.data.add_((torch.randn(m.weight_mask.shape, generator=gen) * 5e-3).cuda())
    tr = vt.Trainer.__new__(vt.Trainer)
    tr.masker = masker
    mean_thr = tr.reset_threshold(model, 0.7)
    assert mean_thr == pytest.approx(g["mean_threshold"], rel=1e-6)
    for n, m in mods:
        assert float(m.threshold) == g["thresholds_after"][n], n
        assert int((m.weight_mask.detach() > m.threshold).sum()) == g["kept_after"][n], n
