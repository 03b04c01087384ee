"""Pins the CPU oracle (oracle/) against golden vectors produced by the UNMODIFIED reference
(tests/golden/make_golden.py).  CPU only; runs in the no-GPU tier."""
import json
import os

import pytest
import torch

from oracle import adamw as o_adamw
from oracle import losses as o_losses
from oracle import lxmert_oracle as lxo
from oracle import masked_ops as o_ops

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RATES = {"Lang": 1 - 0.3, "Vis": 1 - 0.3, "Fus": 1 - 0.3, "P": 0.7}


@pytest.fixture(scope="module")
def ops_gold():
    return torch.load(os.path.join(GOLD, "ops.pt"), weights_only=False)


@pytest.fixture(scope="module")
def tiny_gold():
    return torch.load(os.path.join(GOLD, "tiny_lxmert.pt"), weights_only=False)


def same_value(a, b):
    """float equality with -0 == +0 (CPU kthvalue leaves the sign of a zero unspecified)."""
    return float(a) == float(b)


def test_binarizer(ops_gold):
    out = o_ops.binarize(ops_gold["bin_in"], ops_gold["bin_thr"])
    assert torch.equal(out, ops_gold["bin_out"])
    assert out[0, 0] == 0.0  # strict '>' : a score equal to the threshold is masked out


def test_kth_value_matches_torch_kthvalue(ops_gold):
    for case in ops_gold["kth_cases"]:
        assert same_value(o_ops.kth_value(case["x"], case["k"]), case["v"])
        assert same_value(o_ops.kth_value(case["x"], case["k"], use_abs=True), case["v_abs"])


def test_magnitude_init(ops_gold):
    s, _ = o_ops.magnitude_init(ops_gold["ml_weight"], 0.7, 1e-2)
    assert torch.equal(s, ops_gold["ml_scores_init"])
    n = s.numel()
    assert int((s > 1e-2).sum()) == n - int(n * 0.7)  # no ties in a random weight: exactly n-k survive


def test_masked_linear_fwd_bwd(ops_gold):
    g = ops_gold
    x = g["ml_x"].clone().requires_grad_(True)
    s = g["ml_scores"].clone().requires_grad_(True)
    y = o_ops.masked_linear(x, s, g["ml_weight"], 1e-2, g["ml_bias"])
    y.backward(g["ml_dy"])
    torch.testing.assert_close(y, g["ml_y"], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(x.grad, g["ml_dx"], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(s.grad, g["ml_ds"], rtol=1e-6, atol=1e-8)
    # the literal reference graph on autograd gives the same thing as the hand-written backward
    x2 = g["ml_x"].clone().requires_grad_(True)
    s2 = g["ml_scores"].clone().requires_grad_(True)
    o_ops.masked_linear_reference_form(x2, s2, g["ml_weight"], 1e-2, g["ml_bias"]).backward(g["ml_dy"])
    torch.testing.assert_close(s2.grad, s.grad, rtol=1e-6, atol=1e-8)
    torch.testing.assert_close(x2.grad, x.grad, rtol=1e-6, atol=1e-7)


def test_masked_embedding(ops_gold):
    g = ops_gold
    s = g["emb_scores"].clone().requires_grad_(True)
    e = o_ops.masked_embedding(g["emb_ids"], s, g["emb_weight"], 1e-2, 0)
    e.backward(g["emb_dout"])
    assert torch.equal(e, g["emb_out"])
    torch.testing.assert_close(s.grad, g["emb_ds"], rtol=1e-6, atol=1e-9)
    assert float(s.grad[0].abs().sum()) == 0.0  # padding row gets no gradient


def test_losses(ops_gold):
    r = ops_gold["loss"]
    for kind in ("bce", "lpf", "lmh"):
        logits = r["logits"].clone().requires_grad_(True)
        pooled = r["pooled"].clone().requires_grad_(True)
        if kind == "bce":
            l = o_losses.bce_loss(logits, r["labels"])
        elif kind == "lpf":
            l = o_losses.lpf_loss(logits, r["bias"], r["max_label"], 5)
        else:
            l = o_losses.lmh_loss(pooled, logits, r["bias"], r["labels"], r["lin_w"], r["lin_b"], r["smooth_param"])
        l.backward()
        torch.testing.assert_close(l.detach(), r[kind], rtol=1e-6, atol=1e-6)
        torch.testing.assert_close(logits.grad, r[f"{kind}_dlogits"], rtol=1e-5, atol=1e-8)
        if kind == "lmh":
            torch.testing.assert_close(pooled.grad, r["lmh_dpooled"], rtol=1e-5, atol=1e-9)
    assert same_value(o_losses.vqa_score(r["logits"], r["labels"]), r["score"])


def test_clip_and_adamw(ops_gold):
    t = ops_gold["adamw"]
    ps = [p.clone() for p in t["p0"]]
    states = [o_adamw.new_state(p) for p in ps]
    for step in range(3):
        grads = [g.clone() for g in t["grads"][step]]
        coef, _ = o_adamw.clip_coef(grads, 1.0)
        for p, g, st in zip(ps, grads, states):
            o_adamw.adamw_step(p, g * coef, st, lr=5e-5)
        for p, ref, st, rs in zip(ps, t["p"][step], states, t["sum"][step]):
            torch.testing.assert_close(p, ref, rtol=1e-6, atol=1e-9)
            torch.testing.assert_close(st["sum"], rs, rtol=1e-6, atol=1e-9)


def test_module_census_and_modalities():
    with open(os.path.join(GOLD, "host.json")) as f:
        host = json.load(f)
    names = lxo.module_names()
    assert len(names) == 168
    ref_modal = host["chain_robust"]["modal"]
    assert {n for n, _ in names} <= set(host["chain_robust"]["names"])
    for n, m in names:
        assert ref_modal[n] == m


def _tiny_ctx(g, operand="fp32"):
    cfg = g["config"]
    params = {k: v.clone() for k, v in g["state_dict"].items()}
    for k in params:
        params[k].requires_grad_(k.startswith("classifier.") and k.split(".")[-1] in ("weight_g", "weight_v", "bias"))
    layers = (cfg["l_layers"], cfg["r_layers"], cfg["x_layers"])
    scores, thr, modal = lxo.init_scores(params, RATES, 1e-2, *layers)
    c = lxo.Ctx(params, scores, thr, heads=cfg["num_attention_heads"], operand=operand)
    return c, modal, layers


def test_tiny_lxmert_step_matches_reference(tiny_gold):
    g = tiny_gold
    c, modal, layers = _tiny_ctx(g)
    assert [n for n, _ in lxo.module_names(*layers)] == g["module_names"]
    assert modal == g["modal"]
    for n, s in c.S.items():
        assert int((s > 1e-2).sum()) == g["kept_init"][n]
    batch = g["batch"]
    lmh = {"lin_w": g["lmh_lin_w"], "lin_b": g["lmh_lin_b"], "smooth_param": g["lmh_smooth_param"]}
    for kind in ("normal", "lpf", "lmh"):
        opt_state = {} if kind == "lmh" else None
        out = lxo.training_step(c, batch, kind, opt_state=opt_state, lmh=lmh, layers=layers)
        torch.testing.assert_close(out["logits"], g["logits"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(out["pooled"], g["pooled"], rtol=1e-5, atol=1e-6)
        torch.testing.assert_close(out["loss"], g[f"loss_{kind}"], rtol=1e-5, atol=1e-6)
        names = list(c.S)
        for n, gr in zip(names, out["grads"]):
            ref = g[f"grads_{kind}"][n]
            torch.testing.assert_close(gr, ref, rtol=1e-4, atol=1e-7 * float(ref.abs().max() + 1e-30) + 1e-12)
        for n in g[f"nograd_{kind}"]:
            assert float(out["grads"][names.index(n)].abs().sum()) == 0.0
    assert same_value(out["score"], g["score"])
    torch.testing.assert_close(out["grad_norm"], g["grad_norm_lmh"], rtol=1e-5, atol=0)
    # after clip + AdamW: scores, then the per-modality thresholds of reset_threshold
    for n, s in c.S.items():
        torch.testing.assert_close(s.detach(), g["scores_after"][n], rtol=1e-6, atol=1e-9)
    thr, mean = lxo.reset_thresholds({n: g["scores_after"][n] for n in c.S}, modal, RATES)
    for n in c.S:
        assert same_value(thr[n], g["thresholds_after"][n]), n
        kept = int((g["scores_after"][n] > thr[n]).sum())
        assert kept == g["kept_after"][n]
    assert abs(mean - g["mean_threshold"]) <= 1e-9


def test_tiny_bf16_operand_mode_is_close_to_fp32(tiny_gold):
    """The 'bf16' operand mode only rounds GEMM operands; on the tiny model logits stay within 2e-2 of fp32."""
    g = tiny_gold
    c, _, layers = _tiny_ctx(g, operand="bf16")
    out = lxo.training_step(c, g["batch"], "normal", layers=layers)
    err = float((out["logits"] - g["logits"]).abs().max() / g["logits"].abs().max())
    assert err < 2e-2, err


def test_full_lxmert_config1_matches_reference_and_survey_known_answers():
    """BASELINE config 1 (9/5/5, h=768, B=32, A=2274, seed 49): the oracle against the reference's outputs
    (tests/golden/full_lxmert.pt) and the known-answer values of SURVEY.md section 8(c)."""
    from hg_transformers.modeling_lxmert import LxmertConfig, LxmertForMultipleChoice
    g = torch.load(os.path.join(GOLD, "full_lxmert.pt"), weights_only=False)
    torch.manual_seed(49)
    model = LxmertForMultipleChoice(LxmertConfig(ans_num=2274))
    params = {k: v.detach() for k, v in model.state_dict().items()}
    for k in params:
        params[k].requires_grad_(k.startswith("classifier."))
    scores, thr, modal = lxo.init_scores(params, RATES, 1e-2)
    assert list(scores) == g["module_names"] and modal == g["modal"]
    kept = {n: int((s > 1e-2).sum()) for n, s in scores.items()}
    assert kept == g["kept_init"]
    assert sum(kept.values()) == 62181835 and kept["lxmert.pooler.dense"] == 176948
    assert kept["lxmert.embeddings.word_embeddings"] == 7032269 and kept["lxmert.encoder.visn_fc.box_fc"] == 922
    batch = lxo.synthetic_batch(32, 2274)
    c = lxo.Ctx(params, scores, thr)
    out = lxo.training_step(c, batch, "normal")
    assert abs(float(out["loss"]) - 1577.634766) < 2e-3
    torch.testing.assert_close(out["loss"], g["loss_normal"], rtol=1e-6, atol=0)
    torch.testing.assert_close(out["logits"], g["logits"], rtol=1e-4, atol=2e-6)
    torch.testing.assert_close(out["pooled"], g["pooled"], rtol=1e-4, atol=2e-6)
    assert abs(float(out["logits"].sum()) - (-77.364212)) < 1e-2
    for n, gr in zip(scores, out["grads"]):
        st = g["grad_stats_normal"][n]
        assert abs(float(gr.double().norm()) - st["l2"]) <= 1e-4 * st["l2"] + 1e-12, n
        assert int((gr != 0).sum()) == st["nnz"], n
        samp = gr.reshape(-1)[:: max(1, gr.numel() // 512)][:512]
        torch.testing.assert_close(samp, st["sample"], rtol=1e-3, atol=1e-5 * float(st["sample"].abs().max()) + 1e-12)
    assert int((out["grads"][0] != 0).sum()) == 486912              # embedding dS nnz, SURVEY 8(c)
    lpf = lxo.compute_loss("lpf", out["logits"], out["pooled"], batch)
    assert abs(float(lpf) - 7.543721) < 1e-4


def test_global_threshold_variant_matches_reference(tiny_gold):
    """masking/global_maskers.py + global_mask_trainer_VQA.py on the tiny LXMERT (tests/golden/global_tiny.pt)."""
    g = torch.load(os.path.join(GOLD, "global_tiny.pt"), weights_only=False)
    sd = tiny_gold["state_dict"]
    names = g["module_names"]
    ws = [sd[n + ".weight"] for n in names]
    total = sum(w.numel() for w in ws)
    cut = o_ops.global_kth_value(ws, int(total * g["init_sparsity"]), use_abs=True)
    assert same_value(cut, g["global_weight_threshold"])
    scores = {n: o_ops.magnitude_init_global(w, cut, 1e-2) for n, w in zip(names, ws)}
    assert {n: int((s > 1e-2).sum()) for n, s in scores.items()} == g["kept_init"]
    gen = torch.Generator().manual_seed(g["noise_seed"])
    for n in names:
        scores[n] = scores[n] + torch.randn(scores[n].shape, generator=gen) * 5e-3
    for rate in (0.7, 0.35):
        thr = o_ops.global_kth_value(list(scores.values()), int(total * rate))
        # the reference returns float(torch.tensor([thr] * n_modules).mean()): fp32 mean of identical values
        assert float(torch.tensor([float(thr)] * len(names)).mean()) == g[f"union_threshold_{rate}"]
        assert {n: int((s > float(thr)).sum()) for n, s in scores.items()} == g[f"kept_after_{rate}"]
