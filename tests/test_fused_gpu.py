"""Fused layer kernels (dropout + residual + LayerNorm, GELU) and the engine fast path vs plain torch math."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("H,gdt", [(768, torch.bfloat16), (768, torch.float32), (256, torch.bfloat16), (1024, torch.float32)])
def test_drop_add_layernorm_no_dropout(H, gdt):
    from crvqa import fused
    torch.manual_seed(0)
    M = 1000
    g = torch.randn(M, H, device="cuda").to(gdt).requires_grad_(True)
    res = torch.randn(M, H, device="cuda", requires_grad=True)
    ln = torch.nn.LayerNorm(H, eps=1e-12).cuda()
    ln.weight.data.uniform_(0.5, 1.5)
    ln.bias.data.uniform_(-0.5, 0.5)
    ln.weight.requires_grad_(False)
    ln.bias.requires_grad_(False)
    y32, y16 = fused.drop_add_layernorm(g, res, ln, 0.1, 7, training=False)
    ref = F.layer_norm(g.float() + res, (H,), ln.weight, ln.bias, 1e-12)
    torch.testing.assert_close(y32, ref, rtol=1e-5, atol=1e-5)
    assert torch.equal(y16, y32.bfloat16())
    d32 = torch.randn(M, H, device="cuda")
    d16 = torch.randn(M, H, device="cuda").bfloat16()
    torch.autograd.backward([y32, y16], [d32, d16])
    g_ref = g.detach().float().requires_grad_(True)
    r_ref = res.detach().clone().requires_grad_(True)
    F.layer_norm(g_ref + r_ref, (H,), ln.weight, ln.bias, 1e-12).backward(d32 + d16.float())
    torch.testing.assert_close(res.grad, r_ref.grad, rtol=1e-4, atol=1e-4)
    assert g.grad.dtype == gdt
    torch.testing.assert_close(g.grad.float(), g_ref.grad.bfloat16().float(), rtol=2e-2, atol=2e-2)


def test_drop_add_layernorm_dropout_statistics_and_backward_consistency():
    from crvqa import fused
    M, H, p = 2048, 768, 0.1
    g = torch.ones(M, H, device="cuda", requires_grad=True)
    ln = torch.nn.LayerNorm(H, eps=1e-12).cuda()
    ln.weight.requires_grad_(False)
    ln.bias.requires_grad_(False)
    rng = fused.RngState.get(g.device)
    rng.advance()
    # z = dropout(1): entries are 0 or 1/(1-p); recover the mask from y (normalised z keeps the two levels apart)
    y32, _ = fused.drop_add_layernorm(g, None, ln, p, 11, training=True)
    keep = y32 > 0
    rate = 1.0 - float(keep.float().mean())
    assert abs(rate - p) < 5e-3, rate
    y_again, _ = fused.drop_add_layernorm(g, None, ln, p, 11, training=True)
    assert torch.equal(y_again, y32)                       # same (seed, counter, site) -> same mask
    y_site, _ = fused.drop_add_layernorm(g, None, ln, p, 12, training=True)
    assert not torch.equal(y_site > 0, keep)               # another call site -> another mask
    rng.advance()
    y_next, _ = fused.drop_add_layernorm(g, None, ln, p, 11, training=True)
    assert not torch.equal(y_next > 0, keep)               # next step -> another mask
    # backward regenerates the forward mask: dropped positions get zero gradient
    rng2 = fused.RngState.get(g.device)
    x = torch.randn(M, H, device="cuda", requires_grad=True)
    y, _ = fused.drop_add_layernorm(x, None, ln, p, 21, training=True)
    z_keep = None
    y.backward(torch.randn_like(y))
    # positions whose gradient is exactly zero are the dropped ones; their share must be ~p
    zero_share = float((x.grad == 0).float().mean())
    assert abs(zero_share - p) < 5e-3, zero_share


def test_gelu_bf16():
    from crvqa import fused
    u = (torch.randn(4096, 3072, device="cuda") * 2).bfloat16().requires_grad_(True)
    y = fused.gelu_bf16(u)
    ref = F.gelu(u.detach().float())
    # the kernel's erf is Abramowitz-Stegun 7.1.26 (|erf error| < 1.5e-7) on SFU rcp / ex2, so the fp32 value is within
    # 5e-7 ABSOLUTE of torch's.  Tolerance: that plus one bf16 rounding step (2^-8 relative); in the lower tail
    # (gelu(-3) = -0.004) 5e-7 absolute is a few 1e-5 relative, enough to land on the neighbouring bf16 value there.
    diff = (y.float() - ref).abs()
    assert bool((diff <= ref.abs() * 2.0 ** -8 + 1e-6).all())
    assert float((y.float() != ref.bfloat16().float()).float().mean()) < 0.05
    dy = torch.randn_like(y)
    y.backward(dy)
    u_ref = u.detach().float().requires_grad_(True)
    F.gelu(u_ref).backward(dy.float())
    torch.testing.assert_close(u.grad.float(), u_ref.grad.bfloat16().float(), rtol=1e-2, atol=1e-3)


def test_engine_fast_path_matches_generic_path():
    """Same model, same scores: fused fast path (arena + mask cache) vs generic per-module path, dropout off."""
    import os
    from crvqa import ops
    from hg_transformers._engine import ScoreArena, masked_modules_of
    from oracle import lxmert_oracle as lxo
    from prune_debias_VQA import build_stage2
    cfg = dict(vocab_size=1000, hidden_size=256, num_attention_heads=4, intermediate_size=512, l_layers=2,
               x_layers=2, r_layers=1, visual_feat_dim=128, max_position_embeddings=32)
    model, masker, _ = build_stage2(96, device=torch.device("cuda"), seed=5, config_kwargs=cfg)
    model.eval()
    batch = {k: v.cuda() for k, v in lxo.synthetic_batch(16, 96, seed=5, T=10, R=8, feat=128, vocab=1000).items()}

    def run(zero=True):
        if zero:
            model.zero_grad()
        _, logits, _ = model(batch["ids"], batch["feats"], batch["pos"], labels=batch["target"])
        loss, _ = ops.vqa_loss_bce(logits, batch["target"])
        loss.backward()
        return logits.detach().clone(), float(loss.detach())

    lg0, l0 = run()
    mods = masked_modules_of(model)
    plain = {n: (m.weight_mask.grad.clone() if m.weight_mask.grad is not None else None) for n, m in mods}
    arena = ScoreArena(mods)
    arena.enable_mask_cache()
    assert model.lxmert.encoder._fast_plans() is not None
    arena.begin_step()
    lg1, l1 = run(zero=False)
    arena.finalize_grads()
    err = float((lg1 - lg0).abs().max() / lg0.abs().max())
    assert err < 2e-2, err                                  # bf16 intermediates (QKV, GELU, AO outputs) vs fp32 ones
    assert abs(l1 - l0) <= 1e-3 * abs(l0)
    worst = 0.0
    for n, m in mods:
        if plain[n] is None:
            assert float(m.weight_mask.grad.abs().max()) == 0.0
            continue
        rel = float((m.weight_mask.grad - plain[n]).double().norm() / (plain[n].double().norm() + 1e-30))
        worst = max(worst, rel)
    assert worst < 8e-2, worst
    os.environ["CRVQA_FUSED"] = "0"
    try:
        assert model.lxmert.encoder._fast_plans() is None
    finally:
        os.environ.pop("CRVQA_FUSED")


def _ref_attention(q, k, v, heads, mask):
    B, Sq, H = q.shape
    d = H // heads
    qh = q.float().view(B, Sq, heads, d).transpose(1, 2)
    kh = k.float().view(B, -1, heads, d).transpose(1, 2)
    vh = v.float().view(B, -1, heads, d).transpose(1, 2)
    s = qh @ kh.transpose(-1, -2) / d ** 0.5
    if mask is not None:
        s = s + mask[:, None, None, :]
    return (torch.softmax(s, -1) @ vh).transpose(1, 2).reshape(B, Sq, H)


@pytest.mark.parametrize("Sq,Sk,kind,use_mask", [(20, 20, 0, False), (36, 36, 0, True), (20, 36, 1, False),
                                                 (36, 20, 1, True), (56, 56, 0, False), (7, 64, 2, True)])
def test_small_attention_vs_torch(Sq, Sk, kind, use_mask):
    from crvqa import fused
    torch.manual_seed(Sq * 100 + Sk)
    B, heads, H = 5, 12, 768
    mask = None
    if use_mask:
        mask = torch.zeros(B, Sk, device="cuda")
        mask[:, Sk - 3:] = -10000.0
    if kind == 0:
        qkv = (torch.randn(B, Sq, 3 * H, device="cuda") * 0.7).bfloat16().requires_grad_(True)
        srcs = (qkv,)
        q, k, v = qkv.detach().split(H, -1)
    elif kind == 1:
        qt = (torch.randn(B, Sq, H, device="cuda") * 0.7).bfloat16().requires_grad_(True)
        kv = (torch.randn(B, Sk, 2 * H, device="cuda") * 0.7).bfloat16().requires_grad_(True)
        srcs = (qt, kv)
        q, (k, v) = qt.detach(), kv.detach().split(H, -1)
    else:
        srcs = tuple((torch.randn(B, s, H, device="cuda") * 0.7).bfloat16().requires_grad_(True) for s in (Sq, Sk, Sk))
        q, k, v = (t.detach() for t in srcs)
    out = fused.small_attention(kind, heads, mask, 0.1, 3, False, *srcs)
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    ref = _ref_attention(qr, kr, vr, heads, mask)
    err = float((out.float() - ref).abs().max() / ref.abs().max())
    assert err < 2e-2, err                                      # bf16 probabilities / bf16 output
    do = torch.randn_like(out)
    out.backward(do)
    ref.backward(do.float())
    if kind == 0:
        got = srcs[0].grad.float().split(H, -1)
    elif kind == 1:
        got = (srcs[0].grad.float(),) + tuple(srcs[1].grad.float().split(H, -1))
    else:
        got = tuple(t.grad.float() for t in srcs)
    for g, r in zip(got, (qr.grad, kr.grad, vr.grad)):
        rel = float((g - r).norm() / r.norm())
        assert rel < 2e-2, rel


def test_small_attention_dropout_is_unbiased_and_replayable():
    from crvqa import fused
    B, S, heads, H = 64, 36, 12, 768
    qkv = (torch.randn(B, S, 3 * H, device="cuda") * 0.5).bfloat16()
    base = fused.small_attention(0, heads, None, 0.1, 5, False, qkv).float()
    rng = fused.RngState.get(qkv.device)
    rng.advance()
    a = fused.small_attention(0, heads, None, 0.1, 5, True, qkv).float()
    b = fused.small_attention(0, heads, None, 0.1, 5, True, qkv).float()
    assert torch.equal(a, b)                                    # same step, same site -> same mask
    assert not torch.equal(a, base)
    acc = torch.zeros_like(base)
    n = 24
    for _ in range(n):
        rng.advance()
        acc += fused.small_attention(0, heads, None, 0.1, 5, True, qkv).float()
    rel = float((acc / n - base).norm() / base.norm())
    assert rel < 0.1, rel                                       # E[dropout(P)] = P


@pytest.mark.parametrize("Sq,Sk", [(36, 36), (20, 36), (36, 20), (56, 56)])
def test_small_attention_dropout_forward_and_backward_use_one_mask(Sq, Sk):
    """Training-mode parity: the dropout mask depends only on (rng state, site, batch, head, i, j), so it can be
    read back by running the forward with one-hot V rows; torch attention with THAT mask is then the reference
    for the forward and for dQ / dK / dV (pass A and pass B of the backward regenerate the same mask)."""
    from crvqa import fused
    torch.manual_seed(Sq + Sk)
    B, heads, H, p = 4, 12, 768, 0.1
    d = H // heads
    srcs = tuple((torch.randn(B, s, H, device="cuda") * 0.7).bfloat16().requires_grad_(True) for s in (Sq, Sk, Sk))
    q, k, v = (t.detach() for t in srcs)
    rng = fused.RngState.get(q.device)
    rng.advance()
    onehot = torch.zeros(B, Sk, heads, d, device="cuda")
    onehot[:, torch.arange(Sk), :, torch.arange(Sk)] = 1.0          # V[b, j, h, :] = e_j  (Sk <= 64 = d)
    probe = fused.small_attention(2, heads, None, p, 9, True, q, k, onehot.view(B, Sk, H).bfloat16())
    keep = (probe.view(B, Sq, heads, d)[..., :Sk] != 0).permute(0, 2, 1, 3).float()   # [B, heads, Sq, Sk]
    share = 1.0 - float(keep.mean())
    assert abs(share - p) < 0.02, share
    out = fused.small_attention(2, heads, None, p, 9, True, *srcs)
    qr, kr, vr = (t.float().requires_grad_(True) for t in (q, k, v))
    qh = qr.view(B, Sq, heads, d).transpose(1, 2)
    kh = kr.view(B, Sk, heads, d).transpose(1, 2)
    vh = vr.view(B, Sk, heads, d).transpose(1, 2)
    pd = torch.softmax(qh @ kh.transpose(-1, -2) / d ** 0.5, -1) * keep / (1.0 - p)
    ref = (pd @ vh).transpose(1, 2).reshape(B, Sq, H)
    assert float((out.float() - ref).abs().max() / ref.abs().max()) < 2e-2
    do = torch.randn_like(out)
    out.backward(do)
    ref.backward(do.float())
    for t, r in zip(srcs, (qr, kr, vr)):
        rel = float((t.grad.float() - r.grad).norm() / r.grad.norm())
        assert rel < 2e-2, rel


@pytest.mark.parametrize("Sq,Sk,p", [(36, 36, 0.1), (20, 36, 0.1), (36, 20, 0.0), (20, 20, 0.1), (56, 56, 0.1), (7, 64, 0.1),
                                     (33, 17, 0.1)])
def test_saved_probability_backward_matches_recomputing_backward(Sq, Sk, p, monkeypatch):
    """The training path keeps the signed softmax probabilities of the forward (crv_attention_fwd_p) and its backward
    (crv_attention_bwd_p) never recomputes them; the recomputing backward (crv_attention_bwd) regenerates the same
    dropout decisions from the hash.  Same inputs, same step, same site: identical forward output, and gradients that
    differ only by the bf16 rounding of the stored probabilities (1e-2 norm-wise; measured ~3e-3)."""
    from crvqa import fused
    torch.manual_seed(Sq * 64 + Sk)
    B, heads, H = 6, 12, 768
    mask = torch.zeros(B, Sk, device="cuda")
    mask[:, Sk - 2:] = -10000.0
    base = [(torch.randn(B, s, H, device="cuda") * 0.7).bfloat16() for s in (Sq, Sk, Sk)]
    do = torch.randn(B, Sq, H, device="cuda").bfloat16()
    rng = fused.RngState.get(do.device)
    rng.advance()
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("CRVQA_ATTN_SAVE_P", mode)
        srcs = tuple(t.clone().requires_grad_(True) for t in base)
        out = fused.small_attention(2, heads, mask, p, 11, True, *srcs)
        out.backward(do)
        res[mode] = (out.detach().clone(), [t.grad.float() for t in srcs])
    assert torch.equal(res["1"][0], res["0"][0])
    for a, b in zip(res["1"][1], res["0"][1]):
        assert bool(torch.isfinite(a).all())
        assert float((a - b).norm() / b.norm()) < 1e-2


@pytest.mark.parametrize("dual", [True, False])
@pytest.mark.parametrize("M,H", [(9216, 768), (5120, 768), (37, 256), (515, 1024)])
def test_entry_block_layernorm_average_dropout(dual, M, H):
    """crv_ln_avg_drop_fwd / _bwd against the PyTorch ops of LxmertVisualFeatureEncoder / LxmertEmbeddings
    (hg_transformers/modeling_lxmert.py:576-592, 744-770): without dropout value and gradients to 1e-5 of the output /
    gradient scale; with dropout the kept pattern is read from the output, the torch graph with THAT mask is the
    reference, and the backward must have regenerated the same pattern."""
    import types
    from crvqa import fused
    torch.manual_seed(M + H + dual)
    mk = lambda: types.SimpleNamespace(weight=(torch.rand(H, device="cuda") + 0.5), bias=torch.randn(H, device="cuda") * 0.1,
                                       eps=1e-12, normalized_shape=(H,))
    ln_a, ln_b = mk(), (mk() if dual else None)
    a = torch.randn(M, H, device="cuda", requires_grad=True)
    b = (torch.randn(M, H, device="cuda") * 2 + 0.3).requires_grad_(True) if dual else None
    assert fused.ln_avg_drop_usable(a, b, ln_a, ln_b)

    def ref(a_, b_, keep, scale):
        y = F.layer_norm(a_, (H,), ln_a.weight, ln_a.bias, 1e-12)
        if dual:
            y = (y + F.layer_norm(b_, (H,), ln_b.weight, ln_b.bias, 1e-12)) / 2
        return y * keep * scale

    fused.RngState.get(a.device).advance()
    for p in (0.0, 0.1):
        y32 = fused.ln_avg_drop(a, b, ln_a, ln_b, p, 77, True)
        y16 = y32._crv_bf16
        keep = torch.ones_like(y32)
        if p > 0:
            dense = ref(a.detach(), b.detach() if dual else None, 1.0, 1.0)
            keep = ((y32.detach() != 0) | (dense == 0)).float()
            assert abs(1.0 - float(keep.mean()) - p) < 0.01
        a2 = a.detach().clone().requires_grad_(True)
        b2 = b.detach().clone().requires_grad_(True) if dual else None
        want = ref(a2, b2, keep, 1.0 / (1.0 - p))
        scale = float(want.abs().max())
        assert float((y32.detach() - want.detach()).abs().max()) <= 1e-5 * scale
        assert torch.equal(y16.detach(), y32.detach().bfloat16())
        d32, d16 = torch.randn_like(y32), torch.randn(M, H, device="cuda").bfloat16()
        grads = torch.autograd.grad((y32, y16), (a, b) if dual else (a,), (d32, d16))
        want_g = torch.autograd.grad(want, (a2, b2) if dual else (a2,), d32 + d16.float())
        for g, w in zip(grads, want_g):
            assert float((g - w).abs().max()) <= 2e-5 * float(w.abs().max()), float((g - w).abs().max() / w.abs().max())
    # only one branch needs a gradient (box_fc's input never does, but its weight scores do; the embeddings' ids do not)
    if dual:
        y32 = fused.ln_avg_drop(a.detach(), b, ln_a, ln_b, 0.0, 77, True)
        (gb,) = torch.autograd.grad(y32.sum(), (b,))
        assert bool(torch.isfinite(gb).all())


def test_visualbert_fast_path_matches_generic_path():
    """VisualBERT (BASELINE config 3 in miniature, 20 + 36 = 56 tokens): fused fast path vs generic per-module path on
    the same scores, dropout off -- logits, loss and every score gradient."""
    import logging
    import os
    import types
    from crvqa import ops
    from hg_transformers._engine import ScoreArena, masked_modules_of
    from hg_transformers.modeling_visualbert import VisualBertForMultipleChoice, visualBERTConfig
    from masking import maskers_visualBert as mk
    from masking import sparsity_control as spc
    torch.manual_seed(7)
    cfg = visualBERTConfig(vocab_size=1000, hidden_size=256, num_hidden_layers=3, num_attention_heads=4,
                           intermediate_size=512, visual_embedding_dim=128, ans_num=96, max_position_embeddings=64)
    model = VisualBertForMultipleChoice(cfg).cuda()
    conf = types.SimpleNamespace(masking_scheduler_conf_={"final_sparsity": 0.7, "sparsity_warmup_interval_epoch": 0.1,
                                                          "lambdas_lr": 0.0, "init_epoch": 0, "final_epoch": 1},
                                 logger=logging.getLogger("vb"), num_epochs=1)
    log = logging.getLogger("vb"); log.setLevel(logging.ERROR)
    masker = mk.Masker(masker_scheduler=spc.MaskerScheduler(conf), logger=log, mask_biases=False,
                       structured_masking_info={"structured_masking": None, "structured_masking_types": None,
                                                "force_masking": "bert"},
                       threshold=1e-2, init_scale=2e-2, which_ptl="visual_bert", controlled_init="magnitude")
    masker.patch_modules(model, mk.chain_module_names("visual_bert", list(range(12)), ["K", "Q", "V", "AO", "I", "O", "P", "E"]),
                         "MaskedLinear1")
    model.eval()
    B, T, R = 16, 20, 36
    ids = torch.randint(1, 1000, (B, T), device="cuda")
    feats = torch.randn(B, R, 128, device="cuda")
    target = (torch.rand(B, 96, device="cuda") > 0.97).float()

    def run(zero=True):
        if zero:
            model.zero_grad()
        out = model(input_ids=ids, visual_embeds=feats, labels=target)
        logits = out[1]
        loss, _ = ops.vqa_loss_bce(logits, target)
        loss.backward()
        return logits.detach().clone(), float(loss.detach())

    lg0, l0 = run()
    mods = masked_modules_of(model)
    plain = {n: (m.weight_mask.grad.clone() if m.weight_mask.grad is not None else None) for n, m in mods}
    arena = ScoreArena(mods)
    arena.enable_mask_cache()
    assert model.visual_bert.encoder._fast_plans() is not None
    arena.begin_step()
    lg1, l1 = run(zero=False)
    arena.finalize_grads()
    assert float((lg1 - lg0).abs().max() / lg0.abs().max()) < 2e-2      # bf16 intermediates vs fp32 ones
    assert abs(l1 - l0) <= 1e-3 * abs(l0)
    for n, m in mods:
        if plain[n] is None:
            continue
        rel = float((m.weight_mask.grad - plain[n]).double().norm() / (plain[n].double().norm() + 1e-30))
        assert rel < 8e-2, (n, rel)
    os.environ["CRVQA_FUSED"] = "0"
    try:
        assert model.visual_bert.encoder._fast_plans() is None
    finally:
        os.environ.pop("CRVQA_FUSED")


def test_quick_gelu_bf16_matches_torch():
    """crv_quick_gelu_fwd / _bwd (mPLUG's CLIP tower, mPLUG/models/clip/model.py:25-27) against torch on the same bf16
    input: bf16 rounding of the result only."""
    from crvqa import fused
    torch.manual_seed(9)
    u = (torch.randn(1000, 3072, device="cuda") * 2).bfloat16().requires_grad_(True)
    y = fused.quick_gelu_bf16(u)
    uf = u.detach().float().requires_grad_(True)
    ref = uf * torch.sigmoid(1.702 * uf)
    assert float((y.float() - ref).abs().max()) <= 2 ** -8 * float(ref.abs().max())
    dy = torch.randn_like(y)
    y.backward(dy)
    ref.backward(dy.float())
    assert float((u.grad.float() - uf.grad).abs().max()) <= 2 ** -7 * float(uf.grad.abs().max())


def test_layernorm_bf16_matches_torch():
    """fused.layernorm_bf16 (the LayerNorms of mPLUG's CLIP tower on bf16 activations, mPLUG/models/clip/model.py:12-19:
    fp32 statistics, result in the input's dtype) against torch's fp32 LayerNorm of the same bf16 input: the result
    differs by its bf16 rounding only, the input gradient by bf16 rounding of the output gradient; rows that do not
    fill the last CTA, CLIP's LayerNorm module takes the path by itself and falls back when it may not."""
    from crvqa import fused
    from mPLUG.models.clip.model import LayerNorm
    torch.manual_seed(4)
    M, H = 4 * 577 + 3, 768
    ln = LayerNorm(H).cuda()
    with torch.no_grad():
        ln.weight.copy_(1 + 0.1 * torch.randn(H))
        ln.bias.copy_(0.1 * torch.randn(H))
    x = (torch.randn(M, H, device="cuda") * 1.5 + 0.3).bfloat16().requires_grad_(True)
    assert not fused.layernorm_bf16_usable(x, ln)            # trainable gamma / beta: the PyTorch ops
    with torch.autocast("cuda", dtype=torch.bfloat16):       # as the tower runs: torch normalises in fp32, casts back
        assert ln(x).dtype == torch.bfloat16
    ln.weight.requires_grad = ln.bias.requires_grad = False
    assert fused.layernorm_bf16_usable(x, ln)
    y3 = ln(x.view(M, 1, H))                                 # through the module, any leading shape
    assert y3.dtype == torch.bfloat16 and y3.shape == (M, 1, H)
    assert "LayerNormBf16" in type(y3.grad_fn).__name__, type(y3.grad_fn).__name__
    y = y3.view(M, H)
    xf = x.detach().float().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xf, (H,), ln.weight, ln.bias, ln.eps)
    err = float((y.detach().float() - ref.detach()).abs().max()) / float(ref.detach().abs().max())
    assert err <= 2 ** -8, err
    dy = torch.randn(M, H, device="cuda").bfloat16()
    y.backward(dy)
    ref.backward(dy.float())
    gerr = float((x.grad.float() - xf.grad).abs().max()) / float(xf.grad.abs().max())
    assert gerr <= 2 ** -7, gerr
    assert not fused.layernorm_bf16_usable(x.float(), ln)
    assert not fused.layernorm_bf16_usable(x[:, :700], torch.nn.LayerNorm(700).cuda().requires_grad_(False))


def _fq_reference(q, k, v, mask, H, keep, p):
    B, Lq, D = q.shape
    Lk = k.shape[1]
    qf, kf, vf = (t.detach().float().requires_grad_(True) for t in (q, k, v))
    s = torch.einsum("bihd,bjhd->bhij", qf.view(B, Lq, H, 64), kf.view(B, Lk, H, 64)) / 8.0
    if mask is not None:
        s = s + mask.float()
    P = torch.softmax(s, dim=-1)
    out = torch.einsum("bhij,bjhd->bihd", P * keep / (1.0 - p), vf.view(B, Lk, H, 64)).reshape(B, Lq, D)
    return out, P, (qf, kf, vf)


@pytest.mark.parametrize("B,H,Lq,Lk,mask_kind,p", [
    (3, 12, 6, 593, "row", 0.0), (3, 12, 6, 593, "row", 0.1), (2, 12, 16, 577, "row", 0.1), (4, 12, 6, 6, "causal", 0.1),
    (2, 12, 16, 16, "row", 0.0), (2, 4, 1, 37, None, 0.0), (2, 12, 9, 130, "full", 0.25), (1, 2, 16, 1024, "row1", 0.1)])
def test_few_query_attention_matches_reference_math(B, H, Lq, Lk, mask_kind, p):
    """crv_fq_attention_fwd / _bwd (mPLUG's text-side attention cores, mPLUG/models/modeling_mplug.py:205-300) against
    softmax(Q K^T / 8 + mask) V in fp32 on the same bf16 projections, with the dropout decisions the kernel took (the
    sign of its saved probabilities): saved |P| and the output differ by bf16 rounding, dQ / dK / dV by the bf16
    rounding of P and of the results (2^-6 of the largest entry); about p of the probabilities are dropped."""
    from crvqa import fused
    g = torch.Generator(device="cuda").manual_seed(100 + Lq + Lk)
    D = H * 64
    q = torch.randn(B, Lq, D, device="cuda", generator=g).bfloat16().requires_grad_(True)
    k = torch.randn(B, Lk, D, device="cuda", generator=g).bfloat16().requires_grad_(True)
    v = torch.randn(B, Lk, D, device="cuda", generator=g).bfloat16().requires_grad_(True)
    mask = None
    if mask_kind == "row":
        mask = torch.zeros(B, 1, 1, Lk, device="cuda")
        mask[B - 1, :, :, Lk - max(1, Lk // 5):] = -10000.0
    elif mask_kind == "row1":
        mask = torch.zeros(1, 1, 1, Lk, device="cuda")
        mask[..., ::7] = -10000.0
    elif mask_kind == "causal":
        mask = torch.zeros(B, 1, Lq, Lk, device="cuda").masked_fill_(
            torch.ones(Lq, Lk, device="cuda").triu(1).bool(), -10000.0)
    elif mask_kind == "full":
        mask = (torch.rand(B, 1, Lq, Lk, device="cuda", generator=g) < 0.2).float() * -10000.0
        mask[..., 0] = 0.0
    assert fused.few_query_attention_usable(q, k, v, mask, H)
    site = fused.RngState.new_site()
    out = fused.few_query_attention(q, k, v, mask, H, p, site, training=True)
    assert out.shape == (B, Lq, D) and out.dtype == torch.bfloat16
    probs = out.grad_fn.saved_tensors[3]
    assert probs.shape == (B, H, Lq, Lk)
    keep = (probs.view(torch.int16) >= 0).float()
    ref, P, leaves = _fq_reference(q, k, v, mask, H, keep, p)
    assert float((probs.float().abs() - P).abs().max()) <= 2 ** -8 * float(P.max()) + 1e-6
    if p == 0.0:
        assert bool(keep.all())
    elif keep.numel() > 20000:
        assert abs(1.0 - float(keep.mean()) - p) < 0.02
        again = fused.few_query_attention(q, k, v, mask, H, p, site, training=True)      # same counter: same masks
        assert torch.equal(again.grad_fn.saved_tensors[3], probs)
        fused.RngState.get(q.device).advance()
        other = fused.few_query_attention(q, k, v, mask, H, p, site, training=True)
        assert not torch.equal(other.grad_fn.saved_tensors[3], probs)
    assert float((out.float() - ref).abs().max()) <= 2 ** -6 * float(ref.abs().max())
    dout = torch.randn(B, Lq, D, device="cuda", generator=g).bfloat16()
    out.backward(dout)
    ref.backward(dout.float())
    for name, got, want in zip("qkv", (q.grad, k.grad, v.grad), (t.grad for t in leaves)):
        err = float((got.float() - want).abs().max()) / float(want.abs().max())
        assert err <= 2 ** -6, (name, err)
    # evaluation mode: no dropout whatever p says
    ev = fused.few_query_attention(q, k, v, mask, H, p, site, training=False)
    assert bool((ev.grad_fn.saved_tensors[3].view(torch.int16) >= 0).all())
    assert not fused.few_query_attention_usable(q[:, :, :64], k, v, mask, H)
    assert not fused.few_query_attention_usable(q.float(), k, v, mask, H)


def test_bert_self_attention_takes_the_few_query_kernel(monkeypatch):
    """mPLUG's BertSelfAttention on bf16 projections (the engine's bf16-activation mode): the cross attention of 6 answer
    tokens to 593 image / question tokens through crv_fq_attention equals the scaled_dot_product_attention path
    (CRVQA_MPLUG_FUSED=0) within bf16 rounding, output and input gradients; 593 queries keep the library path."""
    from types import SimpleNamespace
    from mPLUG.models.modeling_mplug import BertSelfAttention
    torch.manual_seed(2)
    cfg = SimpleNamespace(hidden_size=768, num_attention_heads=12, encoder_width=768, attention_probs_dropout_prob=0.1)
    att = BertSelfAttention(cfg, is_cross_attention=True).cuda().eval()
    hid = torch.randn(4, 6, 768, device="cuda", requires_grad=True)
    enc = torch.randn(4, 593, 768, device="cuda", requires_grad=True)
    mask = torch.zeros(4, 1, 1, 593, device="cuda")
    mask[2, ..., 580:] = -10000.0
    dy = torch.randn(4, 6, 768, device="cuda")

    def run(h, e):
        with torch.autocast("cuda", dtype=torch.bfloat16):
            y = att(h, None, e, mask)
        gh, ge = torch.autograd.grad(y, (h, e), dy.to(y.dtype))
        return y, gh, ge

    y1, gh1, ge1 = run(hid, enc)
    assert "FewQueryAttention" in type(y1.grad_fn).__name__
    monkeypatch.setenv("CRVQA_MPLUG_FUSED", "0")
    y0, gh0, ge0 = run(hid, enc)
    assert "FewQueryAttention" not in type(y0.grad_fn).__name__
    for a, b in ((y1, y0), (gh1, gh0), (ge1, ge0)):
        assert float((a.float() - b.float()).abs().max()) <= 2 ** -5 * float(b.float().abs().max())
    monkeypatch.setenv("CRVQA_MPLUG_FUSED", "1")
    with torch.autocast("cuda", dtype=torch.bfloat16):
        big = att(enc, mask, None, None)
    assert "FewQueryAttention" not in type(big.grad_fn).__name__ and big.shape == (4, 593, 768)
