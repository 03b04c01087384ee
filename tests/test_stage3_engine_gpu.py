"""Stage-3 engine (hg_transformers._engine_ft.WeightArena + run_vqa_stage3.ArenaAdam + the fused layer path):
its new kernels against torch math, and the whole engine step against (a) the reference's outputs
(tests/golden/stage3_full.pt) and (b) the per-tensor path of masking/pruned.py + torch.optim.Adam.
Tolerances are written at each assert; bit-exact: pruned positions never move, masks never change."""
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD, "stage3_full.pt"), weights_only=False)


@pytest.mark.parametrize("M,N", [(5120, 768), (9216, 3072), (256, 2304), (100, 8), (7, 1536)])
def test_colsum_bf16(M, N):
    from crvqa import ops
    torch.manual_seed(M + N)
    x = torch.randn(M, N, device="cuda").bfloat16()
    out = torch.full((N,), 3.0, device="cuda")
    ops.colsum_bf16(x, out)
    ref = x.double().sum(0)
    assert float((out.double() - ref).abs().max()) <= 1e-5 * float(x.double().abs().sum(0).max())
    first = out.clone()
    ops.colsum_bf16(x, out)
    assert torch.equal(out, first)                      # deterministic
    ops.colsum_bf16(x, out, accumulate=True)
    torch.testing.assert_close(out, 2 * first, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("H,M,p", [(768, 1000, 0.0), (768, 5120, 0.1), (256, 37, 0.0), (1024, 515, 0.0)])
def test_layernorm_backward_parameter_gradients(H, M, p):
    """dgamma / dbeta of the fused dropout + residual + LayerNorm against torch autograd on the same pre-norm input."""
    from crvqa import fused
    torch.manual_seed(H + M)
    g = torch.randn(M, H, device="cuda").bfloat16().requires_grad_(True)
    res = torch.randn(M, H, device="cuda", requires_grad=True)
    ln = torch.nn.LayerNorm(H, eps=1e-12).cuda()
    ln.weight.data.uniform_(0.5, 1.5)
    ln.bias.data.uniform_(-0.5, 0.5)
    fused.RngState.get(g.device).advance()
    y32, y16 = fused.drop_add_layernorm(g, res, ln, p, 31, training=True)
    d32 = torch.randn(M, H, device="cuda")
    d16 = torch.randn(M, H, device="cuda").bfloat16()
    torch.autograd.backward([y32, y16], [d32, d16])
    # recover z = dropout(g) + res from the output: xhat = (y - beta) / gamma
    xhat = (y32.detach() - ln.bias.detach()) / ln.weight.detach()
    d = (d32 + d16.float())
    dgamma, dbeta = (d * xhat).double().sum(0), d.double().sum(0)
    tol = 1e-4 * float((d * xhat).abs().double().sum(0).max())
    assert float((ln.weight.grad.double() - dgamma).abs().max()) < tol
    assert float((ln.bias.grad.double() - dbeta).abs().max()) < 1e-5 * float(d.abs().double().sum(0).max())
    if p == 0.0:
        w2, b2 = ln.weight.detach().clone().requires_grad_(True), ln.bias.detach().clone().requires_grad_(True)
        g2, r2 = g.detach().float().requires_grad_(True), res.detach().clone().requires_grad_(True)
        F.layer_norm(g2 + r2, (H,), w2, b2, 1e-12).backward(d)
        torch.testing.assert_close(ln.weight.grad, w2.grad, rtol=1e-4, atol=1e-4 * float(w2.grad.abs().max()))
        torch.testing.assert_close(ln.bias.grad, b2.grad, rtol=1e-4, atol=1e-4 * float(b2.grad.abs().max()))
        torch.testing.assert_close(res.grad, r2.grad, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("wd", [0.0, 0.01])
def test_adam_torch_mode_matches_torch_optim_adam(wd):
    """crv_adamw_step / crv_adamw_segmented in mode 1 against torch.optim.Adam over several steps, including the
    clip coefficient and the masked bf16 operand refresh."""
    from crvqa import ops
    torch.manual_seed(5)
    n = 3 * 8192 + 4096 + 64
    p0 = torch.randn(n, device="cuda")
    mask = (torch.rand(n, device="cuda") > 0.7).float()
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
    # flat launch
    pf, mf, vf = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    # segmented launch with operand refresh over the first 3 * 8192 elements
    ps, ms, vs = p0.clone(), torch.zeros(n, device="cuda"), torch.zeros(n, device="cuda")
    gs = torch.zeros(n, device="cuda")
    w16, wm = mask.bfloat16(), torch.zeros(n, dtype=torch.bfloat16, device="cuda")
    rows = [((c0) // 8, min(8192, n - c0), 0, 1 if c0 < 3 * 8192 else 0) for c0 in range(0, n, 8192)]
    chunks = torch.tensor(rows, dtype=torch.int32, device="cuda")
    for t in range(1, 6):
        g = torch.randn(n, device="cuda") * mask * (3.0 if t == 2 else 0.01)
        sumsq = (g.double() ** 2).sum().float().reshape(())
        clip = min(1.0, 1.0 / (float(sumsq.sqrt()) + 1e-6))
        ref.grad = g * clip
        opt.step()
        ops.adamw_step_flat(pf, g, mf, vf, None, 1e-3, t, 0.9, 0.999, 1e-8, wd, sumsq, 1.0, mode=ops.ADAM_TORCH)
        gs.copy_(g)
        ops.adamw_segmented(ps, gs, ms, vs, None, chunks, None, w16, wm, 1e-3, t, 0.9, 0.999, 1e-8, wd, sumsq, 1.0,
                            zero_grad=True, mode=ops.ADAM_TORCH)
        assert float(gs.abs().max()) == 0.0
        for got in (pf, ps):
            torch.testing.assert_close(got, ref.detach(), rtol=2e-6, atol=2e-7)
        assert torch.equal(wm[: 3 * 8192], (ps * mask).bfloat16()[: 3 * 8192])
    st = opt.state[ref]
    torch.testing.assert_close(ms, st["exp_avg"], rtol=1e-5, atol=1e-9)
    torch.testing.assert_close(vs, st["exp_avg_sq"], rtol=1e-5, atol=1e-12)
    if wd == 0.0:
        assert torch.equal(ps[mask == 0], p0[mask == 0])          # zero gradient, zero decay: pruned entries never move


def _stage3_model(gold, dropout=None, A=None):
    import run_vqa_stage3 as s3
    from crvqa import ops
    from hg_transformers.modeling_lxmert import LxmertConfig, LxmertForMultipleChoice
    torch.manual_seed(49)
    model = LxmertForMultipleChoice(LxmertConfig(ans_num=A or gold["A"])).cuda()
    bert = model.lxmert
    mods = dict(bert.named_modules())
    names = s3.trained_mask_module_names()
    ws = [mods[n].weight.detach() for n in names]
    ks = [max(1, int(w.numel() * gold["zero_rate"])) for w in ws]
    thr = ops.kth_value_batched(ws, ks, use_abs=True)
    mask = {f"lxmert.{n}.weight_mask": (w.abs() > thr[i]) for i, (n, w) in enumerate(zip(names, ws))}
    s3.pruning_model_with_mask(bert, mask, "lxmert")
    if dropout is not None:
        for m in model.modules():
            if isinstance(m, torch.nn.Dropout):
                m.p = dropout
    return model, mask, s3


def _trainer(model, s3, B, loss="lmh", lr=5e-5):
    from hg_transformers.data.data_collator import TrimCollator
    from hg_transformers.mask_trainer_VQA import Trainer
    from hg_transformers.training_args import TrainingArguments
    from prune_debias_VQA import ModelArguments
    targs = TrainingArguments(output_dir="/tmp/crvqa_stage3", per_gpu_train_batch_size=B, logging_steps=1000, seed=49,
                              training_type="FT_trainedMask", FT_type=loss, save_steps=0, dataloader_num_workers=0,
                              learning_rate=lr)
    opt, sch = s3.init_optimizer(model, targs, B * 100)
    tr = Trainer(model=model, args=targs, model_args=ModelArguments(), data_collator=TrimCollator(),
                 optimizers=(opt, sch), masker=None)
    return tr, opt, sch


def _inputs(B, A, seed=49):
    from oracle import lxmert_oracle as lxo
    host = lxo.synthetic_batch(B, A, seed=seed)
    return [host[k].cuda() if k else torch.arange(B) for k in ["ids", "feats", "pos", "target", None, None, "bias",
                                                                "max_label"]]


def test_engine_forward_backward_against_reference(gold, monkeypatch):
    """The golden test of test_stage3_gpu.py through the ENGINE: arena views, fused layer path (grouped 2-CTA GEMMs
    where the shapes allow, fused LayerNorm with dgamma / dbeta, column-sum bias gradients), gradients read from G."""
    monkeypatch.setenv("CRVQA_KEEP_GRADS", "1")
    from oracle import lxmert_oracle as lxo
    model, mask, s3 = _stage3_model(gold)
    tr, opt, sch = _trainer(model, s3, gold["B"], loss="normal")
    tr._setup_engine(opt)
    assert tr.arena is not None and model.lxmert.encoder._fast_plans() is not None
    assert sorted(n for n, p in model.named_parameters() if p.requires_grad) == gold["trainable"]
    assert s3.see_weight_rate(model, "lxmert") == pytest.approx(gold["zero_rate_pct"], rel=0, abs=1e-9)
    model.eval()
    tr._zero_grad(opt)
    batch = {k: v.cuda() for k, v in lxo.synthetic_batch(gold["B"], gold["A"]).items()}
    loss, logits, pooled = model(batch["ids"], batch["feats"], batch["pos"], labels=batch["target"])[:3]
    loss.backward()
    tr.arena.finalize_grads()
    scale = float(gold["logits"].abs().max())
    assert float((logits.detach().cpu() - gold["logits"]).abs().max()) < 1e-2 * scale      # bf16 operands, 19 layers
    assert float(loss) == pytest.approx(float(gold["loss_normal"]), rel=2e-3)
    params = dict(model.named_parameters())
    worst = 0.0
    for n, st in gold["grad_stats_normal"].items():
        g = params[n].grad
        if n.endswith("key.bias") or st["l2"] < 1e-6:
            continue
        r = abs(float(g.double().norm()) - st["l2"]) / st["l2"]
        worst = max(worst, r)
        assert r < 5e-2, (n, r)                         # same bar as the per-tensor path (bf16 noise floor)
        if n.endswith("weight_orig") and n[: -len("_orig")] + "_mask" in mask:
            assert bool((g[~mask[n[: -len("_orig")] + "_mask"]] == 0).all()), n
    assert worst > 0.0
    for n in gold["nograd_normal"]:                     # tensors the reference leaves without a gradient: zeros here
        assert float(params[n].grad.abs().max()) == 0.0, n


def test_engine_steps_match_per_tensor_path(gold, monkeypatch):
    """Four optimisation steps (dropout off, LMH loss): engine (one-launch Adam over the arena, fused layers) against
    the per-tensor path (masking/pruned.py modules, torch clip_grad_norm_, torch.optim.Adam)."""
    B, A = 16, gold["A"]
    inputs = _inputs(B, A)
    monkeypatch.setenv("CRVQA_FT_ENGINE", "0")
    m_ref, mask, s3 = _stage3_model(gold, dropout=0.0)
    tr_ref, opt_ref, sch_ref = _trainer(m_ref, s3, B)
    m_ref.train()
    ref_losses = []
    for _ in range(4):
        opt_ref.zero_grad()
        loss, _ = tr_ref._training_step(m_ref, inputs, opt_ref)
        torch.nn.utils.clip_grad_norm_(m_ref.parameters(), 1.0)
        opt_ref.step(); sch_ref.step()
        ref_losses.append(float(loss))
    monkeypatch.setenv("CRVQA_FT_ENGINE", "1")
    m_eng, _, _ = _stage3_model(gold, dropout=0.0)
    tr, opt, sch = _trainer(m_eng, s3, B)
    tr.debias_loss_fn.load_state_dict(tr_ref.debias_loss_fn.state_dict())
    tr._setup_engine(opt)
    assert tr.arena is not None
    q = m_eng.lxmert.encoder.layer[0].attention.self.query
    w0, m0 = q.weight_orig.detach().clone(), q.weight_mask.clone()
    tr._zero_grad(opt)
    losses = []
    for _ in range(4):
        loss, _ = tr._device_step(m_eng, inputs, opt)
        sch.step()
        losses.append(float(loss))
    for a, b in zip(losses, ref_losses):
        assert a == pytest.approx(b, rel=5e-3), (losses, ref_losses)
    assert losses[-1] < losses[0]
    pr, pe = dict(m_ref.named_parameters()), dict(m_eng.named_parameters())
    # after 4 Adam steps of lr 5e-5 every trained entry moved by <= ~2e-4; the two paths must agree on the MOVE
    for n in ("lxmert.encoder.layer.0.attention.self.query.weight_orig", "lxmert.encoder.x_layers.2.lang_inter.dense.bias",
              "lxmert.encoder.r_layers.1.output.LayerNorm.weight", "lxmert.pooler.dense.weight_orig",
              "classifier.main.3.weight_v" if "classifier.main.3.weight_v" in pr else sorted(pr)[0]):
        # Adam moves an entry by <= lr per step whatever its gradient's size, so an entry whose (tiny) gradient changes
        # sign under the two paths' different bf16 rounding can differ by up to 2 * 4 * lr = 4e-4: a hard cap at that
        # bound, and all but 1e-4 of the entries within 1.5e-4 (measured: 2 of 589 824 beyond it, max 1.9e-4)
        gap = (pe[n].detach() - pr[n].detach()).abs()
        assert float(gap.max()) <= 4e-4, (n, float(gap.max()))
        assert float((gap > 1.5e-4).float().mean()) <= 1e-4, (n, float((gap > 1.5e-4).float().mean()))
    assert torch.equal(q.weight_mask, m0)
    moved = q.weight_orig.detach() != w0
    assert bool(moved[m0 == 1].any()) and not bool(moved[m0 == 0].any())
    # the GEMM operand the next forward reads is bf16(weight_orig * mask) of the UPDATED weights
    assert torch.equal(q._wm, (q.weight_orig.detach() * q.weight_mask).bfloat16())
    # state_dict keys are unchanged by the arena
    assert sorted(m_eng.state_dict()) == sorted(m_ref.state_dict())


def test_engine_graph_replay_matches_eager(gold, monkeypatch):
    B, A = 16, gold["A"]
    inputs = _inputs(B, A)
    out = []
    for graph in (False, True):
        model, _, s3 = _stage3_model(gold, dropout=0.0)
        tr, opt, sch = _trainer(model, s3, B)
        tr._setup_engine(opt)
        tr._zero_grad(opt)
        stepper = tr._make_graphed_step(model, opt, sch) if graph else None
        if graph:
            assert stepper is not None
        losses = []
        for _ in range(7):
            if graph:
                loss, _ = stepper.step(inputs)
            else:
                loss, _ = tr._device_step(model, inputs, opt)
                sch.step()
            losses.append(float(loss))
        if graph:
            assert stepper.graph is not None
        out.append((losses, model.lxmert.encoder.layer[3].output.dense.weight_orig.detach().clone(),
                    int(opt.state[opt.param_groups[0]["params"][0]]["step"])))
    (l0, w0, s0), (l1, w1, s1) = out
    assert s0 == s1 == 7
    for a, b in zip(l0, l1):
        assert a == pytest.approx(b, rel=2e-3), (l0, l1)
    torch.testing.assert_close(w1, w0, rtol=0, atol=1e-4)
