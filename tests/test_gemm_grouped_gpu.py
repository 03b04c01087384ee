"""Grouped 2-CTA launches (crv_masked_gemm_grouped): every member of a group against fp32 torch math on the same
bf16-rounded operands (2e-3, north_star), the fused GELU epilogues against the standalone GELU kernels' definition,
shared-output score gradients, ragged row counts, and groups longer than one launch."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return float((a.float() - b).abs().max() / b.abs().max().clamp_min(1e-30))


def _rand(*shape, scale=1.0, seed=0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def test_group_of_mixed_problems_matches_fp32_math():
    from crvqa import ops
    M1, M2, H, FF = 5120, 9216, 768, 3072
    x1, x2 = _rand(M1, H, seed=1).bfloat16(), _rand(M2, H, seed=2).bfloat16()
    w_a = _rand(H, H, scale=0.02, seed=3)
    w_b = _rand(FF, H, scale=0.02, seed=4)
    bias_a, bias_b = _rand(H, seed=5), _rand(FF, seed=6)
    dy1, dy2 = _rand(M1, H, seed=7).bfloat16(), _rand(M2, FF, seed=8).bfloat16()
    y1 = torch.empty(M1, H, device="cuda")
    y2 = torch.empty(M2, FF, dtype=torch.bfloat16, device="cuda")
    dx1 = torch.empty(M1, H, dtype=torch.bfloat16, device="cuda")
    ds2 = torch.full((FF, H), 7.0, device="cuda")
    wa16, wb16 = w_a.bfloat16(), w_b.bfloat16()
    ops.gemm_grouped([
        ops.gemm_problem(ops.GEMM_FWD, x1, wa16, y1, bias=bias_a),
        ops.gemm_problem(ops.GEMM_FWD, x2, wb16, y2, bias=bias_b),
        ops.gemm_problem(ops.GEMM_DX, dy1, wa16, dx1),
        ops.gemm_problem(ops.GEMM_DS, dy2, x2, ds2, w_f32=w_b),
    ])
    torch.cuda.synchronize()
    assert _rel(y1, x1.float() @ wa16.float().t() + bias_a) < 2e-3
    assert _rel(y2, x2.float() @ wb16.float().t() + bias_b) < 1e-2           # bf16 output rounding
    assert _rel(dx1, dy1.float() @ wa16.float()) < 1e-2
    assert _rel(ds2, (dy2.float().t() @ x2.float()) * w_b) < 2e-3            # overwrites (7.0 is gone), fp32 W


@pytest.mark.parametrize("M", [256, 640, 1152, 9216])
def test_dx_ds_pair_ragged_rows_and_accumulate(M):
    from crvqa import ops
    N, K = 768, 768
    x, dy = _rand(M, K, seed=11).bfloat16(), _rand(M, N, seed=12).bfloat16()
    w = _rand(N, K, scale=0.02, seed=13)
    w16 = w.bfloat16()
    dx = torch.empty(M, K, dtype=torch.bfloat16, device="cuda")
    ds = torch.empty(N, K, device="cuda")
    pr = lambda acc: [ops.gemm_problem(ops.GEMM_DX, dy, w16, dx),
                      ops.gemm_problem(ops.GEMM_DS, dy, x, ds, w_f32=w, accumulate=acc)]
    ops.gemm_grouped(pr(False))
    torch.cuda.synchronize()
    ref_ds = (dy.float().t() @ x.float()) * w
    assert _rel(dx, dy.float() @ w16.float()) < 1e-2
    assert _rel(ds, ref_ds) < 2e-3
    ops.gemm_grouped(pr(True))
    torch.cuda.synchronize()
    assert _rel(ds, 2 * ref_ds) < 2e-3


def test_shared_module_two_modalities_one_output():
    """Cross layers apply ONE module to both modalities: two DS problems of a call name the same dS and must add."""
    from crvqa import ops
    Ml, Mv, N, K = 5120, 9216, 2304, 768
    xl, xv = _rand(Ml, K, seed=21).bfloat16(), _rand(Mv, K, seed=22).bfloat16()
    dyl, dyv = _rand(Ml, N, seed=23).bfloat16(), _rand(Mv, N, seed=24).bfloat16()
    w = _rand(N, K, scale=0.02, seed=25)
    w16 = w.bfloat16()
    ds = torch.full((N, K), 3.0, device="cuda")
    dxl = torch.empty(Ml, K, dtype=torch.bfloat16, device="cuda")
    dxv = torch.empty(Mv, K, dtype=torch.bfloat16, device="cuda")
    ref = (dyl.float().t() @ xl.float() + dyv.float().t() @ xv.float()) * w
    for order in (0, 1):    # both DS in one launch (order 0) and split over two launches of the call (order 1)
        ds.fill_(3.0)
        if order == 0:
            probs = [ops.gemm_problem(ops.GEMM_DS, dyl, xl, ds, w_f32=w),
                     ops.gemm_problem(ops.GEMM_DS, dyv, xv, ds, w_f32=w, accumulate=True),
                     ops.gemm_problem(ops.GEMM_DX, dyl, w16, dxl), ops.gemm_problem(ops.GEMM_DX, dyv, w16, dxv)]
        else:
            probs = [ops.gemm_problem(ops.GEMM_DX, dyl, w16, dxl), ops.gemm_problem(ops.GEMM_DS, dyl, xl, ds, w_f32=w),
                     ops.gemm_problem(ops.GEMM_DX, dyv, w16, dxv), ops.gemm_problem(ops.GEMM_DX, dyv, w16, dxv),
                     ops.gemm_problem(ops.GEMM_DS, dyv, xv, ds, w_f32=w, accumulate=True)]
        ops.gemm_grouped(probs)
        torch.cuda.synchronize()
        assert _rel(ds, ref) < 2e-3, order
        assert _rel(dxl, dyl.float() @ w16.float()) < 1e-2 and _rel(dxv, dyv.float() @ w16.float()) < 1e-2


@pytest.mark.parametrize("M", [5120, 1152])
def test_fused_gelu_epilogues(M):
    """FF1 forward: aux = bf16(x W^T + b), out = bf16(gelu(aux)) -- the standalone kernel's definition (gelu of the
    ROUNDED pre-activation).  FF2 dX: out = bf16((dY Wm) * gelu'(u))."""
    from crvqa import fused, ops
    H, FF = 768, 3072
    x = _rand(M, H, seed=31).bfloat16()
    w1, b1 = _rand(FF, H, scale=0.04, seed=32).bfloat16(), _rand(FF, scale=0.5, seed=33)
    a = torch.empty(M, FF, dtype=torch.bfloat16, device="cuda")
    u = torch.empty(M, FF, dtype=torch.bfloat16, device="cuda")
    ops.gemm_grouped([ops.gemm_problem(ops.GEMM_FWD, x, w1, a, bias=b1, aux=u, act=ops.ACT_GELU)])
    torch.cuda.synchronize()
    u_ref = x.float() @ w1.float().t() + b1
    assert _rel(u, u_ref) < 1e-2
    assert torch.equal(a, fused.gelu_bf16(u))                                  # same function of the same rounded u
    assert _rel(a, F.gelu(u.float())) < 1e-2
    # FF2 backward
    w2 = _rand(H, FF, scale=0.02, seed=34).bfloat16()
    dy = _rand(M, H, seed=35).bfloat16()
    du = torch.empty(M, FF, dtype=torch.bfloat16, device="cuda")
    ops.gemm_grouped([ops.gemm_problem(ops.GEMM_DX, dy, w2, du, aux=u, act=ops.ACT_GELU)])
    torch.cuda.synchronize()
    uf = u.float().requires_grad_(True)
    F.gelu(uf).backward(dy.float() @ w2.float())
    assert _rel(du, uf.grad) < 1e-2


def test_multi_linear_autograd_matches_single_launch_path(monkeypatch):
    """crvqa.fused.multi_linear (forward group + backward dX/dS group, fused GELU) against the one-GEMM-per-launch
    path on the same arena: same outputs and gradients up to bf16 rounding of intermediates."""
    from crvqa import ops
    from hg_transformers._engine import ScoreArena, masked_modules_of
    from oracle import lxmert_oracle as lxo
    from prune_debias_VQA import build_stage2
    cfg = dict(vocab_size=1000, hidden_size=256, num_attention_heads=4, intermediate_size=512, l_layers=2,
               x_layers=2, r_layers=1, visual_feat_dim=128, max_position_embeddings=32)
    model, masker, _ = build_stage2(96, device=torch.device("cuda"), seed=5, config_kwargs=cfg)
    model.eval()
    batch = {k: v.cuda() for k, v in lxo.synthetic_batch(32, 96, seed=5, T=10, R=8, feat=128, vocab=1000).items()}
    arena = ScoreArena(masked_modules_of(model))
    arena.enable_mask_cache()

    def run():
        arena.begin_step()
        _, logits, _ = model(batch["ids"], batch["feats"], batch["pos"], labels=batch["target"])
        loss, _ = ops.vqa_loss_bce(logits, batch["target"])
        loss.backward()
        arena.finalize_grads()
        return logits.detach().clone(), arena.grads.clone()

    monkeypatch.setenv("CRVQA_GROUPED", "0")
    lg0, g0 = run()
    monkeypatch.setenv("CRVQA_GROUPED", "1")
    c0 = ops.lib.crv_launch_count()
    lg1, g1 = run()
    grouped_launches = ops.lib.crv_launch_count() - c0
    assert float((lg1 - lg0).abs().max() / lg0.abs().max()) < 5e-3
    assert float((g1 - g0).double().norm() / g0.double().norm()) < 2e-2
    print("launches of one fwd+bwd, grouped:", grouped_launches)
