"""Stage-3 frozen-mask fine-tune on the GPU: the pruned modules (masking/pruned.py, through the C ABI) against
the CPU oracle (oracle/stage3.py) and the reference's outputs (tests/golden/stage3_full.pt).
Bit-exact: masks, kept counts, zero rate, the zero pattern of dW_orig, the bf16 pruned operand.
Floating point: 2e-3 per GEMM on identical bf16 operands (fp32 accumulate); end to end the bf16-operand
noise floor documented in DESIGN.md section 2 (tolerances written at each assert)."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def gold():
    return torch.load(os.path.join(GOLD, "stage3_full.pt"), weights_only=False)


def test_mul_cast_bf16_bit_exact():
    from crvqa import ops
    torch.manual_seed(1)
    for n in (8, 4096 + 3, 768 * 3072):
        w = torch.randn(n, device="cuda")
        m = (torch.rand(n, device="cuda") > 0.7).float()
        if n % 8:                       # odd length: only through a 16-byte aligned base
            w, m = w.clone(), m.clone()
        assert torch.equal(ops.mul_cast_bf16(w, m), (w * m).bfloat16())


@pytest.mark.parametrize("M,N,K,bias", [(640, 768, 768, True), (1152, 3072, 768, True), (9216, 768, 3072, False),
                                        (100, 768, 2048, True)])
def test_pruned_linear_vs_oracle(M, N, K, bias):
    from masking.pruned import PrunedLinear
    from oracle import stage3 as o3
    torch.manual_seed(M + N)
    x = torch.randn(M, K) * 0.5
    w = torch.randn(N, K) * 0.02
    m = (torch.rand(N, K) > 0.7).float()
    b = torch.randn(N) * 0.1 if bias else None
    dy = torch.randn(M, N)
    lin = PrunedLinear(torch.nn.Parameter(w.cuda()), m.cuda(), torch.nn.Parameter(b.cuda()) if bias else None)
    xg = x.cuda().requires_grad_(True)
    y = lin(xg)
    y.backward(dy.cuda())
    xo, wo = x.clone().requires_grad_(True), w.clone().requires_grad_(True)
    bo = b.clone().requires_grad_(True) if bias else None
    yo = o3.pruned_linear(xo, wo, m, bo, operand="bf16")
    yo.backward(dy)
    def rel(a, r):
        return float((a.cpu() - r).norm() / r.norm())
    assert rel(y.detach(), yo.detach()) < 2e-3
    assert rel(xg.grad, xo.grad) < 2e-3
    assert rel(lin.weight_orig.grad, wo.grad) < 2e-3
    assert torch.equal(lin.weight_orig.grad.cpu() != 0, wo.grad != 0) or bool((lin.weight_orig.grad.cpu()[m == 0] == 0).all())
    assert bool((lin.weight_orig.grad.cpu()[m == 0] == 0).all())          # exact zeros where the mask is zero
    if bias:
        assert rel(lin.bias.grad, bo.grad) < 1e-5
    assert sorted(lin.state_dict()) == (["bias"] if bias else []) + ["weight_mask", "weight_orig"]


def test_l1_unstructured_mask_exact_count_and_cut():
    from masking.pruned import l1_unstructured_mask
    from oracle import stage3 as o3
    torch.manual_seed(3)
    for shape, amount in (((768, 768), 0.7), ((3072, 768), 0.35), ((30522, 768), 0.7), ((768, 4), 0.5)):
        w = torch.randn(shape) * 0.02
        got = l1_unstructured_mask(w.cuda(), amount).cpu()
        ref = o3.l1_unstructured_mask(w, amount)
        assert int(got.sum()) == int(ref.sum()) == w.numel() - round(amount * w.numel())
        cut = w.abs().reshape(-1).kthvalue(round(amount * w.numel())).values
        differ = got != ref
        assert bool((w.abs()[differ] == cut).all())      # only inside a tie group at the cut (torch: unspecified)
        assert int(differ.sum()) <= 2 * int((w.abs() == cut).sum())


def _stage3_model(gold):
    import run_vqa_stage3 as s3
    from crvqa import ops
    from hg_transformers.modeling_lxmert import LxmertConfig, LxmertForMultipleChoice
    torch.manual_seed(49)
    model = LxmertForMultipleChoice(LxmertConfig(ans_num=gold["A"])).cuda()
    bert = model.lxmert
    mods = dict(bert.named_modules())
    names = s3.trained_mask_module_names()
    ws = [mods[n].weight.detach() for n in names]
    ks = [max(1, int(w.numel() * gold["zero_rate"])) for w in ws]
    thr = ops.kth_value_batched(ws, ks, use_abs=True)           # the stand-in trained mask: |W| > k-th |W|
    mask = {f"lxmert.{n}.weight_mask": (w.abs() > thr[i]) for i, (n, w) in enumerate(zip(names, ws))}
    s3.pruning_model_with_mask(bert, mask, "lxmert")
    return model, mask, s3


def test_full_lxmert_stage3_against_reference(gold):
    """Whole pruned LXMERT (9/5/5, A=2274, B=8, eval mode): masks and bookkeeping bit-exact, logits / loss /
    gradients of all trainable tensors against the fp32 reference."""
    from oracle import lxmert_oracle as lxo
    model, mask, s3 = _stage3_model(gold)
    kept = {k[len("lxmert."):-len(".weight_mask")]: int(v.sum()) for k, v in mask.items()}
    assert kept == gold["kept"]
    assert s3.see_weight_rate(model, "lxmert") == pytest.approx(gold["zero_rate_pct"], rel=0, abs=1e-9)
    assert sorted(n for n, p in model.named_parameters() if p.requires_grad) == gold["trainable"]
    model.eval()
    batch = {k: v.cuda() for k, v in lxo.synthetic_batch(gold["B"], gold["A"]).items()}
    loss, logits, pooled = model(batch["ids"], batch["feats"], batch["pos"], labels=batch["target"])[:3]
    loss.backward()
    scale = float(gold["logits"].abs().max())
    assert float((logits.detach().cpu() - gold["logits"]).abs().max()) < 1e-2 * scale      # bf16 operands, 19 layers
    assert float(loss) == pytest.approx(float(gold["loss_normal"]), rel=2e-3)
    stats = gold["grad_stats_normal"]
    params = dict(model.named_parameters())
    assert sorted(n for n, p in params.items() if p.requires_grad and p.grad is None) == gold["nograd_normal"]
    worst = 0.0
    for n, st in stats.items():
        g = params[n].grad
        if n.endswith("key.bias") or st["l2"] < 1e-6:
            continue                                    # identically zero in exact arithmetic (softmax shift)
        r = abs(float(g.double().norm()) - st["l2"]) / st["l2"]
        worst = max(worst, r)
        assert r < 5e-2, (n, r)                         # gradient norms: bf16 noise floor through 19 layers
        if n.endswith("weight_orig"):
            m = mask[n[: -len("_orig")] + "_mask"]
            assert bool((g[~m] == 0).all()), n          # exact zeros where the mask is zero
    assert worst > 0.0


def test_mag_pruning_matches_reference_counts(gold):
    import run_vqa_stage3 as s3
    from hg_transformers.modeling_lxmert import LxmertConfig, LxmertForMultipleChoice
    torch.manual_seed(49)
    model = LxmertForMultipleChoice(LxmertConfig(ans_num=gold["A"])).cuda()
    s3.mag_pruning(model.lxmert, 0.7)
    pm = {n: m for n, m in model.lxmert.named_modules() if hasattr(m, "weight_mask")}
    assert sorted(pm) == gold["mag_pruned_modules"]
    assert {n: int(m.weight_mask.sum()) for n, m in pm.items()} == gold["mag_kept"]
    for n, samp in gold["mag_mask_sample"].items():
        flat = pm[n].weight_mask.reshape(-1)
        assert torch.equal(flat[:: max(1, flat.numel() // 512)][:512].bool().cpu(), samp), n


def test_stage3_training_steps_reduce_loss_and_keep_masked_weights_frozen(gold):
    """A few Adam steps with the stage-2 Trainer (training_type FT_trainedMask, LMH loss): the loss goes down,
    the masks do not change and pruned positions of weight_orig never move."""
    from hg_transformers.data.data_collator import TrimCollator
    from hg_transformers.mask_trainer_VQA import Trainer
    from hg_transformers.training_args import TrainingArguments
    from oracle import lxmert_oracle as lxo
    from prune_debias_VQA import ModelArguments
    model, mask, s3 = _stage3_model(gold)
    targs = TrainingArguments(output_dir="/tmp/crvqa_stage3", per_gpu_train_batch_size=8, logging_steps=1000, seed=49,
                              training_type="FT_trainedMask", FT_type="lmh", save_steps=0, dataloader_num_workers=0,
                              learning_rate=5e-5)
    opt, sch = s3.init_optimizer(model, targs, 8 * 100)
    margs = ModelArguments()
    tr = Trainer(model=model, args=targs, model_args=margs, data_collator=TrimCollator(), optimizers=(opt, sch), masker=None)
    host = lxo.synthetic_batch(8, gold["A"])
    inputs = [host[k].cuda() if k else torch.arange(8) for k in ["ids", "feats", "pos", "target", None, None, "bias", "max_label"]]
    q = model.lxmert.encoder.layer[0].attention.self.query
    w0, m0 = q.weight_orig.detach().clone(), q.weight_mask.clone()
    model.train()
    losses = []
    for _ in range(6):
        opt.zero_grad()
        loss, _ = tr._training_step(model, inputs, opt)
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        opt.step(); sch.step()
        losses.append(float(loss))
    assert losses[-1] < losses[0], losses
    assert torch.equal(q.weight_mask, m0)
    moved = q.weight_orig.detach() != w0
    assert bool(moved[m0 == 1].any()) and not bool(moved[m0 == 0].any())
