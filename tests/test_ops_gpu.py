"""GPU parity of every C-ABI op against the reference's golden vectors (tests/golden/ops.pt) and the
CPU oracle.  Integer / index / mask work is bit-exact; floating point within the stated tolerance."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


@pytest.fixture(scope="module")
def g():
    return torch.load(os.path.join(GOLD, "ops.pt"), weights_only=False)


def dev(t):
    return t.cuda()


def test_binarize_bit_exact(g):
    from crvqa import ops
    out, cnt = ops.binarize(dev(g["bin_in"]), g["bin_thr"], want_count=True)
    assert torch.equal(out.cpu(), g["bin_out"])
    assert int(cnt) == int(g["bin_out"].sum())
    b = ops.binarize(dev(g["bin_in"]), dev(g["bin_thr"]), as_bool=True)
    assert b.dtype == torch.bool and torch.equal(b.cpu(), g["bin_out"].bool())


def test_kth_value_golden_cases(g):
    from crvqa import ops
    cases = g["kth_cases"]
    got = ops.kth_value_batched([dev(c["x"]) for c in cases], [c["k"] for c in cases]).cpu()
    got_abs = ops.kth_value_batched([dev(c["x"]) for c in cases], [c["k"] for c in cases], use_abs=True).cpu()
    for i, c in enumerate(cases):
        assert float(got[i]) == float(c["v"]), i          # == treats -0 and +0 alike
        assert float(got_abs[i]) == float(c["v_abs"]), i


@pytest.mark.parametrize("seed", [0, 1])
def test_kth_value_batched_vs_oracle(seed):
    """Ragged batch: tiny / large / tie-heavy / negative / constant segments, ranks at both ends."""
    from crvqa import ops
    from oracle import masked_ops as o
    gen = torch.Generator().manual_seed(seed)
    segs = [torch.randn(1, generator=gen), torch.randn(3072, generator=gen) * 1e-2,
            torch.randn(589824, generator=gen) * 0.02, torch.rand(2359296, generator=gen) * 0.04 - 0.01,
            torch.where(torch.rand(1572864, generator=gen) < 0.7, torch.zeros(1572864), torch.full((1572864,), 0.02)),
            torch.full((70001,), 0.02), -torch.rand(8193, generator=gen), torch.randn(8191, generator=gen).abs() * 1e-30]
    ks = [1, 2150, 412876, 1651507, 1101004, 70001, 1, 8191]
    got = ops.kth_value_batched([dev(s) for s in segs], ks).cpu()
    for i, (s, k) in enumerate(zip(segs, ks)):
        assert float(got[i]) == float(o.kth_value(s, k)), (i, float(got[i]))
    ks2 = [1, 1, 1, 2359296, 1572864, 1, 8193, 4000]
    got = ops.kth_value_batched([dev(s) for s in segs], ks2).cpu()
    for i, (s, k) in enumerate(zip(segs, ks2)):
        assert float(got[i]) == float(o.kth_value(s, k)), (i, float(got[i]))


def test_kth_value_sortedness_property_large():
    """BASELINE-size segment (word embeddings, 23.4 M): the result must have exactly-consistent ranks."""
    from crvqa import ops
    gen = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(23440896, device="cuda", generator=gen) * 0.02
    k = int(23440896 * 0.7)
    v = ops.kth_value_batched([x], [k])[0]
    below = int((x < v).sum())
    upto = int((x <= v).sum())
    assert below < k <= upto


def _kth_ref(x, k, use_abs=False):
    v = x.abs() if use_abs else x
    return torch.kthvalue(v.reshape(-1).float(), k).values


def test_kth_value_front_end_edge_cases():
    """Segments large enough for the sample -> filter -> select front end: NaN / inf members, ranks at both ends
    (open pivot windows), |x| keys, an unaligned segment, pivots that tie massively, and a segment whose
    stratified sample is unrepresentative on purpose (every sampled run holds huge values), which must fall
    back to the full radix select.  All bit-exact against torch.kthvalue."""
    from crvqa import ops
    gen = torch.Generator(device="cuda").manual_seed(11)
    n = 700001
    base = torch.randn(n, device="cuda", generator=gen) * 0.01
    with_inf = base.clone(); with_inf[::1000] = float("inf"); with_inf[1::1000] = float("-inf")
    with_nan = base.clone(); with_nan[5::997] = float("nan")
    ties = torch.where(torch.rand(n, device="cuda", generator=gen) < 0.7, 0.0, 0.02)
    ties[::3] += torch.randn(ties[::3].numel(), device="cuda", generator=gen) * 1e-3
    unaligned = torch.randn(n + 3, device="cuda", generator=gen)[3:]
    assert unaligned.data_ptr() % 16 != 0
    adversarial = torch.randn(589824, device="cuda", generator=gen) * 0.01
    idx = torch.arange(589824, device="cuda")
    run = 589824 // 2048
    adversarial[(idx % run) < 8] = 1000.0          # exactly the elements the sampler looks at
    neg_zero = torch.zeros(n, device="cuda"); neg_zero[::2] = -0.0; neg_zero[::7] = 1.0; neg_zero[::11] = -1.0
    cases = [(base, 1), (base, n), (base, 3), (base, n - 2), (base, n // 2), (with_inf, 600), (with_inf, n - 500),
             (with_inf, n // 3), (with_nan, n - 5), (with_nan, n // 2), (ties, int(n * 0.7)), (ties, int(n * 0.2)),
             (ties, int(n * 0.9)), (unaligned, n // 5), (adversarial, int(589824 * 0.7)), (adversarial, 589824 - 100),
             (neg_zero, n // 2), (torch.full((n,), 0.02, device="cuda"), n // 2)]
    for use_abs in (False, True):
        got = ops.kth_value_batched([c[0] for c in cases], [c[1] for c in cases], use_abs=use_abs)
        for i, (x, k) in enumerate(cases):
            ref = _kth_ref(x, k, use_abs)
            if torch.isnan(ref):
                assert torch.isnan(got[i]), (i, use_abs)
            else:
                assert float(got[i]) == float(ref), (i, use_abs, float(got[i]), float(ref))


def test_kth_value_small_workspace_takes_radix_core():
    """With the count-only workspace size of the C ABI every segment takes the 3-pass radix core; same answers."""
    import ctypes
    from crvqa import _lib, ops
    gen = torch.Generator(device="cuda").manual_seed(12)
    segs = [torch.randn(900001, device="cuda", generator=gen), torch.rand(40000, device="cuda", generator=gen),
            torch.randn(300000, device="cuda", generator=gen).abs()]
    ks = [450000, 17, 299999]
    count = len(segs)
    nbytes = _lib.lib.crv_kth_value_workspace_bytes(count)
    ns = (ctypes.c_longlong * count)(*[s.numel() for s in segs])
    assert nbytes < _lib.lib.crv_kth_value_workspace_bytes_for(ns, count)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    out = torch.empty(count, device="cuda")
    ptrs = (ctypes.c_void_p * count)(*[s.data_ptr() for s in segs])
    kk = (ctypes.c_longlong * count)(*ks)
    _lib.check(_lib.lib.crv_kth_value_batched(ptrs, ns, kk, count, 0, out.data_ptr(), ws.data_ptr(), nbytes, None), "kth")
    for i, (x, k) in enumerate(zip(segs, ks)):
        assert float(out[i]) == float(_kth_ref(x, k))
    assert torch.equal(out, ops.kth_value_batched(segs, ks))


def test_kth_plan_reuse_tracks_changing_scores():
    """The cached plan (trainer path) holds pointers, not values: new scores and new ranks give new thresholds."""
    from crvqa import ops
    gen = torch.Generator(device="cuda").manual_seed(13)
    segs = [torch.randn(589824, device="cuda", generator=gen) for _ in range(5)] + [torch.randn(3072, device="cuda", generator=gen)]
    plan = ops.KthPlan(segs)
    for rate in (0.3, 0.7, 0.95):
        for s in segs:
            s.add_(torch.randn(s.shape, device="cuda", generator=gen) * 0.1)
        ks = [max(1, int(s.numel() * rate)) for s in segs]
        got = plan(ks)
        for i, (x, k) in enumerate(zip(segs, ks)):
            assert float(got[i]) == float(_kth_ref(x, k)), (rate, i)


def test_magnitude_init(g):
    from crvqa import ops
    w = dev(g["ml_weight"])
    k = int(w.numel() * 0.7)
    thr = ops.kth_value_batched([w], [k], use_abs=True)
    s = ops.magnitude_init(w, thr[0:1], 2.0 * 1e-2, 0.0 * 1e-2)
    assert torch.equal(s.cpu(), g["ml_scores_init"])


def _masked_module(g, name, weight, bias, scores, padding_idx=None):
    from masking.maskers import MaskedLinear1
    info = {"structured_masking": None, "structured_masking_types": None, "force_masking": "bert", "ptl_config": None}
    w = torch.nn.Parameter(dev(weight), requires_grad=False)
    b = torch.nn.Parameter(dev(bias), requires_grad=False) if bias is not None else None
    m = MaskedLinear1(weight=w, bias=b, mask_biases=False, name=name, padding_idx=padding_idx,
                      threshold=torch.tensor(1e-2), init_sparsity=0.7, init_scale=2e-2, controlled_init="magnitude",
                      structured_masking_info=info)
    return m


def test_masked_linear_module_vs_reference_and_oracle(g):
    """Reference MaskedLinear1 golden (fp32) vs the tcgen05 path (K = 24 rows are TMA-legal): bf16
    operand rounding bounds the gap at 2e-2 of the output scale; the mask itself is bit-exact."""
    m = _masked_module(g, "x.dense", g["ml_weight"], g["ml_bias"], None)
    assert torch.equal(m.weight_mask.detach().cpu(), g["ml_scores_init"])   # magnitude init, bit-exact
    m.weight_mask.data.copy_(dev(g["ml_scores"]))
    x = dev(g["ml_x"]).requires_grad_(True)
    y = m(x)
    y.backward(dev(g["ml_dy"]))
    for got, ref in ((y, g["ml_y"]), (x.grad, g["ml_dx"]), (m.weight_mask.grad, g["ml_ds"])):
        err = float((got.detach().cpu() - ref).abs().max() / ref.abs().max())
        assert err < 2e-2, err
    mw, mb = m.get_masks()
    assert mb is None and torch.equal(mw.cpu(), (g["ml_scores"] > 1e-2).float())


@pytest.mark.parametrize("M,N,K", [(60, 768, 768), (640, 3072, 768), (1152, 768, 3072)])
def test_masked_linear_tcgen05_vs_oracle(M, N, K):
    """TMA / tcgen05 path: against the oracle with bf16-rounded operands 2e-3 (north_star tolerance);
    against exact fp32 math the gap is the bf16 operand rounding (reported, bounded at 2e-2)."""
    from masking.maskers import MaskedLinear1  # noqa: F401
    from oracle import masked_ops as o
    gen = torch.Generator().manual_seed(11)
    w = torch.randn(N, K, generator=gen) * 0.02
    bias = torch.randn(N, generator=gen) * 0.1
    x = torch.randn(2, M // 2, K, generator=gen)
    dy = torch.randn(2, M // 2, N, generator=gen)
    m = _masked_module(None, "enc.dense", w, bias, None)
    m.weight_mask.data.add_(dev(torch.randn(N, K, generator=gen) * 0.01))
    s_cpu = m.weight_mask.detach().cpu()
    xg = dev(x).requires_grad_(True)
    y = m(xg)
    y.backward(dev(dy))
    for operand, tol in (("bf16", 2e-3), ("fp32", 2e-2)):
        xo = x.clone().requires_grad_(True)
        so = s_cpu.clone().requires_grad_(True)
        yo = o.masked_linear(xo, so, w, 1e-2, bias, operand)
        yo.backward(dy)
        for got, ref in ((y, yo), (xg.grad, xo.grad), (m.weight_mask.grad, so.grad)):
            err = float((got.detach().cpu() - ref.detach()).abs().max() / ref.detach().abs().max())
            assert err < tol, (operand, err)


def test_masked_embedding(g):
    m = _masked_module(g, "emb.word_embeddings", g["emb_weight"], None, None, padding_idx=0)
    assert torch.equal(m.weight_mask.detach().cpu(), g["emb_scores"])
    e = m(dev(g["emb_ids"]))
    assert torch.equal(e.cpu(), g["emb_out"])                                # gather + mask: bit-exact
    e.backward(dev(g["emb_dout"]))
    torch.testing.assert_close(m.weight_mask.grad.cpu(), g["emb_ds"], rtol=1e-6, atol=1e-9)
    assert float(m.weight_mask.grad[0].abs().sum()) == 0.0                   # padding row


def test_losses(g):
    from crvqa import ops
    r = g["loss"]
    logits, labels, bias = dev(r["logits"]), dev(r["labels"]), dev(r["bias"])
    lg = logits.clone().requires_grad_(True)
    loss, score = ops.vqa_loss_bce(lg, labels)
    loss.backward()
    torch.testing.assert_close(loss.cpu(), r["bce"], rtol=2e-6, atol=1e-6)
    torch.testing.assert_close(lg.grad.cpu(), r["bce_dlogits"], rtol=1e-5, atol=1e-8)
    assert float(score) == float(r["score"])

    lg = logits.clone().requires_grad_(True)
    loss, score = ops.vqa_loss_lpf(lg, bias, dev(r["max_label"]), 5.0, labels)
    loss.backward()
    torch.testing.assert_close(loss.cpu(), r["lpf"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(lg.grad.cpu(), r["lpf_dlogits"], rtol=1e-4, atol=1e-8)
    assert float(score) == float(r["score"])

    from hg_transformers.vqa_debias_loss_functions import LearnedMixin
    lm = LearnedMixin(0.36).cuda()
    lm.bias_lin.weight.data.copy_(r["lin_w"])
    lm.bias_lin.bias.data.copy_(r["lin_b"])
    lm.smooth_param.data.copy_(r["smooth_param"])
    lg = logits.clone().requires_grad_(True)
    pooled = dev(r["pooled"]).requires_grad_(True)
    loss = lm(pooled, lg, bias, labels, "cuda")
    loss.backward()
    torch.testing.assert_close(loss.cpu(), r["lmh"], rtol=1e-5, atol=1e-5)
    torch.testing.assert_close(lg.grad.cpu(), r["lmh_dlogits"], rtol=1e-4, atol=1e-8)
    torch.testing.assert_close(pooled.grad.cpu(), r["lmh_dpooled"], rtol=1e-4, atol=1e-8)
    assert float(lm.last_score) == float(r["score"])


def test_clip_adamw(g):
    from optimization import AdamW
    from crvqa import ops
    t = g["adamw"]
    ps = [torch.nn.Parameter(dev(p.clone())) for p in t["p0"]]
    opt = AdamW([{"params": [p]} for p in ps], lr=5e-5, eps=1e-8)
    for step in range(3):
        sumsq = torch.zeros((), device="cuda")
        for p, gr in zip(ps, t["grads"][step]):
            p.grad = dev(gr.clone())
            ops.sumsq_into(p.grad, sumsq)
        opt.set_clip(sumsq, 1.0)
        opt.step()
        for p, ref, rs in zip(ps, t["p"][step], t["sum"][step]):
            torch.testing.assert_close(p.detach().cpu(), ref, rtol=2e-6, atol=1e-9)
            torch.testing.assert_close(opt.state[p]["sum"].cpu(), rs, rtol=2e-6, atol=1e-9)


def test_cast_and_apply_mask():
    from crvqa import ops
    gen = torch.Generator().manual_seed(2)
    x = torch.randn(1000003, generator=gen)
    assert torch.equal(ops.to_bf16(dev(x)).cpu(), x.bfloat16())
    s = torch.rand(4099, generator=gen)
    w = torch.randn(4099, generator=gen).bfloat16()
    out = ops.apply_mask_bf16(dev(w), dev(s), torch.tensor(0.5))
    assert torch.equal(out.cpu(), torch.where(s > 0.5, w, torch.zeros_like(w)))


def test_apply_mask_segmented_bit_exact():
    """The launch that writes EVERY masked operand of the timed step (ScoreArena.refresh_masked): per-module
    thresholds, ragged module sizes, ties at the threshold, +-0 -- bit for bit against oracle binarize x W."""
    from crvqa import ops
    from oracle import masked_ops as o
    gen = torch.Generator().manual_seed(11)
    sizes = [8192 * 3, 768 * 768, 8, 24, 8192 + 40, 3072 * 768, 136 * 72]
    offs, off = [], 0
    for n in sizes:
        offs.append(off)
        off += (n + 63) // 64 * 64
    w = torch.randn(off, generator=gen).bfloat16()
    s = torch.rand(off, generator=gen) * 0.02
    s[torch.rand(off, generator=gen) < 0.3] = 0.0
    s[torch.rand(off, generator=gen) < 0.05] = -0.0
    thr = torch.tensor([0.0, 0.01, 0.02, -1.0, 5e-5, 0.0139, 0.007])
    for i, n in enumerate(sizes):           # exact ties at every module's threshold
        s[offs[i]: offs[i] + n: 7] = thr[i]
    rows = []
    for i, n in enumerate(sizes):
        for c0 in range(0, n, 8192):
            rows.append(((offs[i] + c0) // 8, min(8192, n - c0), i, 0))
    chunks = torch.tensor(rows, dtype=torch.int32)
    wm = torch.full((off,), 7.0).bfloat16().cuda()
    ops.apply_mask_segmented(dev(w), dev(s), dev(thr), dev(chunks), wm)
    wm = wm.cpu()
    for i, n in enumerate(sizes):
        sl = slice(offs[i], offs[i] + n)
        mask = o.binarize(s[sl], float(thr[i]))
        want = torch.where(mask > 0, w[sl], torch.zeros_like(w[sl]))
        assert torch.equal(wm[sl].view(torch.int16), want.view(torch.int16)), i
        pad = wm[offs[i] + n: (offs[i + 1] if i + 1 < len(sizes) else off)]
        assert bool((pad.float() == 7.0).all())      # alignment padding between modules is never written


def test_secondary_debias_losses_on_gpu():
    """RUBI_loss and BiasProduct through the fused CUDA kernels against the reference's own
    values and logit gradients (tests/golden/secondary_losses.pt)."""
    from crvqa import ops
    from hg_transformers._trainer_core import RUBI_loss
    from hg_transformers.vqa_debias_loss_functions import BiasProduct
    g = torch.load(os.path.join(GOLD, "secondary_losses.pt"), weights_only=False)
    lg = dev(g["logits"]).requires_grad_(True)
    loss = RUBI_loss(lg, dev(g["bias"]), dev(g["max_label"]))
    loss.backward()
    assert abs(float(loss) - float(g["rubi"])) <= 1e-5 * abs(float(g["rubi"]))
    torch.testing.assert_close(lg.grad.cpu(), g["rubi_dlogits"], rtol=1e-4, atol=1e-7)
    l2, score = ops.vqa_loss_rubi(lg.detach(), dev(g["bias"]), dev(g["max_label"]), dev(g["labels"]))
    want = g["labels"].gather(1, g["logits"].argmax(1, keepdim=True)).sum()
    assert float(l2) == float(loss) and abs(float(score) - float(want)) < 1e-6
    bp = BiasProduct().cuda()
    if g["bp_smooth_param"] is not None:
        bp.smooth_param.data.copy_(g["bp_smooth_param"])
    lg2 = dev(g["logits"]).requires_grad_(True)
    loss = bp(dev(g["hidden"]), lg2, dev(g["bias"]), dev(g["labels"]))
    loss.backward()
    assert abs(float(loss) - float(g["bp"])) <= 1e-5 * abs(float(g["bp"]))
    torch.testing.assert_close(lg2.grad.cpu(), g["bp_dlogits"], rtol=1e-4, atol=1e-7)


@pytest.mark.parametrize("M,N,K", [(9216, 768, 4), (300, 770, 3), (77, 768, 8), (129, 64, 12), (5, 9, 1)])
def test_masked_linear_small_k(M, N, K):
    """The fp32 SIMT masked linear for inner dimensions TMA cannot take (LXMERT's box_fc: K = 4; reference
    masking/maskers.py:359-366 with a 4-wide input): K <= 8 runs the coalesced kernels (lanes along N, atomic row-slice
    partials for dS), larger K the general ones.  Forward / dX / dS against fp32 torch math on the same mask."""
    from crvqa import ops
    g = torch.Generator().manual_seed(M + N + K)
    x = torch.randn(M, K, generator=g).cuda().requires_grad_(True)
    w = (torch.randn(N, K, generator=g) * 0.3).cuda()
    s = torch.rand(N, K, generator=g).cuda()
    s[0, 0] = 0.5                                        # a tie: strict > masks it out
    thr = torch.tensor(0.5, device="cuda")
    b = torch.randn(N, generator=g).cuda()
    scores = s.clone().requires_grad_(True)
    y = ops.MaskedLinearSmallKFn.apply(x, scores, w, thr, b)
    # references in float64: other tests of the suite switch torch's fp32 matmuls to TF32 for the whole process
    wm = (w * (s > thr).float()).double()
    ref = x.detach().double() @ wm.t() + b.double()
    assert float((y - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
    dy = torch.randn(M, N, generator=g).cuda()
    y.backward(dy)
    ref_dx = dy.double() @ wm
    ref_ds = (dy.double().t() @ x.detach().double()) * w.double()
    assert float((x.grad - ref_dx).abs().max()) <= 1e-5 * float(ref_dx.abs().max())
    assert float((scores.grad - ref_ds).abs().max()) <= 2e-5 * float(ref_ds.abs().max())


def test_momentum_update_is_bit_identical_to_the_torch_expression():
    """crv_momentum_update over many separately allocated tensors against param_m = param_m * m + param * (1 - m)
    (mPLUG/models/model_vqa_mplug.py:152-156; the three foreach passes of the PyTorch path): same two rounded products,
    same rounded sum -- bit for bit, including sizes that are not multiples of 4 and tensors larger than one chunk."""
    from crvqa import ops
    g = torch.Generator().manual_seed(12)
    shapes = [(768, 768), (3,), (1,), (30522, 32), (17, 5), (16384,), (16385,), (4, 4, 4, 3)]
    online = [torch.randn(*sh, generator=g).cuda() for sh in shapes]
    twins = [(t + 0.01 * torch.randn(*t.shape, generator=g).cuda()).contiguous() for t in online]
    want = [b.clone() for b in twins]
    m = 0.995
    for _ in range(3):
        fresh = torch._foreach_mul(online, 1.0 - m)
        torch._foreach_mul_(want, m)
        torch._foreach_add_(want, fresh)
    plan = ops.MomentumPlan(online, twins)
    assert plan.key == ops.MomentumPlan.key_of(online, twins)
    for _ in range(3):
        plan.run(m)
    for a, b in zip(twins, want):
        assert torch.equal(a, b)
