"""mPLUG masking path, CPU tier:

* oracle/mplug_masking.py against the reference's outputs (tests/golden/mplug_skeleton.pt, make_golden_mplug.py);
* the host logic of the drop-in ``mPLUG/masking/maskers.py`` / ``vqa_mplug.py`` / ``engine.py`` with the oracle
  injected as a fake kernel backend (no CUDA compute runs here), against the same golden file;
* the bf16-score comparison threshold as a brute-force property.
"""
import contextlib
import hashlib
import io
import logging
import os
import sys
import types

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import mplug_skeleton as sk  # noqa: E402
from oracle import masked_ops as o_ops  # noqa: E402
from oracle import mplug_masking as om  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mplug_skeleton.pt")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


def digest(tensors):
    h = hashlib.sha256()
    for k in sorted(tensors):
        h.update(k.encode())
        h.update(tensors[k].detach().cpu().float().contiguous().numpy().tobytes())
    return h.hexdigest()


def masked(model):
    return [(n, m) for n, m in model.named_modules() if hasattr(m, "threshold")]


def thr_record(model):
    rec = {}
    for n, m in masked(model):
        t = m.threshold
        rec[n] = (float(t), str(t.dtype).replace("torch.", "") if torch.is_tensor(t) else type(t).__name__)
    return rec


def perturb(model, seed, scale):
    g = torch.Generator().manual_seed(seed)
    for _, m in masked(model):
        m.weight_mask.data.add_((torch.randn(m.weight_mask.shape, generator=g) * scale).to(m.weight_mask.device))


def kept(model):
    return {n: int(m.get_masks()[0].sum()) for n, m in masked(model)}


def quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def fresh(gold):
    model = sk.build(gold["skeleton_seed"])
    assert digest(model.state_dict()) == gold["state_dict_sha256"], "skeleton weights differ from the golden run"
    return model


# ----------------------------------------------------------------------------- oracle vs reference
def test_oracle_chain_module_names(gold):
    abbr = {"visual_encoder": ["AO_visual", "I_visual", "O_visual", "AO", "I", "O", "E"],
            "text_encoder": ["K", "Q", "V", "AO", "I", "O", "E"],
            "fusion_encoder": ["SK", "SQ", "SV", "SAO", "CK", "CQ", "CV", "CAO", "I", "O", "E"],
            "text_decoder": ["SK", "SQ", "SV", "SAO", "CK", "CQ", "CV", "CAO", "I", "O", "E"]}
    for tower, ab in abbr.items():
        assert sorted(om.chain_module_names(tower, list(range(3)), ab)) == gold["chain"][tower]


def _oracle_train_state(model):
    for p in model.parameters():
        p.grad = None
    model.train()
    loss = model(*sk.batch())
    loss.backward()
    return float(loss.detach()), {n: m.weight_mask.grad for n, m in masked(model)}


def test_oracle_variant_a_matches_reference(gold):
    A = gold["A"]
    model = fresh(gold)
    names = sk.names_to_mask(om.chain_module_names)
    assert sorted(names) == A["names_tobe_masked"]
    om.patch(model, names, init_sparsity=A["init_sparsity"], controlled_init="magnitude_soft")
    assert [n for n, _ in masked(model)] == A["module_names"]
    assert sorted(n for n, p in model.named_parameters() if p.requires_grad) == A["trainable"]
    assert thr_record(model) == A["init_thresholds"]
    assert kept(model) == A["kept_init"]
    for n, m in masked(model):      # Masker.init_masks: binarised against the MASKER's 1e-2, not the module's threshold
        assert np.array_equal(om.packed(om.mask_of(m.weight_mask, 1e-2)), A["init_masks"][f"{n}_weight_mask"])
    loss, grads = _oracle_train_state(model)
    assert loss == pytest.approx(A["loss"], rel=1e-5)
    for n, g in grads.items():
        if A["grads"][n] is None:
            assert g is None                    # attn.out_proj: nn.MultiheadAttention reads .weight, never calls it
        else:
            assert torch.allclose(g, A["grads"][n], rtol=1e-4, atol=1e-7), n
    masks = [m.get_masks()[0] for _, m in masked(model)]
    sizes = [(n, p.numel()) for n, p in model.named_parameters()]
    assert round(om.see_sparsity(masks, sizes), 2) == A["start_see_sparsity"]
    assert round(om.zero_rate(masks), 2) == A["start_zero_rate"]

    perturb(model, *A["perturb"])
    for r in A["resets"]:
        mean = om.reset_threshold(model, r["rate"])
        assert thr_record(model) == r["thresholds"]
        assert kept(model) == r["kept"]
        assert mean == r["mean"]
    for n, m in masked(model):
        assert np.array_equal(om.packed(m.get_masks()[0]), A["after_masks"][n + ".weight"])
    before = thr_record(model)
    om.reset_threshold(model, 1e-4)
    assert (before == thr_record(model)) == A["tiny_rate_moves_nothing"]
    loss, grads = _oracle_train_state(model)
    assert loss == pytest.approx(A["after_train"]["loss"], rel=1e-5)
    for n, g in grads.items():
        want = A["after_train"]["grad_norms"][n]
        assert (g is None) if want is None else float(g.norm()) == pytest.approx(want, rel=1e-4)

    # (D) the same scores seen through a bf16 model copy
    D = gold["D"]
    assert digest({n: m.weight_mask for n, m in masked(model)}) == D["fp32_scores_sha256"]
    assert thr_record(model) == D["fp32_thresholds"]
    for _, m in masked(model):
        m.score_dtype = torch.bfloat16
    assert kept(model) == D["kept_before"]
    for r in D["resets"]:
        mean = om.reset_threshold(model, r["rate"])
        assert thr_record(model) == r["thresholds"]
        assert mean == r["mean"]
        for n, m in masked(model):
            assert np.array_equal(om.packed(m.get_masks()[0]), r["masks"][n]), n


def test_oracle_ramp_and_global_variants_match_reference(gold):
    from masking import sparsity_control as sp   # the drop-in's scheduler: pure host code
    import types
    B0 = gold["B0"]
    model = fresh(gold)
    names = sk.names_to_mask(om.chain_module_names)
    om.patch(model, names, init_sparsity=0.0, controlled_init="magnitude_soft")
    assert thr_record(model) == B0["init_thresholds"] and kept(model) == B0["kept_init"]
    assert all(v == (0.0, "int") for v in B0["init_thresholds"].values())
    assert B0["reset_at_zero"].startswith("RuntimeError")         # the reference cannot average a list of ints

    B = gold["B"]
    model = fresh(gold)
    om.patch(model, names, init_sparsity=B["init_sparsity"], controlled_init="magnitude_soft")
    conf = types.SimpleNamespace(
        masking_scheduler_conf_={"lambdas_lr": 0.0, "sparsity_warmup": "automated_gradual_sparsity",
                                 "sparsity_warmup_interval_epoch": 0.1, "init_epoch": 0.0, "final_epoch": 4,
                                 "final_sparsity": 0.7, "init_sparsity": 0.1},
        logger=logging.getLogger("t"))
    sched = sp.MaskerScheduler(conf)
    for r in B["ramp"]:
        _, target, changed = sched.step(cur_epoch=r["epoch"])
        assert target == r["target"] and changed == r["changed"]
        mean = om.reset_threshold(model, target)
        assert thr_record(model) == r["thresholds"] and kept(model) == r["kept"] and mean == r["mean"]
    cs = B["constant_scores"]
    mod = dict(masked(model))[cs["module"]]
    mod.weight_mask.data.fill_(0.25)
    om.reset_threshold(model, 0.5)
    assert float(mod.threshold) == cs["threshold_after"] == cs["threshold_before"]

    C = gold["C"]
    model = fresh(gold)
    cut = om.patch(model, names, init_sparsity=C["init_sparsity"], controlled_init="magnitude", global_prune=True)
    assert cut == C["global_weight_threshold"]
    assert kept(model) == C["kept_init"] and thr_record(model) == C["init_thresholds"]
    loss, grads = _oracle_train_state(model)
    assert loss == pytest.approx(C["loss"], rel=1e-5)
    for n, g in grads.items():
        want = C["grad_norms"][n]
        assert (g is None) if want is None else float(g.norm()) == pytest.approx(want, rel=1e-4)
    perturb(model, *C["perturb"])
    for r in C["global_resets"]:
        mean = om.reset_threshold(model, r["rate"], global_prune=True)
        assert thr_record(model) == r["thresholds"] and kept(model) == r["kept"] and mean == r["mean"]


# ----------------------------------------------------------------------------- bf16 comparison threshold
def test_bf16_score_threshold_is_exact():
    """bf16(S) > bf16(t)  <=>  S > T for every fp32 S, checked on dense neighbourhoods of both rounding boundaries."""
    from mPLUG.masking.maskers import bf16_score_threshold
    g = torch.Generator().manual_seed(0)
    ts = torch.cat([torch.randn(400, generator=g) * 0.05, torch.randn(100, generator=g) * 30,
                    torch.tensor([0.0, -0.0, 1e-2, 1.0, -1.0, 0.0078125, 0.00390625, 1e-40, -1e-40, 3e-39])])
    T = bf16_score_threshold(ts)
    assert T.dtype == torch.float32 and T.shape == ts.shape
    offs = torch.cat([torch.arange(-70000, 70000, 13), torch.arange(-33000, -32500), torch.arange(32500, 33000),
                      torch.arange(-4, 5)]).to(torch.int64)
    for t, Ti in zip(ts, T):
        t16 = t.to(torch.bfloat16)
        base = int(t16.float().view(torch.int32))
        key = (-(base & 0x7FFFFFFF) if base < 0 else base) + offs         # monotone integer image of fp32
        bits = torch.where(key < 0, (-key) | 0x80000000, key)
        bits = torch.where(bits >= 2 ** 31, bits - 2 ** 32, bits).to(torch.int32)
        S = bits.view(torch.float32)
        S = S[torch.isfinite(S)]
        assert torch.equal(S.to(torch.bfloat16) > t16, S > Ti), float(t)


# ----------------------------------------------------------------------------- drop-in host logic, oracle as backend
@pytest.fixture()
def oracle_backend(monkeypatch):
    from crvqa import ops

    def kth(tensors, ks, use_abs=False):
        return torch.tensor([float(o_ops.kth_value(t, int(k), use_abs=use_abs)) for t, k in zip(tensors, ks)])

    def mag(weight, w_thr, hi, lo):
        keep = weight.detach().abs() > float(w_thr)
        return torch.where(keep, torch.full_like(weight, hi), torch.full_like(weight, lo))

    def binz(scores, thr, want_count=False, as_bool=False):
        m = o_ops.binarize(scores.detach(), float(thr))
        out = m.bool() if as_bool else m
        return (out, m.sum().long()) if want_count else out

    monkeypatch.setattr(ops, "kth_value_batched", kth)
    monkeypatch.setattr(ops, "magnitude_init", mag)
    monkeypatch.setattr(ops, "binarize", binz)
    monkeypatch.setattr(ops, "_stage", lambda t: t)
    return ops


def _conf(**over):
    from mPLUG.masking.mask_config import MaskConfigs
    conf = MaskConfigs()
    for k, v in over.items():
        setattr(conf, k, v)
    return conf


def _init(model, **over):
    from mPLUG import vqa_mplug
    return quiet(vqa_mplug.init_masker, _conf(**over), model, weight_types=sk.WEIGHT_TYPES,
                 layers_to_mask=sk.LAYERS)


def test_dropin_chain_names_and_config(gold):
    from mPLUG.masking import maskers
    from mPLUG.masking.mask_config import MaskConfigs
    from mPLUG import vqa_mplug
    for tower, names in gold["chain"].items():
        ab = {"visual_encoder": ["AO_visual", "I_visual", "O_visual", "AO", "I", "O", "E"],
              "text_encoder": ["K", "Q", "V", "AO", "I", "O", "E"]}.get(
            tower, ["SK", "SQ", "SV", "SAO", "CK", "CQ", "CV", "CAO", "I", "O", "E"])
        assert sorted(maskers.chain_module_names(tower, list(range(3)), ab)) == names
    c = MaskConfigs()
    assert (c.zero_rate, c.threshold, c.controlled_init, c.masker_update_step, c.train_classifier) == \
        (0.5, 1e-2, "magnitude_soft", 100, True)
    assert vqa_mplug.encode_maskconfig(c) is c.__dict__ and vqa_mplug.encode_maskconfig(3) == 3
    full = vqa_mplug.names_to_mask(c)
    # 12 x 2 ViT MLPs + 6 x 6 text + 6 x 10 fusion + 12 x 10 decoder, each with its _m twin
    assert len(full) == 2 * (24 + 36 + 60 + 120)
    c.mask_classifier = True
    assert "text_decoder_m.cls.predictions.transform.dense" in vqa_mplug.names_to_mask(c)


def test_dropin_masker_host_logic_variant_a(gold, oracle_backend):
    from mPLUG.masking import maskers
    A = gold["A"]
    model = fresh(gold)
    masker = _init(model, zero_rate=0.7)
    assert sorted(masker.masker_scheduler.conf.masking_scheduler_conf_) == sorted(
        ["lambdas_lr", "sparsity_warmup", "sparsity_warmup_interval_epoch", "init_epoch", "final_epoch",
         "final_sparsity"])
    assert [n for n, _ in masked(model)] == A["module_names"]
    assert sorted(n for n, p in model.named_parameters() if p.requires_grad) == A["trainable"]
    assert thr_record(model) == A["init_thresholds"]
    assert kept(model) == A["kept_init"]
    for n, m in masked(model):
        assert np.array_equal(om.packed(masker.init_masks[f"{n}_weight_mask"]), A["init_masks"][f"{n}_weight_mask"])
        assert m.weight is dict(model.named_parameters())[n + ".weight"] and not m.weight.requires_grad
        assert isinstance(m, maskers.MaskedLinear1)
    assert round(quiet(maskers.see_sparsity, model), 2) == A["start_see_sparsity"]
    assert round(quiet(maskers.save_model_mask, model, is_save=False), 2) == A["start_zero_rate"]

    perturb(model, *A["perturb"])
    for r in A["resets"]:
        mean = maskers.reset_threshold(model, r["rate"])
        assert thr_record(model) == r["thresholds"]
        assert kept(model) == r["kept"]
        assert mean == r["mean"]
    before = thr_record(model)
    maskers.reset_threshold(model, 1e-4)
    assert before == thr_record(model)

    # bf16 score mode: fp32 master scores, masks and thresholds of the reference's bf16 model copy
    D = gold["D"]
    maskers.set_score_dtype(model, torch.bfloat16)
    assert kept(model) == D["kept_before"]
    for r in D["resets"]:
        mean = maskers.reset_threshold(model, r["rate"])
        assert thr_record(model) == r["thresholds"]
        assert mean == r["mean"]
        for n, m in masked(model):
            assert np.array_equal(om.packed(m.get_masks()[0]), r["masks"][n]), n


def test_dropin_masker_host_logic_ramp_global_and_export(gold, oracle_backend, tmp_path):
    from mPLUG import vqa_mplug
    from mPLUG.masking import maskers
    B0 = gold["B0"]
    model = fresh(gold)
    _init(model, zero_rate=0.7, init_sparsity=0.0, final_sparsity_epoch=4)
    assert thr_record(model) == B0["init_thresholds"] and kept(model) == B0["kept_init"]
    assert maskers.reset_threshold(model, 0.0) == 0.0   # documented difference: the reference raises here

    B = gold["B"]
    model = fresh(gold)
    masker = _init(model, zero_rate=0.7, init_sparsity=0.1, final_sparsity_epoch=4)
    out_dir = str(tmp_path / "masks")
    for r in B["ramp"]:
        mean, target = quiet(vqa_mplug.update_masks, model, masker, r["epoch"], out_dir)
        assert target == r["target"] and mean == r["mean"]
        assert thr_record(model) == r["thresholds"] and kept(model) == r["kept"]
    saved = torch.load(os.path.join(out_dir, "mask.pt"))
    assert sorted(saved) == sorted(n + ".weight" for n, _ in masked(model))
    for n, m in masked(model):
        assert saved[n + ".weight"].dtype == torch.float32 and torch.equal(saved[n + ".weight"], m.get_masks()[0])
    cs = B["constant_scores"]
    mod = dict(masked(model))[cs["module"]]
    mod.weight_mask.data.fill_(0.25)
    maskers.reset_threshold(model, 0.5)
    assert float(mod.threshold) == cs["threshold_after"]

    C = gold["C"]
    model = fresh(gold)
    masker = _init(model, zero_rate=0.6, init_sparsity=0.5, controlled_init="magnitude", global_prune=True)
    assert float(masker.global_threshold) == C["global_weight_threshold"]
    assert kept(model) == C["kept_init"] and thr_record(model) == C["init_thresholds"]
    perturb(model, *C["perturb"])
    for r in C["global_resets"]:
        mean = maskers.reset_threshold(model, r["rate"], global_prune=True)
        assert thr_record(model) == r["thresholds"] and kept(model) == r["kept"] and mean == r["mean"]


def test_load_mask_and_prune_reparametrises_the_named_modules(gold, oracle_backend, tmp_path):
    from mPLUG import vqa_mplug
    from mPLUG.masking.pruned import PrunedLinear   # the shared module, as this package sees it
    model = fresh(gold)
    g = torch.Generator().manual_seed(1)
    target = "text_encoder.encoder.layer.0.intermediate.dense"
    mask = (torch.rand(model.text_encoder.encoder.layer[0].intermediate.dense.weight.shape, generator=g) > 0.5).float()
    torch.save({"module." + target + ".weight": mask}, tmp_path / "mask.pt")
    quiet(vqa_mplug.load_mask_and_prune, str(tmp_path), model)
    mod = model.text_encoder.encoder.layer[0].intermediate.dense
    assert isinstance(mod, PrunedLinear) and torch.equal(mod.weight_mask, mask)
    assert "text_encoder.encoder.layer.0.intermediate.dense.weight_orig" in dict(model.named_parameters())


def test_engine_backward_advances_the_dropout_counter_of_the_fused_kernels(monkeypatch):
    """The fused layer kernels hash their dropout masks from a device-side (seed, counter); MaskTrainEngine.backward()
    moves the counter of every device that drew masks, once per call, after the backward has been queued."""
    from crvqa import fused
    from mPLUG.engine import MaskTrainEngine
    events = []

    class Counter:
        def advance(self):
            events.append("advance")

    monkeypatch.setattr(fused.RngState, "_per_device", {("cuda", 0): Counter(), ("cuda", 1): Counter()})
    net = torch.nn.Linear(4, 2)
    net.weight.register_hook(lambda g: events.append("backward"))
    eng = MaskTrainEngine(net, torch.optim.SGD(net.parameters(), lr=0.1))
    eng.backward(eng(torch.randn(3, 4)).sum())
    assert events == ["backward", "advance", "advance"]
    eng.backward(eng(torch.randn(3, 4)).sum())
    assert events.count("advance") == 4


def test_engine_step_protocol_on_cpu():
    """The DeepSpeed-engine stand-in: backward / clip / optimiser step / counters, on a plain module."""
    from mPLUG.engine import MaskTrainEngine
    torch.manual_seed(0)
    net = torch.nn.Linear(8, 4)
    opt = torch.optim.SGD(net.parameters(), lr=0.1)
    eng = MaskTrainEngine(net, opt, gradient_clipping=0.5)
    x = torch.randn(16, 8)
    w0 = net.weight.detach().clone()
    loss = eng(x).pow(2).sum()
    eng.backward(loss)
    g = net.weight.grad.clone()
    gb = net.bias.grad.clone()
    norm = torch.sqrt(g.pow(2).sum() + gb.pow(2).sum())
    eng.step()
    assert eng.global_steps == 1 and net.weight.grad is None
    assert float(eng.last_grad_norm) == pytest.approx(float(norm), rel=1e-6)
    assert torch.allclose(net.weight, w0 - 0.1 * g * (0.5 / (norm + 1e-6)), atol=1e-7)
    assert [n for n, _ in eng.named_parameters()] == ["weight", "bias"]


# ----------------------------------------------------------------------------- engine: data parallel (gloo, 2 ranks)
def _engine_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                    "compress-robust-vqa_b200"))
    from mPLUG.engine import MaskTrainEngine
    torch.manual_seed(0)                                   # identical replicas
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    net[0].bias.requires_grad = False                      # a frozen tensor must stay out of the exchange
    opt = torch.optim.SGD([p for p in net.parameters() if p.requires_grad], lr=0.05)
    eng = MaskTrainEngine(net, opt, gradient_clipping=1e9)
    g = torch.Generator().manual_seed(100)
    xs = [torch.randn(4, 6, generator=g) for _ in range(world)]       # every rank knows every shard
    # what one process would do on the concatenated batch (mean of per-rank mean losses == mean over the union)
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    ref.load_state_dict(net.state_dict())
    for step in range(2):
        loss = eng(xs[rank]).pow(2).mean()
        eng.backward(loss)
        eng.step()
        ref.zero_grad()
        torch.stack([ref(x).pow(2).mean() for x in xs]).mean().backward()
        with torch.no_grad():
            for n, p in ref.named_parameters():
                if n != "0.bias":
                    p -= 0.05 * p.grad
        for (n, p), (_, pr) in zip(net.named_parameters(), ref.named_parameters()):
            assert torch.allclose(p, pr, atol=1e-6), (rank, step, n)
    assert eng.global_steps == 2
    dist.barrier()
    dist.destroy_process_group()
    q.put(rank)


def test_engine_allreduces_trainable_gradients_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_engine_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(2)) == [0, 1]


def _fused_engine_worker(rank, world, port, q):
    """Two data-parallel ranks through the engine's multi-tensor step (kernels emulated, gloo): gradients averaged,
    one clip coefficient, identical parameters on both ranks == one process on the union batch with torch's AdamW."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.path.join(root, "compress-robust-vqa_b200"), root, os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from crvqa import ops
    from mPLUG import engine as eng_mod
    calls = _emulated_multi_kernels(None, ops)
    eng_mod._FusedAdamW.device_types = ("cpu", "cuda")
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    net[0].bias.requires_grad = False
    groups = [{"params": [net[0].weight, net[2].weight], "weight_decay": 0.05},
              {"params": [net[2].bias], "weight_decay": 0.0}]
    opt = torch.optim.AdamW(groups, lr=0.02)
    eng = eng_mod.MaskTrainEngine(net, opt, gradient_clipping=0.05)
    ref = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    ref.load_state_dict(net.state_dict())
    ref[0].bias.requires_grad = False
    ropt = torch.optim.AdamW([{"params": [ref[0].weight, ref[2].weight], "weight_decay": 0.05},
                              {"params": [ref[2].bias], "weight_decay": 0.0}], lr=0.02)
    g = torch.Generator().manual_seed(100)
    xs = [torch.randn(4, 6, generator=g) for _ in range(world)]
    for step in range(3):
        loss = eng(xs[rank]).pow(2).mean()
        eng.backward(loss)
        eng.step()
        ropt.zero_grad()
        torch.stack([ref(x).pow(2).mean() for x in xs]).mean().backward()
        norm = torch.nn.utils.clip_grad_norm_([p for p in ref.parameters() if p.requires_grad], 0.05)
        ropt.step()
        assert float(norm) > 0.05 and float(eng.last_grad_norm) == pytest.approx(float(norm), rel=1e-5)
        for (n, p), (_, pr) in zip(net.named_parameters(), ref.named_parameters()):
            assert torch.allclose(p, pr, rtol=1e-5, atol=1e-7), (rank, step, n)
    assert calls["sumsq"] == 3 and calls["adamw"] == 6 and eng.global_steps == 3
    gathered = [torch.zeros_like(net[2].weight) for _ in range(world)]
    dist.all_gather(gathered, net[2].weight.detach())
    assert torch.equal(gathered[0], gathered[1])                 # bit-identical replicas
    dist.barrier()
    dist.destroy_process_group()
    q.put(rank)


def test_fused_engine_step_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_fused_engine_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(2)) == [0, 1]


def test_held_masked_operand_is_rebuilt_exactly_when_it_must(gold, oracle_backend, monkeypatch):
    """The engine-managed operand cache of the masked modules: built once, reused while nothing changes, rebuilt after
    the engine's step (drop), after a threshold refresh (new threshold object) and after a score-dtype switch."""
    from mPLUG.engine import MaskTrainEngine
    from mPLUG.masking import maskers
    calls = []

    def fake_apply(w16, scores, thr):
        calls.append(float(thr))
        return (w16.float() * (scores > thr).float()).to(torch.bfloat16)

    monkeypatch.setattr(oracle_backend, "apply_mask_bf16", fake_apply)
    monkeypatch.setattr(oracle_backend, "to_bf16", lambda x: x.to(torch.bfloat16))
    model = fresh(gold)
    _init(model, zero_rate=0.7)
    mod = dict(masked(model))["text_encoder.encoder.layer.0.intermediate.dense"]
    dev = torch.device("cpu")
    assert mod._held_masked_weight(mod._threshold_on(dev)) is None            # off until an engine opts in
    eng = MaskTrainEngine(model, torch.optim.SGD([mod.weight_mask], lr=0.1), bf16=False)
    a = mod._held_masked_weight(mod._threshold_on(dev))
    b = mod._held_masked_weight(mod._threshold_on(dev))
    assert a is b and len(calls) == 1
    assert torch.equal(a.float(), mod.weight.bfloat16().float() * mod.get_masks()[0])
    mod.weight_mask.grad = torch.ones_like(mod.weight_mask)
    eng.step()                                                                # scores moved
    c = mod._held_masked_weight(mod._threshold_on(dev))
    assert c is not a and len(calls) == 2
    maskers.reset_threshold(model, 0.5)                                       # new threshold objects
    d = mod._held_masked_weight(mod._threshold_on(dev))
    assert d is not c and len(calls) == 3 and calls[-1] == float(mod.threshold)
    maskers.set_score_dtype(model, torch.bfloat16)                            # same threshold, other comparison
    e = mod._held_masked_weight(mod._threshold_on(dev))
    assert e is not d and len(calls) == 4
    assert calls[-1] == float(maskers.bf16_score_threshold(torch.tensor(float(mod.threshold))))
    eng.invalidate_masks()
    assert mod._wm is None


# ----------------------------------------------------------------------------- engine: fused clip + AdamW step, host logic
def _at(addr, n, ctype, dtype):
    """numpy view of n elements at a raw host address (the tensors the address tables name live on the CPU here)."""
    import ctypes
    return np.frombuffer((ctype * n).from_address(int(addr)), dtype=dtype)


def _emulated_multi_kernels(monkeypatch, ops):
    """crv_sumsq_multi / crv_adamw_multi restated in numpy over the SAME address and row tables the engine hands to the
    CUDA kernels (include/crvqa.h; csrc/elementwise.cu sumsq_multi_kernel / adamw_multi_kernel), so the engine's host
    logic -- tables, groups, step counters, plan rebuilds, operand hand-over -- runs on a box without a GPU."""
    import ctypes
    f32 = np.float32
    calls = {"sumsq": 0, "adamw": 0, "upload": 0}

    def upload(values, dtype, device, out=None):
        calls["upload"] += 1
        t = torch.tensor(values, dtype=dtype)
        if out is None:
            return t
        out.copy_(t)
        return out

    def sumsq_multi(ptrs, rows, acc):
        calls["sumsq"] += 1
        tot = 0.0
        for t, first8, n, _ in rows.tolist():
            x = _at(int(ptrs[t]) + first8 * 32, n, ctypes.c_float, f32)
            tot += float(np.sum(x.astype(np.float64) ** 2))
        acc += tot

    def adamw_multi(p, g, m, v, w16, wm, thr, rows, lr, step, b1, b2, eps, wd, total_sumsq=None, max_norm=1.0):
        calls["adamw"] += 1
        clip = f32(1.0)
        if total_sumsq is not None:
            clip = min(f32(1.0), f32(max_norm) / (np.sqrt(f32(total_sumsq.item())) + f32(1e-6)))
        decay, omb1, omb2 = f32(1.0 - lr * wd), f32(1.0 - b1), f32(1.0 - b2)
        step_size, inv_bc2 = f32(lr / (1.0 - b1 ** step)), f32(1.0 / np.sqrt(1.0 - b2 ** step))
        for t, first8, n, flags in rows.tolist():
            off = first8 * 8
            P, G, M, V = (_at(int(tab[t]) + off * 4, n, ctypes.c_float, f32) for tab in (p, g, m, v))
            gc = G * clip
            P *= decay
            M += (gc - M) * omb1
            V[:] = V * f32(b2) + omb2 * gc * gc
            P -= step_size * (M / (np.sqrt(V) * inv_bc2 + f32(eps)))
            if flags & 1:
                T = _at(int(thr[t]), 1, ctypes.c_float, f32)[0]
                W = _at(int(w16[t]) + off * 2, n, ctypes.c_uint16, np.uint16)
                _at(int(wm[t]) + off * 2, n, ctypes.c_uint16, np.uint16)[:] = np.where(P > T, W, np.uint16(0))

    put = monkeypatch.setattr if monkeypatch is not None else setattr     # a spawned worker has no monkeypatch fixture
    put(ops, "upload", upload)
    put(ops, "sumsq_multi", sumsq_multi)
    put(ops, "adamw_multi", adamw_multi)
    return calls


def test_fused_engine_step_host_logic_follows_clip_plus_torch_adamw(gold, oracle_backend, monkeypatch):
    """MaskTrainEngine.step() with a stock torch AdamW: address / row tables per parameter group, torch-compatible
    optimiser state, plan rebuilds (new threshold objects, parameters that start receiving gradients), the refreshed
    masked operand handed to the modules -- against clip_grad_norm_ + torch.optim.AdamW.step on cloned tensors fed the
    same gradients.  The two kernels are emulated in numpy over the engine's own tables (no GPU here)."""
    from mPLUG import engine as eng_mod
    from mPLUG import optim as mplug_optim
    from mPLUG.masking import maskers
    ops = oracle_backend
    applied = []

    def fake_apply(w16, scores, thr):
        applied.append(1)
        return (w16.float() * (scores > thr).float()).to(torch.bfloat16)

    monkeypatch.setattr(ops, "apply_mask_bf16", fake_apply)
    monkeypatch.setattr(ops, "to_bf16", lambda x: x.to(torch.bfloat16))
    calls = _emulated_multi_kernels(monkeypatch, ops)
    monkeypatch.setattr(eng_mod._FusedAdamW, "device_types", ("cpu", "cuda"))
    model = fresh(gold)
    _init(model, zero_rate=0.6, init_sparsity=0.2, final_sparsity_epoch=2)
    args = types.SimpleNamespace(opt="adamW", lr=2e-3, weight_decay=0.02)
    opt = mplug_optim.create_optimizer(args, model)
    eng = eng_mod.MaskTrainEngine(model, opt, gradient_clipping=0.05, bf16=True)
    trainable = [p for g in opt.param_groups for p in g["params"]]
    twin = [[p.detach().clone().requires_grad_(True) for p in g["params"]] for g in opt.param_groups]
    ref = torch.optim.AdamW([{"params": twin[0], "weight_decay": 0.0}, {"params": twin[1], "weight_decay": 0.02}],
                            lr=2e-3)
    flat_twin = [q for grp in twin for q in grp]
    held = [m for _, m in masked(model) if m.holds_masked_operand()]
    assert len(held) >= 20
    late = {id(held[0].weight_mask), id(held[3].weight_mask)}       # these get no gradient in the first two steps
    gen = torch.Generator().manual_seed(7)
    plans = []
    for it in range(5):
        for g, h in zip(opt.param_groups, ref.param_groups):
            g["lr"] = h["lr"] = 2e-3 * (1.0 - 0.1 * it)
        for p, q in zip(trainable, flat_twin):
            if it < 2 and id(p) in late:
                p.grad = q.grad = None
                continue
            p.grad = torch.randn(p.shape, generator=gen) * 0.01
            q.grad = p.grad.clone()
        if it == 3:
            maskers.reset_threshold(model, 0.5)                       # new threshold objects: the plan must follow
        for m in held:                                               # a forward would fetch the operand here
            m._held_masked_weight(m._threshold_on(torch.device("cpu")))
        n_applied = len(applied)
        before = dict(calls)
        eng.step()
        plans.append(eng._fused.plan)
        if it >= 2:
            # two scores joined with step counters behind the others' in their group: the kernels take one step number
            # per launch, so that group is stepped by two launches from now on
            assert not eng._fused.plan.uniform_steps and calls["adamw"] == before["adamw"] + 3
            assert calls["sumsq"] == before["sumsq"] + 1
        want_norm = torch.nn.utils.clip_grad_norm_([q for q in flat_twin if q.grad is not None], 0.05)
        ref.step()
        assert float(eng.last_grad_norm) == pytest.approx(float(want_norm), rel=1e-5) and float(want_norm) > 0.05
        for p, q in zip(trainable, flat_twin):
            assert p.grad is None
            assert torch.allclose(p.detach(), q.detach(), rtol=2e-6, atol=2e-8), float((p - q).abs().max())
            if len(opt.state[p]):
                assert float(opt.state[p]["step"]) == float(ref.state[q]["step"])
                assert torch.allclose(opt.state[p]["exp_avg_sq"], ref.state[q]["exp_avg_sq"], rtol=1e-5, atol=1e-30)
        if it < 2:
            assert calls["sumsq"] == before["sumsq"] + 1 and calls["adamw"] == before["adamw"] + 2
            refreshed = [m for m in eng._fused.plan.mods if m is not None]
            assert len(refreshed) == len(held) - 2
            for m in refreshed:
                thr = m._threshold_on(torch.device("cpu"))
                assert m._wm_key == m._wm_key_now(thr)
                assert torch.equal(m._wm.float(), m.weight.bfloat16().float() * m.get_masks()[0])
                assert m._held_masked_weight(thr) is m._wm            # no rebuild on the next forward
            assert len(applied) == n_applied
            assert held[0]._wm is None                                # outside the pass: dropped, rebuilt on demand
    assert plans[0] is plans[1] and plans[1] is not plans[2] and plans[2] is not plans[3] and plans[3] is plans[4]
    # a plain optimizer.step() on SOME of the tensors splits a run of equal step counts: the plan is rebuilt around it
    for p in trainable:
        p.grad = None
    lone = held[5].weight_mask
    lone.grad = torch.zeros_like(lone)
    opt.step()
    for p, q in zip(trainable, flat_twin):
        p.grad = torch.randn(p.shape, generator=gen) * 0.01
        q.grad = p.grad.clone()
    flat_twin[[id(p) for p in trainable].index(id(lone))].grad = None     # the twin takes the lone step now ...
    before = dict(calls)
    eng.step()
    assert eng._fused.plan is not plans[4] and calls["adamw"] == before["adamw"] + 4
    sd = opt.state_dict()
    assert len(sd["state"]) == len(trainable)
    steps = {int(st["step"]) for st in opt.state.values()}
    assert steps == {4, 6, 7}                  # late joiners, everyone else, the one tensor stepped on its own as well


def test_fused_engine_step_uniform_run_rebuilds_plan_on_new_thresholds(gold, oracle_backend, monkeypatch):
    """Every trainable tensor has a gradient from the first step (the training loop's case): all steps take the fused
    pass, a threshold refresh rebuilds the plan once, and the operand after it is the new mask's."""
    from mPLUG import engine as eng_mod
    from mPLUG.masking import maskers
    ops = oracle_backend
    monkeypatch.setattr(ops, "apply_mask_bf16", lambda w16, s, t: (w16.float() * (s > t).float()).to(torch.bfloat16))
    monkeypatch.setattr(ops, "to_bf16", lambda x: x.to(torch.bfloat16))
    calls = _emulated_multi_kernels(monkeypatch, ops)
    monkeypatch.setattr(eng_mod._FusedAdamW, "device_types", ("cpu", "cuda"))
    model = fresh(gold)
    _init(model, zero_rate=0.6, init_sparsity=0.2, final_sparsity_epoch=2)
    params = [p for p in model.parameters() if p.requires_grad]
    opt = torch.optim.AdamW(params, lr=1e-3, weight_decay=0.0)
    eng = eng_mod.MaskTrainEngine(model, opt, gradient_clipping=1.0, bf16=True)
    twin = [p.detach().clone().requires_grad_(True) for p in params]
    ref = torch.optim.AdamW(twin, lr=1e-3, weight_decay=0.0)
    gen = torch.Generator().manual_seed(11)
    plans = []
    for it in range(4):
        for p, q in zip(params, twin):
            p.grad = torch.randn(p.shape, generator=gen) * 1e-3       # norm below 1: the clip stays off
            q.grad = p.grad.clone()
        if it == 2:
            maskers.reset_threshold(model, 0.55)
        eng.step()
        plans.append(eng._fused.plan)
        assert float(torch.nn.utils.clip_grad_norm_(twin, 1.0)) < 1.0
        ref.step()
        for p, q in zip(params, twin):
            assert torch.allclose(p.detach(), q.detach(), rtol=2e-6, atol=2e-8)
        for m in (m for m in plans[-1].mods if m is not None):
            assert torch.equal(m._wm.float(), m.weight.bfloat16().float() * m.get_masks()[0])
    assert calls["adamw"] == 4 and calls["sumsq"] == 4
    assert plans[0] is plans[1] and plans[1] is not plans[2] and plans[2] is plans[3]
    assert calls["upload"] == 2 * 2 + 4         # two tables per plan build + the gradient addresses of every step
    # CRVQA_MPLUG_FUSED=0 keeps the PyTorch passes
    monkeypatch.setenv("CRVQA_MPLUG_FUSED", "0")
    assert not eng_mod._FusedAdamW.wanted(opt)
    assert not eng_mod._FusedAdamW.wanted(torch.optim.SGD(params, lr=0.1))


def test_optimizer_groups_and_cosine_schedule_match_reference():
    """mPLUG/optim + mPLUG/scheduler against the reference packages (tests/golden/mplug_host.json): same parameter
    groups (members, lr, weight decay) for create_optimizer / create_two_optimizer, same learning rate after every
    scheduler.step(t) incl. the warm-up line, repeated / out-of-order calls, restarts with t_mul and decay_rate."""
    import json
    import types

    from mPLUG.optim import create_optimizer, create_two_optimizer
    from mPLUG.scheduler import create_scheduler
    with open(os.path.join(os.path.dirname(GOLD), "mplug_host.json")) as f:
        G = json.load(f)
    model = sk.build()
    for n, p in model.named_parameters():
        p.requires_grad = ("predictions" in n) or n.endswith("intermediate.dense.weight")
    names = {id(p): n for n, p in model.named_parameters()}
    vis = {id(p): "visual_encoder." + n for n, p in model.visual_encoder.named_parameters()}

    def groups(opt):
        return [{"lr": g["lr"], "weight_decay": g["weight_decay"],
                 "params": [names.get(id(p), vis.get(id(p))) for p in g["params"]]} for g in opt.param_groups]

    a = types.SimpleNamespace(**G["opt"])
    two = create_two_optimizer(a, model)
    assert groups(two) == G["two_optimizer_groups"]
    one = create_optimizer(a, model)
    assert groups(one) == G["optimizer_groups"]
    assert isinstance(one, torch.optim.AdamW) and isinstance(two, torch.optim.AdamW)
    sch, epochs = create_scheduler(types.SimpleNamespace(**G["sched"]), two)
    assert epochs == G["num_epochs"]
    assert [g["lr"] for g in two.param_groups] == G["lr_after_init"]
    for t, want in G["lr_by_step"]:
        sch.step(t)
        assert [g["lr"] for g in two.param_groups] == pytest.approx(want, rel=1e-12, abs=0), t
    sch2, _ = create_scheduler(types.SimpleNamespace(**dict(G["sched"], warmup_epochs=0, lr_cycle_limit=2,
                                                            lr_cycle_mul=2.0, decay_rate=0.5, epochs=3)), one)
    for t, want in G["lr_by_step_cycles"]:
        sch2.step(t)
        assert [g["lr"] for g in one.param_groups] == pytest.approx(want, rel=1e-12, abs=0), t
    assert sch2.get_cycle_length() == G["cycle_length"]
    with pytest.raises(NotImplementedError):
        create_scheduler(types.SimpleNamespace(**dict(G["sched"], sched="tanh")), one)
    with pytest.raises(NotImplementedError):
        create_optimizer(types.SimpleNamespace(**dict(G["opt"], opt="lookahead_adam")), model)


def test_driver_utils():
    from mPLUG import utils
    d = utils.AttrDict({"opt": "adamW", "lr1": 3e-5})
    assert d.opt == "adamW" and d["lr1"] == 3e-5
    d.lr2 = 5e-6
    assert d["lr2"] == 5e-6
    assert utils.get_rank() == 0 and utils.get_world_size() == 1 and utils.is_main_process()
    assert utils.compute_n_params(torch.nn.Linear(1000, 2000)) == "2.0M"
    log = utils.MetricLogger(delimiter="  ")
    log.add_meter("loss", utils.SmoothedValue(window_size=1, fmt="{value:.4f}"))
    seen = []
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()) as out:
        for x in log.log_every([1.0, 3.0], 1, "hdr"):
            log.update(loss=torch.tensor(x), lr=0.5)
            seen.append(x)
    assert seen == [1.0, 3.0] and "loss: 3.0000" in out.getvalue()
    assert log.meters["loss"].global_avg == 2.0 and log.global_avg() == "loss: 2.0000  lr: 0.5000"
    args = types.SimpleNamespace()
    for k in ("RANK", "WORLD_SIZE"):
        os.environ.pop(k, None)
    with contextlib.redirect_stdout(io.StringIO()):
        utils.init_distributed_mode(args)
    assert args.distributed is False


def test_bf16_score_threshold_property_random_bit_patterns():
    """Same equivalence on uniformly random fp32 BIT PATTERNS for both threshold and score (all exponents, both signs,
    subnormals), plus scores placed exactly on and next to the two rounding boundaries of every threshold."""
    from mPLUG.masking.maskers import bf16_score_threshold
    g = torch.Generator().manual_seed(123)
    tbits = torch.randint(-2 ** 31, 2 ** 31 - 1, (4000,), generator=g, dtype=torch.int64).to(torch.int32)
    t = tbits.view(torch.float32)
    t = t[torch.isfinite(t) & (t.abs() < 1e38)]
    T = bf16_score_threshold(t)
    t16 = t.to(torch.bfloat16)
    sbits = torch.randint(-2 ** 31, 2 ** 31 - 1, (t.numel(), 64), generator=g, dtype=torch.int64).to(torch.int32)
    S = sbits.view(torch.float32)
    ok = torch.isfinite(S)
    want = S.to(torch.bfloat16) > t16[:, None]
    got = S > T[:, None]
    assert torch.equal(want[ok], got[ok])
    # the boundary itself and its two fp32 neighbours, for every threshold
    for delta in (-1, 0, 1):
        Tb = T.view(torch.int32).to(torch.int64)
        key = torch.where(Tb < 0, -(Tb & 0x7FFFFFFF), Tb) + delta
        bits = torch.where(key < 0, (-key) | 0x80000000, key)
        bits = torch.where(bits >= 2 ** 31, bits - 2 ** 32, bits).to(torch.int32)
        Sb = bits.view(torch.float32)
        fin = torch.isfinite(Sb)
        assert torch.equal((Sb.to(torch.bfloat16) > t16)[fin], (Sb > T)[fin]), delta


def test_oracle_training_trajectory_matches_reference(gold):
    """Six AdamW steps on scores + LM head with a mask update every two steps (scheduler.step -> reset_threshold): the
    oracle-patched network follows the reference's losses, thresholds (bf16, exact) and kept counts."""
    import types

    from masking import sparsity_control as sp
    T = gold["T"]
    model = fresh(gold)
    om.patch(model, sk.names_to_mask(om.chain_module_names), init_sparsity=T["init_sparsity"],
             controlled_init="magnitude_soft")
    assert thr_record(model) == T["init_thresholds"] and kept(model) == T["kept_init"]
    sched = sp.MaskerScheduler(types.SimpleNamespace(
        masking_scheduler_conf_={"lambdas_lr": 0.0, "sparsity_warmup": "automated_gradual_sparsity",
                                 "sparsity_warmup_interval_epoch": 0.1, "init_epoch": 0.0, "final_epoch": 2,
                                 "final_sparsity": 0.7, "init_sparsity": 0.3}, logger=logging.getLogger("t")))
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=T["lr"], weight_decay=0.0)
    data = sk.batch()
    model.train()
    for step, want in enumerate(T["steps"]):
        loss = model(*data)
        opt.zero_grad()
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in model.parameters() if p.requires_grad and p.grad is not None], 1.0)
        opt.step()
        assert float(loss.detach()) == pytest.approx(want["loss"], rel=1e-4), step
        if "target" in want:
            _, target, _ = sched.step(cur_epoch=(step + 1) // 2)
            assert target == want["target"]
            mean = om.reset_threshold(model, target)
            assert thr_record(model) == want["thresholds"], step
            assert mean == want["mean"]
            got = kept(model)
            assert all(abs(got[n] - want["kept"][n]) <= 2 for n in got), step
