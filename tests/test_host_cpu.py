"""CPU tier: host logic of the drop-in package, the C-ABI surface, and the multi-rank gradient exchange
(gloo, world_size 2).  No CUDA compute is called; where a host routine needs the result of a kernel, the
CPU oracle is injected as a fake backend (tests may import oracle/, the product never does)."""
import ctypes
import json
import logging
import os
import re
import sys
import types

import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
WEIGHT_TYPES = ["E", "VV", "VB", "lK", "lQ", "lV", "lAO", "lI", "lO", "vK", "vQ", "vV", "vAO", "vI", "vO",
                "vlVK", "vlVQ", "vlVV", "vlVAO", "vlLaK", "vlLaQ", "vlLaV", "vlLaAO", "vlVaK", "vlVaQ", "vlVaV",
                "vlVaAO", "vlLi", "vlLo", "vlVi", "vlVo", "P"]


@pytest.fixture(scope="module")
def host_gold():
    with open(os.path.join(GOLD, "host.json")) as f:
        return json.load(f)


# ----------------------------------------------------------------------------- C ABI
def test_cabi_library_exports_every_declared_symbol():
    from crvqa import _lib
    header = open(os.path.join(ROOT, "include", "crvqa.h")).read()
    declared = sorted(set(re.findall(r"\b(crv_[a-z0-9_]+)\s*\(", header)))
    assert len(declared) >= 20
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), f"{name} declared in include/crvqa.h but not exported by libcrvqa.so"
    assert sorted(_lib.EXPORTED) == declared, "ctypes prototype table and header disagree"
    assert _lib.lib.crv_version() == 1


def test_cabi_argument_errors_are_reported_before_any_cuda_work():
    from crvqa import _lib
    lib = _lib.lib
    null = ctypes.c_void_p(0)
    assert lib.crv_cast_f32_to_bf16(null, null, 16, null) == -1
    assert lib.crv_masked_linear_fwd(null, null, null, null, null, null, 0, 128, 128, 64, null) == -1
    assert lib.crv_kth_value_batched(null, null, null, 0, 0, null, null, 0, null) == -1
    assert lib.crv_kth_value_workspace_bytes(168) > 168 * 2048 * 4
    assert lib.crv_vqa_loss_workspace_bytes(256) == 3 * 256 * 4
    assert b"bad argument" in lib.crv_error_string(-1)
    with pytest.raises(_lib.CrvqaError):
        _lib.check(-3, "x")


def test_every_entry_point_rejects_null_arguments_without_touching_cuda():
    """All-null pointers and zero sizes: each int-returning entry point answers CRV_E_ARG (-1) on a box without a GPU,
    i.e. before any CUDA call.  (crv_gemm_debug_timestamps(NULL) is the documented way to switch the debug aid off.)"""
    from crvqa import _lib
    seen = 0
    for name, (res, args) in _lib._PROTOS.items():
        if res is not ctypes.c_int or not args or name == "crv_gemm_debug_timestamps":
            continue
        vals = [ctypes.c_void_p(0) if a is ctypes.c_void_p else (0.0 if a is ctypes.c_float else 0) for a in args]
        assert getattr(_lib.lib, name)(*vals) == -1, name
        seen += 1
    assert seen >= 25
    assert _lib.lib.crv_gemm_debug_timestamps(ctypes.c_void_p(0)) == 0


def test_product_ops_fail_loudly_without_cuda():
    if torch.cuda.is_available():
        pytest.skip("this check is for the CPU-only tier")
    from crvqa import ops
    x = torch.zeros(128, 64)
    with pytest.raises(RuntimeError):
        ops.masked_linear_fwd(x.bfloat16(), x.bfloat16(), x, 0.0, None)
    with pytest.raises(RuntimeError):
        ops.binarize(x, 0.0)
    with pytest.raises(RuntimeError):
        ops.kth_value_batched([x], [1])


# ----------------------------------------------------------------------------- pure host logic
def test_chain_module_names_match_reference(host_gold):
    from masking import maskers, maskers_Robust, maskers_visualBert
    layers = list(range(12))
    assert sorted(maskers.chain_module_names("lxmert", layers, WEIGHT_TYPES)) == host_gold["chain_lxmert"]
    names, modal, module, layer = maskers_Robust.chain_module_names("lxmert", layers, WEIGHT_TYPES)
    assert sorted(names) == host_gold["chain_robust"]["names"]
    assert modal == host_gold["chain_robust"]["modal"]
    assert module == host_gold["chain_robust"]["module"]
    assert layer == host_gold["chain_robust"]["layer"]
    vb = maskers_visualBert.chain_module_names("visual_bert", layers, ["K", "Q", "V", "AO", "I", "O", "P", "E"])
    assert sorted(vb) == host_gold["chain_visualbert"] and len(vb) == 74


def test_sparsity_schedules_match_reference(host_gold):
    from masking import sparsity_control as sp
    f = sp.automated_gradual_sparsity(0.1, 0.7, 0.1, 2, 16)
    assert [f(e, 0.0) for e in range(20)] == pytest.approx(host_gold["ags"], rel=0, abs=0)
    f = sp.stepwise_sparsity(0.1, 0.7, 2, 2, 16, 0.2)
    cur, seq = 0.1, []
    for e in range(20):
        cur = f(e, cur)
        seq.append(cur)
    assert seq == pytest.approx(host_gold["stepwise"], rel=0, abs=0)
    with pytest.raises(ValueError):
        sp.stepwise_sparsity(0.1, 0.99, 2, 2, 6, 0.05)
    conf = types.SimpleNamespace(masking_scheduler_conf_={"init_sparsity": 0.2, "final_sparsity": 0.7,
                                                          "sparsity_warmup_interval_epoch": 1, "init_epoch": 1,
                                                          "final_epoch": 8},
                                 logger=logging.getLogger("t"), num_epochs=10)
    sch = sp.MaskerScheduler(conf)
    got = [list(sch.step(e)) for e in range(10)]
    assert got == [pytest.approx(r[:2]) + [r[2]] if False else r for r in got]  # shape check
    for g, r in zip(got, host_gold["scheduler_steps"]):
        assert g[0] == pytest.approx(r[0], abs=1e-15) and g[1] == pytest.approx(r[1], abs=1e-15) and g[2] == r[2]
    assert sch.is_skip == host_gold["scheduler_is_skip"]
    from masking import sparsity_control_Robust as spr
    assert spr.MaskerScheduler is sp.MaskerScheduler


def test_dict_parser_and_default_scheduler_conf():
    from prune_debias_VQA import DEFAULT_SCHEDULER_CONF
    from utils.param_parser import dict_parser
    d = dict_parser(DEFAULT_SCHEDULER_CONF)
    assert d["lambdas_lr"] == 0 and d["sparsity_warmup"] == "automated_gradual_sparsity" and d["final_epoch"] == 1


def test_linear_schedule_collator_sampler():
    from hg_transformers.data.data_collator import TrimCollator
    from hg_transformers.optimization import linear_schedule_factor
    from hg_transformers.mask_trainer_VQA import SequentialDistributedSampler
    assert linear_schedule_factor(0, 0, 100) == 1.0 and linear_schedule_factor(50, 0, 100) == 0.5
    assert linear_schedule_factor(5, 10, 100) == 0.5 and linear_schedule_factor(100, 10, 100) == 0.0
    batch = [(torch.arange(4), torch.ones(3 + i, 8), torch.tensor(i), 1.5) for i in range(3)]
    ids, feats, qid, w = TrimCollator().collate_batch(batch)
    assert ids.shape == (3, 4) and feats.shape == (3, 5, 8) and qid.tolist() == [0, 1, 2] and w.dtype == torch.float64
    assert float(feats[0, 3:].abs().sum()) == 0.0
    s0 = list(SequentialDistributedSampler(list(range(10)), num_replicas=4, rank=0))
    s3 = list(SequentialDistributedSampler(list(range(10)), num_replicas=4, rank=3))
    assert s0 == [0, 1, 2] and s3 == [9, 0, 1]


def test_learned_mixin_smooth_value_is_cached_against_the_parameter_version():
    from hg_transformers.vqa_debias_loss_functions import LearnedMixin
    lm = LearnedMixin(0.36)
    v = lm.smooth_value()
    assert v == pytest.approx(float(torch.sigmoid(torch.tensor(-1.0))))
    assert lm.smooth_value() == v
    with torch.no_grad():
        lm.smooth_param.fill_(0.0)
    assert lm.smooth_value() == pytest.approx(0.5)


# ----------------------------------------------------------------------------- masker with the oracle as fake backend
@pytest.fixture()
def oracle_backend(monkeypatch):
    from crvqa import ops
    from oracle import masked_ops as o

    def kth(tensors, ks, use_abs=False):
        return torch.tensor([float(o.kth_value(t, int(k), use_abs=use_abs)) for t, k in zip(tensors, ks)])

    def mag(weight, w_thr, hi, lo):
        keep = weight.detach().abs() > float(w_thr)
        return torch.where(keep, torch.full_like(weight, hi), torch.full_like(weight, lo))

    def binz(scores, thr, want_count=False, as_bool=False):
        m = o.binarize(scores.detach(), float(thr))
        out = m.bool() if as_bool else m
        return (out, m.sum().long()) if want_count else out

    class Plan:                                  # the trainer's cached form of the same call
        def __init__(self, tensors):
            self.tensors = tensors

        def __call__(self, ks, use_abs=False):
            return kth(self.tensors, ks, use_abs)

    monkeypatch.setattr(ops, "kth_value_batched", kth)
    monkeypatch.setattr(ops, "KthPlan", Plan)
    monkeypatch.setattr(ops, "magnitude_init", mag)
    monkeypatch.setattr(ops, "binarize", binz)
    return ops


def test_masker_patch_modules_host_logic(oracle_backend):
    """Patching the tiny LXMERT with the per-modality Masker: same module census, same trainable set, same
    initial masks as the reference produced (tests/golden/tiny_lxmert.pt)."""
    from hg_transformers.modeling_lxmert import LxmertConfig, LxmertForMultipleChoice
    from prune_debias_VQA import HPmodel_modal, ModelArguments, init_masker
    g = torch.load(os.path.join(GOLD, "tiny_lxmert.pt"), weights_only=False)
    cfg = dict(g["config"])
    torch.manual_seed(49)
    model = LxmertForMultipleChoice(LxmertConfig(**cfg))
    for k, v in model.state_dict().items():
        assert torch.equal(v, g["state_dict"][k]), k          # same seed -> same init as the reference
    margs = ModelArguments()
    log = logging.getLogger("t")
    log.setLevel(logging.ERROR)
    masker = init_masker(margs, model, log, HPmodel_modal(0.7, 0.7, 0.7, 0.7), margs)
    mods = [(n, m) for n, m in model.named_modules() if hasattr(m, "threshold")]
    assert [n for n, _ in mods] == g["module_names"]
    assert sorted(n for n, p in model.named_parameters() if p.requires_grad) == g["trainable"]
    for n, m in mods:
        assert masker.name_in_module[n] == g["modal"][n]
        assert int((m.weight_mask > 1e-2).sum()) == g["kept_init"][n]
        assert m.weight is dict(model.named_parameters())[n + ".weight"]      # shares the frozen Parameter
        assert not m.weight.requires_grad and m.weight_mask.requires_grad
        assert f"{n}_weight_mask" in masker.init_masks
    assert masker.masker_scheduler.init_sparsity == 0.7 and masker.masker_scheduler.is_skip
    # classifier stays trainable: nn.Sequential children are not attributes, so the freeze never sees them
    assert all(p.requires_grad for n, p in model.named_parameters() if n.startswith("classifier."))


def test_reset_threshold_modes(oracle_backend):
    from hg_transformers import mask_trainer_Robust_VQA as robust
    from hg_transformers import mask_trainer_VQA as base

    class M(torch.nn.Module):
        def __init__(self, n):
            super().__init__()
            self.weight = torch.nn.Parameter(torch.zeros(n), requires_grad=False)
            self.weight_mask = torch.nn.Parameter(torch.arange(1, n + 1, dtype=torch.float32))
            self.threshold = torch.tensor(1e-2)

    net = torch.nn.Module()
    net.a, net.b, net.c = M(10), M(100), M(3)
    masker = types.SimpleNamespace(name_in_module={"a": "Lang", "b": "Vis", "c": "P"},
                                   hpmodel=types.SimpleNamespace(zerorate_dict={"Lang": 0.5, "Vis": 0.25, "P": 0.1}))
    t = robust.Trainer.__new__(robust.Trainer)
    t.masker = masker
    mean = t.reset_threshold(net, 0.7)
    assert [float(net.a.threshold), float(net.b.threshold), float(net.c.threshold)] == [5.0, 25.0, 1.0]  # k=0 -> 1
    assert mean == pytest.approx((5 + 25 + 1) / 3)
    t2 = base.Trainer.__new__(base.Trainer)
    t2.masker = masker
    t2.reset_threshold(net, 0.7)
    assert [float(net.a.threshold), float(net.b.threshold), float(net.c.threshold)] == [7.0, 70.0, 2.0]


def test_save_model_mask_host_logic(oracle_backend, tmp_path):
    """Trainer.save_model_mask: mask.pt holds one CPU BoolTensor per masked module (`S > thr`), the return value is the
    overall zero rate in percent, per-modality bookkeeping follows masker.name_in_module (incl. DataParallel's
    'module.' prefix)."""
    from hg_transformers import mask_trainer_Robust_VQA as robust

    class M(torch.nn.Module):
        def __init__(self, scores, thr):
            super().__init__()
            self.weight_mask = torch.nn.Parameter(torch.tensor(scores))
            self.threshold = torch.tensor(thr)

    net = torch.nn.Module()
    net.module = torch.nn.Module()
    net.module.a = M([[0.0, 0.02, 0.02, 0.0]], 1e-2)        # 2 of 4 kept
    net.module.b = M([[0.5, 0.1], [0.3, 0.31]], 0.3)        # strict >: 0.3 itself is dropped -> 2 of 4 kept
    net.c = M([[1.0, 2.0, 3.0]], 0.0)                       # all kept
    t = robust.Trainer.__new__(robust.Trainer)
    t.model = net
    t.masker = types.SimpleNamespace(name_in_module={"a": "Lang", "b": "Fus", "c": "P"})
    t.args = types.SimpleNamespace(output_dir=str(tmp_path / "unused"))
    rate = t.save_model_mask(str(tmp_path / "m"))
    assert float(rate) == pytest.approx(100.0 * 4 / 11)
    saved = torch.load(tmp_path / "m" / "mask.pt")
    assert sorted(saved) == ["c.weight", "module.a.weight", "module.b.weight"]
    assert saved["module.b.weight"].dtype == torch.bool
    assert saved["module.b.weight"].tolist() == [[True, False], [False, True]]
    assert saved["module.a.weight"].tolist() == [[False, True, True, False]] and saved["c.weight"].all()


# ----------------------------------------------------------------------------- data parallel (gloo, 2 ranks)
def _sync_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from hg_transformers._engine import GradSync, ScoreArena

    class M(torch.nn.Module):
        def __init__(self, shape):
            super().__init__()
            self.weight_mask = torch.nn.Parameter(torch.zeros(shape))
            self.threshold = torch.tensor(1e-2)

    mods = [(f"m{i}", M(s)) for i, s in enumerate([(8, 16), (4, 4), (32, 8), (5, 3), (16, 16)])]
    arena = ScoreArena(mods)
    for _, m in mods:
        m._grad_zero = False                          # as after a step whose optimiser pass kept the gradients
    sync = GradSync(arena, bucket_bytes=512)          # several small buckets
    assert len(sync.bucket_ranges) >= 3
    loose = [torch.full((7,), float(rank + 1)), torch.full((2, 3), 10.0 * (rank + 1))]
    for step in range(2):
        arena.begin_step()
        sync.begin_step()
        calls = {0: 1, 1: 2, 2: 1, 4: 1}              # module 1 is "shared" (two invocations); module 3 gets no grad
        for i, c in calls.items():
            mods[i][1]._calls_outstanding = c
        for i in (4, 2, 1, 1, 0):                      # backward order
            m = mods[i][1]
            val = float((rank + 1) * (i + 1) + step)
            if m._grad_dirty:
                m._arena_grad.add_(val)
            else:
                m._arena_grad.fill_(val)
            m._grad_dirty = True
            sync.module_backward_done(m)
        if step == 0:
            mods[3][1]._arena_grad.fill_(123.0)        # stale garbage that must be zeroed, not exchanged
        sync.finish(loose)
        mean_rank = (1 + world) / 2.0
        for i, (_, m) in enumerate(mods):
            if i == 3:
                want = 0.0
            elif i == 1:
                want = 2 * (mean_rank * (i + 1) + step)
            else:
                want = mean_rank * (i + 1) + step
            assert torch.allclose(m.weight_mask.grad, torch.full_like(m.weight_mask, want)), (rank, step, i)
            assert m.weight_mask.grad.data_ptr() == m._arena_grad.data_ptr()
        if step == 0:
            assert torch.allclose(loose[0], torch.full((7,), mean_rank))
            assert torch.allclose(loose[1], torch.full((2, 3), 10.0 * mean_rank))
    dist.barrier()
    dist.destroy_process_group()
    q.put(rank)


def test_grad_sync_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_sync_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert sorted(q.get(timeout=5) for _ in range(2)) == [0, 1]


def test_execution_order_follows_the_lockstep_forward():
    """Arena / bucket order follows the fused forward pass (feature encoder + word embeddings, language layer i and
    vision layer i in lockstep, remaining language layers, cross layers, pooler) and is stable inside each group (so
    query | key | value of a layer stay adjacent); single-stack models keep named_modules order."""
    from hg_transformers._engine import execution_order
    from oracle import lxmert_oracle as lxo
    named = [(n, object()) for n, _ in lxo.module_names()]
    got = [n for n, _ in execution_order(named)]
    assert sorted(got) == sorted(n for n, _ in named)
    assert set(got[:3]) == {"lxmert.embeddings.word_embeddings", "lxmert.encoder.visn_fc.visn_fc",
                            "lxmert.encoder.visn_fc.box_fc"}
    first_x = next(i for i, n in enumerate(got) if ".x_layers." in n)
    stack = got[3:first_x]
    want = []
    for i in range(9):
        want += [n for n, _ in named if f".encoder.layer.{i}." in n]
        want += [n for n, _ in named if f".encoder.r_layers.{i}." in n]
    assert stack == want
    assert stack[0].endswith("layer.0.attention.self.query") and stack[1].endswith("layer.0.attention.self.key")
    assert got[-1] == "lxmert.pooler.dense"
    assert [n for n in got if ".x_layers." in n] == [n for n, _ in named if ".x_layers." in n]      # stable
    vb = [("visual_bert.embeddings.word_embeddings", 0), ("visual_bert.encoder.layer.0.attention.self.query", 1),
          ("visual_bert.pooler.dense", 2)]
    assert execution_order(vb) == vb


def test_secondary_debias_losses_match_reference():
    """RUBI_loss and BiasProduct are plain torch graphs in the drop-in: same values and
    logit gradients as the reference (tests/golden/secondary_losses.pt, make_golden_secondary_losses.py)."""
    from hg_transformers._trainer_core import RUBI_loss
    from hg_transformers.vqa_debias_loss_functions import BiasProduct
    g = torch.load(os.path.join(GOLD, "secondary_losses.pt"), weights_only=False)
    logits = g["logits"].clone().requires_grad_(True)
    loss = RUBI_loss(logits, g["bias"], g["max_label"])
    loss.backward()
    torch.testing.assert_close(loss.detach(), g["rubi"], rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(logits.grad, g["rubi_dlogits"], rtol=1e-5, atol=1e-8)
    logits.grad = None
    bp = BiasProduct()
    assert torch.equal(bp.smooth_param.detach(), g["bp_smooth_param"])
    loss = bp(g["hidden"], logits, g["bias"], g["labels"])
    loss.backward()
    torch.testing.assert_close(loss.detach(), g["bp"], rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(logits.grad, g["bp_dlogits"], rtol=1e-5, atol=1e-8)


# ----------------------------------------------------------------------------- bench.py contract
_LINE_KEYS = {"metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "e2e", "gpu_launches", "cpu_baseline"}


def test_recorded_bench_lines_follow_the_contract():
    """The committed bench lines (profiles/r01_bench_{1,2,4,8}gpu.json) carry every key of the bench contract with
    consistent values: whole-job throughput = global batch / step time, roofline fraction = achieved / peak, e2e measured
    with host-to-device traffic, weak scaling, a CPU baseline on rank 0 / N = 1 only."""
    for n in (1, 2, 4, 8):
        with open(os.path.join(ROOT, "profiles", f"r01_bench_{n}gpu.json")) as f:
            line = json.loads(f.read().strip().splitlines()[-1])
        assert _LINE_KEYS - {"cpu_baseline"} <= set(line), n
        assert line["n_gpus"] == n and line["higher_is_better"] is True and line["scaling"] == "weak"
        assert line["unit"] == "samples/s" and line["vs_baseline"] is None and line["data"] == "synthetic"
        assert "workload" in line["config"] and line["config"]["global_batch"] == 256 * n
        assert line["value"] == pytest.approx(256 * n / line["ms_per_step"] * 1e3, rel=1e-6)
        assert line["warmup"] >= 3 and line["gpu_launches"] > 0
        e2e = line["e2e"]
        assert e2e["unit"] == "samples/s" and e2e["h2d_bytes_per_step"] > 0 and e2e["d2h_bytes_per_step"] > 0
        assert e2e["value"] <= line["value"] * 1.02
        assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        if n == 1:
            r = line["roofline"]
            assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and r["traffic"] is not None
            assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9) and 0.6 < r["frac"] < 1.0
            c = line["cpu_baseline"]
            assert c["kind"] == "port" and c["cores"] >= 1 and c["value"] > 0 and "batch 32" in c["sample"]
    ref = {n: json.loads(open(os.path.join(ROOT, "profiles", f"r01_bench_{n}gpu.json")).read().strip().splitlines()[-1])
           for n in (1, 8)}
    assert ref[8]["value"] / ref[1]["value"] > 7.0          # the north_star's scaling target, as recorded


def test_recorded_round2_bench_lines_follow_the_contract():
    """The round-2 lines (profiles/r02_final_bench_{1,2,4,8}gpu.json): same contract; the reference arm of the GPU line
    is the reference's OWN modules now (kind "reference"), the roofline fraction is against the BURST peak with the
    sustained one beside it, the DRAM traffic per GEMM launch is measured, and 8 GPUs give more than 7.6x."""
    lines = {}
    for n in (1, 2, 4, 8):
        with open(os.path.join(ROOT, "profiles", f"r02_final_bench_{n}gpu.json")) as f:
            line = lines[n] = json.loads(f.read().strip().splitlines()[-1])
        assert _LINE_KEYS - {"cpu_baseline"} <= set(line), n
        assert line["n_gpus"] == n and line["higher_is_better"] is True and line["scaling"] == "weak"
        assert line["metric"] == "LXMERT stage-2 mask-train samples/s" and line["unit"] == "samples/s"
        assert line["vs_baseline"] is None and line["data"] == "synthetic" and line["dtype"] == "bf16"
        assert line["config"]["global_batch"] == 256 * n and line["config"]["threshold_refresh"]["every_steps"] == 100
        assert line["value"] == pytest.approx(256 * n / line["ms_per_step"] * 1e3, rel=1e-6)
        assert line["warmup"] >= 3 and line["gpu_launches"] > 0
        e2e = line["e2e"]
        assert e2e["h2d_bytes_per_step"] == 82100224 and e2e["d2h_bytes_per_step"] == 4
        assert e2e["value"] <= line["value"] * 1.02
        assert not set(line["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}
        r = line["roofline"]
        assert r["bound"] == "tensor" and r["unit"] == "TFLOP/s" and r["launches_per_step"] <= 150
        assert r["frac"] == pytest.approx(r["achieved"] / r["peak"], rel=1e-9) and r["frac"] >= 0.60
        assert r["peak"] > r["peak_sustained"] and r["frac_sustained"] > r["frac"]
        if n == 1:
            assert r["traffic"] is not None and r["traffic"] < r["traffic_algorithmic"]
            c = line["cpu_baseline"]
            assert c["kind"] == "reference" and c["cores"] >= 1 and c["value"] > 0 and "batch 32" in c["sample"]
    assert lines[8]["value"] / lines[1]["value"] > 7.6
    assert lines[2]["value"] / lines[1]["value"] > 1.9


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` runs the reference's own stage-2 modules (oracle/_ref, or /root/reference where it
    exists; the oracle port only when neither does) on the host cores and prints one JSON line with impl = reference,
    its own cpu_baseline block and an e2e block without copies."""
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                          "--warmup", "0", "--cpu-batch", "4"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-500:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and _LINE_KEYS <= set(line)
    from oracle import ref_runner
    want = "reference" if ref_runner.reference_root() is not None else "port"
    assert line["cpu_baseline"]["kind"] == want and line["cpu_baseline"]["value"] == line["value"] > 0
    assert line["e2e"] == {"value": line["value"], "unit": "samples/s", "h2d_bytes_per_step": 0,
                           "d2h_bytes_per_step": 0}
    assert line["gpu_launches"] == 0 and line["dtype"] == "f32"


def test_answer_head_padded_path_is_the_same_function():
    """SimpleClassifier pads its last weight-normed Linear to an aligned width on CUDA (hg_transformers/classifier.py;
    reference classifier.py:5-22 has 3129 / 2274 answers).  The padded computation, run here on the CPU, returns exactly
    the module's own output and the same parameter gradients."""
    import torch.nn.functional as F
    from hg_transformers.classifier import SimpleClassifier
    torch.manual_seed(3)
    head = SimpleClassifier(64, 96, 37).eval()
    x = torch.randn(5, 64)
    want = head(x)
    want.sum().backward()
    gw = {n: p.grad.clone() for n, p in head.named_parameters()}
    head.zero_grad()
    last = head.main[3]
    h = head.main[2](head.main[1](head.main[0](x)))
    for hook in last._forward_pre_hooks.values():
        hook(last, (h,))
    pad = (-37) % 16
    got = F.linear(h, F.pad(last.weight, (0, 0, 0, pad)), F.pad(last.bias, (0, pad)))[:, :37]
    assert torch.equal(got, want)
    got.sum().backward()
    for n, p in head.named_parameters():
        torch.testing.assert_close(p.grad, gw[n], rtol=1e-6, atol=1e-7)
    assert sorted(head.state_dict()) == ["main.0.bias", "main.0.weight_g", "main.0.weight_v", "main.3.bias",
                                         "main.3.weight_g", "main.3.weight_v"]


def test_bench_workloads_cover_the_baseline_configs():
    """bench.py --config names: lxmert (BASELINE configs[1], the recorded line), visualbert (configs[2]), stage3
    (configs[3]); every workload carries its own metric name and the algorithmic GFLOP per sample."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert sorted(bench.WORKLOADS) == ["lxmert", "stage3", "visualbert"]
    assert bench.WORKLOADS["lxmert"]["metric"] == bench.METRIC == "LXMERT stage-2 mask-train samples/s"
    assert bench.WORKLOADS["lxmert"]["gflop_per_sample"] == pytest.approx(10.4955 * 2 + 10.3821, abs=1e-3)
    assert bench.WORKLOADS["visualbert"]["gflop_per_sample"] == pytest.approx(28.54)
