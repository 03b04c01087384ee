"""mPLUG-VQA network, CPU tier: the drop-in ``mPLUG/models`` (plain torch until the masker patches it) against the
reference's network run in the build container (tests/golden/mplug_model_tiny.pt, make_golden_mplug_model.py): same
state_dict keys (strict load), same loss and per-parameter gradient norms; then the host logic of masking it, with the
oracle as the fake kernel backend."""
import os
import sys
import types

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from test_mplug_cpu import kept, masked, oracle_backend, quiet, thr_record  # noqa: E402,F401

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mplug_model_tiny.pt")
TWINS = ("visual_encoder", "text_encoder", "fusion_encoder", "text_decoder")


@pytest.fixture(scope="module")
def gold():
    return torch.load(GOLD, weights_only=False)


def build(gold, device="cpu"):
    from mPLUG.models.model_vqa_mplug import MPLUG
    config = dict(gold["config"], bert_config=dict(gold["bert"]))
    model = MPLUG(config=config, tokenizer=types.SimpleNamespace(pad_token_id=0))
    assert sorted(model.state_dict()) == gold["state_dict_keys"]
    full = dict(gold["online"])
    for k, v in gold["online"].items():             # the twins start as copies of the online towers
        tower = k.split(".")[0]
        if tower in TWINS:
            full[tower + "_m" + k[len(tower):]] = v
    model.load_state_dict(full, strict=True)
    return model.eval().to(device)                  # dropout off, as in the golden run; gradients still flow


def batch(gold, device="cpu"):
    B, res, V = 4, gold["config"]["image_res"], gold["bert"]["vocab_size"]
    g = torch.Generator().manual_seed(5)
    image = torch.randn(B, 3, res, res, generator=g)
    q_ids = torch.randint(1, V, (B, 7), generator=g)
    q_att = torch.ones(B, 7, dtype=torch.long)
    q_att[1, 5:] = 0
    q_ids[1, 5:] = 0
    k = [2, 1, 3, 2]
    n = sum(k)
    a_ids = torch.randint(1, V, (n, 5), generator=g)
    a_att = torch.ones(n, 5, dtype=torch.long)
    a_ids[0, 3:] = 0
    a_att[0, 3:] = 0
    a_ids[5, 4:] = 0
    a_att[5, 4:] = 0
    weights = torch.rand(n, generator=g) + 0.2
    bias = torch.rand(n, generator=g) * 0.5
    d = torch.device(device)
    question = types.SimpleNamespace(input_ids=q_ids.to(d), attention_mask=q_att.to(d))
    answer = types.SimpleNamespace(input_ids=a_ids.to(d), attention_mask=a_att.to(d))
    return image.to(d), question, answer, k, weights.to(d), bias.to(d)


def run(model, gold, with_bias, device="cpu"):
    image, question, answer, k, weights, bias = batch(gold, device)
    for p in model.parameters():
        p.grad = None
    loss = model(image, question, answer, train=True, alpha=0.4, k=k, weights=weights, bias=bias if with_bias else None)
    loss.backward()
    return float(loss.detach()), {n: float(p.grad.norm()) for n, p in model.named_parameters() if p.grad is not None}


def test_dense_network_matches_reference(gold):
    model = build(gold)
    assert (model.text_decoder.cls.predictions.decoder.weight
            is model.text_decoder.bert.embeddings.word_embeddings.weight) == gold["tied"]
    twin0 = {k: v.clone() for k, v in model.state_dict().items() if k.startswith("text_encoder_m.")}
    loss, norms = run(model, gold, with_bias=False)
    assert loss == pytest.approx(gold["dense_loss"], rel=1e-5)
    assert sorted(norms) == sorted(gold["dense_grad_norms"])
    for n, want in gold["dense_grad_norms"].items():
        assert norms[n] == pytest.approx(want, rel=2e-4, abs=1e-9), n
    loss_b, _ = run(model, gold, with_bias=True)
    assert loss_b == pytest.approx(gold["dense_loss_bias"], rel=1e-5)
    moved = any(not torch.equal(v, twin0[k]) for k, v in model.state_dict().items() if k in twin0)
    assert moved == gold["twin_moved"]              # the momentum update runs although its logits are never weighted in
    # the skipped twin forward changes nothing: with it switched on the loss is identical
    type(model).run_unused_distill_forward = True
    try:
        again, _ = run(model, gold, with_bias=True)
    finally:
        type(model).run_unused_distill_forward = False
    assert again == pytest.approx(loss_b, rel=1e-6)


def test_rank_answer_matches_reference(gold):
    from mPLUG.models.model_vqa_mplug import tile
    model = build(gold)
    image, question, answer, k, weights, bias = batch(gold)
    states, atts = model.encode_question(image, question)
    R = gold["rank"]
    assert float(states.norm()) == pytest.approx(R["states_norm"], rel=1e-5)
    cand = answer.input_ids.clone()
    cand[:, 0] = R["bos"]
    with torch.no_grad():
        ids, probs = model.rank_answer(states, atts, cand, answer.attention_mask, R["k"])
    assert torch.equal(ids, R["topk_ids"])
    assert torch.allclose(probs, R["topk_probs"], rtol=1e-4, atol=1e-7)
    assert torch.equal(tile(torch.arange(6).view(2, 3), 0, 3), gold["tile"])


def test_full_size_configuration_builds_the_reference_census():
    """mPLUG-base at 384 px on the meta device: parameter count and the census of maskable modules."""
    from mPLUG import vqa_mplug
    from mPLUG.masking.mask_config import MaskConfigs
    from mPLUG.models.model_vqa_mplug import MPLUG
    config = dict(image_res=384, vision_width=768, distill=True, clip_name="ViT-B-16",
                  bert_config=dict(stride_layer=3, fusion_layers=6, text_encoder_layers=6, text_decode_layers=12))
    with torch.device("meta"):
        model = MPLUG(config=config)
    names = vqa_mplug.names_to_mask(MaskConfigs())
    found = [n for n, m in model.named_modules() if n in names]
    assert len(found) == len(names) == 480          # every name the reference tables produce exists in the network
    assert all(isinstance(dict(model.named_modules())[n], torch.nn.Linear) for n in found)
    assert model.visual_encoder.visual.positional_embedding.shape == (577, 768)
    online = sum(p.numel() for n, p in model.named_parameters() if "_m." not in n)
    assert 420e6 < online < 435e6                   # ViT-B/16 86 M, text 6 layers, fusion and decoder 12 each, 3 tables


def test_masking_the_network_host_logic(gold, oracle_backend):
    from mPLUG import vqa_mplug
    from mPLUG.masking.mask_config import MaskConfigs
    from mPLUG.masking import maskers
    G = gold["masked"]
    model = build(gold)
    conf = MaskConfigs()
    conf.zero_rate = 0.5
    quiet(vqa_mplug.init_masker, conf, model, layers_to_mask=gold["layers_to_mask"])
    assert [n for n, _ in masked(model)] == G["module_names"]
    assert sorted(n for n, p in model.named_parameters() if p.requires_grad) == G["trainable"]
    assert thr_record(model) == G["thresholds"]
    assert kept(model) == G["kept"]
    mean = maskers.reset_threshold(model, 0.7)
    r = G["reset_0.7"]
    assert mean == r["mean"] and thr_record(model) == r["thresholds"] and kept(model) == r["kept"]


def test_beam_search_generation_matches_reference():
    """MPLUG.forward(train=False) -> predictor.TextGenerator: same token sequences and scores as the reference's beam
    search for three (beam, min_length, max_length) settings, the full ranked lists, and the early-close case where
    finished questions leave the batch (tests/golden/mplug_generation_tiny.pt, make_golden_mplug_generation.py)."""
    from mPLUG.models.model_vqa_mplug import MPLUG
    g = torch.load(os.path.join(os.path.dirname(GOLD), "mplug_generation_tiny.pt"), weights_only=False)
    config = dict(g["config"], bert_config=dict(g["bert"]))
    model = MPLUG(config=config, tokenizer=types.SimpleNamespace(pad_token_id=0))
    model.load_state_dict(g["state_dict"], strict=True)
    model.eval()
    B, res = 5, g["config"]["image_res"]
    gen = torch.Generator().manual_seed(9)
    image = torch.randn(B, 3, res, res, generator=gen)
    q_ids = torch.randint(1, 100, (B, 6), generator=gen)
    q_att = torch.ones(B, 6, dtype=torch.long)
    q_att[2, 4:] = 0
    q_ids[2, 4:] = 0
    image = image * torch.arange(1, B + 1).view(-1, 1, 1, 1).float()
    question = types.SimpleNamespace(input_ids=q_ids, attention_mask=q_att)

    def same(ids, scores, want):
        assert [[t.tolist() for t in q] for q in ids] == [[t.tolist() for t in q] for q in want["ids"]]
        for got_q, want_q in zip(scores, want["scores"]):
            assert [float(s) for s in got_q] == pytest.approx(want_q, rel=1e-4)

    bg = model.beam_generator
    for key in ((3, 1, 6), (2, 0, 4), (1, 2, 5)):
        bg.beam_size, bg.min_length, bg.max_length = key
        ids, scores = model(image, question, None, train=False, k=None)
        same(ids, scores, g["runs"][key])
    bg.beam_size, bg.min_length, bg.max_length = 3, 1, 6
    states, atts = model.encode_question(image, question)
    same(*bg.translate_batch([states, atts], out_size=3), g["runs"]["ranked3"])
    with torch.no_grad():
        model.text_decoder.cls.predictions.bias[102] = g["runs"]["early_close_bias"]
    ids, scores = bg.translate_batch([states, atts], out_size=3)
    same(ids, scores, g["runs"]["early_close"])
    assert sorted({len(t) for q in ids for t in q}) != [7]          # some hypotheses ended before max_length
    with pytest.raises(NotImplementedError):
        bg.translate_batch([states, atts], do_sample=True)


def test_evaluation_loop_with_a_toy_tokenizer(tmp_path):
    """vqa_mplug.evaluation / evaluate / cal_metric / save_result on the generation fixture: answers are the decoded best
    hypotheses without the special tokens; the accuracy is the mean soft score from the label file."""
    import json

    from mPLUG import vqa_mplug
    from mPLUG.models.model_vqa_mplug import MPLUG
    g = torch.load(os.path.join(os.path.dirname(GOLD), "mplug_generation_tiny.pt"), weights_only=False)
    model = MPLUG(config=dict(g["config"], bert_config=dict(g["bert"])), tokenizer=types.SimpleNamespace(pad_token_id=0))
    model.load_state_dict(g["state_dict"], strict=True)
    bg = model.beam_generator
    bg.beam_size, bg.min_length, bg.max_length = 3, 1, 6

    class Enc(types.SimpleNamespace):
        def to(self, device):
            return Enc(input_ids=self.input_ids.to(device), attention_mask=self.attention_mask.to(device))

    class Tok:
        special = {101: "[CLS]", 102: "[SEP]", 0: "[PAD]"}

        def __call__(self, questions, padding="longest", return_tensors="pt", **kw):
            ids = torch.stack(questions)
            return Enc(input_ids=ids, attention_mask=(ids != 0).long())

        def decode(self, ids):
            return " ".join(self.special.get(int(t), f"w{int(t)}") for t in ids)

    B, res = 5, g["config"]["image_res"]
    gen = torch.Generator().manual_seed(9)
    image = torch.randn(B, 3, res, res, generator=gen)
    q_ids = torch.randint(1, 100, (B, 6), generator=gen)
    q_ids[2, 4:] = 0
    image = image * torch.arange(1, B + 1).view(-1, 1, 1, 1).float()
    loader = [(image[:3], list(q_ids[:3]), torch.tensor([10, 11, 12])), (image[3:], list(q_ids[3:]), torch.tensor([13, 14]))]
    out = vqa_mplug.evaluation(model, loader, Tok(), torch.device("cpu"), {"k_test": 128})
    want = g["runs"][(3, 1, 6)]["ids"]
    assert [r["question_id"] for r in out] == [10, 11, 12, 13, 14]
    for r, q in zip(out, want):
        assert r["answer"] == " ".join(f"w{int(t)}" for t in q[0] if int(t) not in (0, 101, 102))
    labels = [{"question_id": r["question_id"], "label": {r["answer"]: 0.5 + 0.1 * i}} for i, r in enumerate(out)]
    labels[1]["label"] = {"something else": 1.0}
    (tmp_path / "labels.json").write_text(json.dumps(labels))
    assert vqa_mplug.cal_metric(out, [str(tmp_path / "labels.json")]) == pytest.approx((0.5 + 0.7 + 0.8 + 0.9) / 5)
    stats = vqa_mplug.evaluate(model, loader, [str(tmp_path / "labels.json")], Tok(), torch.device("cpu"),
                               {"k_test": 128}, str(tmp_path / "out"))
    assert stats == {"acc": "{:.4f}".format((0.5 + 0.7 + 0.8 + 0.9) / 5)}
    assert json.loads((tmp_path / "out" / "vqa_answer.json").read_text()) == out


class _ToyTokenizer:
    """Whitespace tokenizer over a tiny vocabulary with BERT's special ids ([PAD] 0, [CLS] 101 -> here 1, [SEP] 2)."""
    pad_token_id = 0

    def __init__(self, words):
        self.ids = {w: i + 3 for i, w in enumerate(words)}

    def __call__(self, texts, padding="longest", truncation=False, max_length=None, return_tensors="pt"):
        rows = []
        for t in texts:
            toks = [1] + [self.ids[w] for w in t.replace("[SEP]", " ").split()] + [2]
            rows.append(toks[:max_length] if truncation and max_length else toks)
        width = max(len(r) for r in rows)
        ids = torch.tensor([r + [0] * (width - len(r)) for r in rows])

        class Enc(types.SimpleNamespace):
            def to(self, device):
                return Enc(input_ids=self.input_ids.to(device), attention_mask=self.attention_mask.to(device))

        return Enc(input_ids=ids, attention_mask=(ids != 0).long())


def test_reference_signature_train_loop_on_cpu(gold):
    """vqa_mplug.train with the reference's argument list on the dense tiny network (plain torch on the CPU): collate ->
    tokenise -> engine step, alpha ramp and scheduler warm-up calls in epoch 0, stats dict; train_pretokenized too."""
    from torch.utils.data import DataLoader

    from mPLUG import vqa_mplug
    from mPLUG.dataset import SyntheticVQAImageDataset, vqa_bias_collate_fn, vqa_collate_fn
    from mPLUG.engine import MaskTrainEngine
    from mPLUG.optim import create_two_optimizer
    from mPLUG.scheduler import create_scheduler
    model = build(gold)
    words = ("what", "color", "is", "the", "cat", "two", "red", "yes", "no", "dog")
    tok = _ToyTokenizer(words)
    data = SyntheticVQAImageDataset(6, image_res=gold["config"]["image_res"], words=words, with_bias=True, seed=3)
    item = data[0]
    assert item[0].shape == (3, 64, 64) and isinstance(item[1], str) and len(item[2]) == len(item[3]) == len(item[4])
    loader = DataLoader(data, batch_size=3, collate_fn=vqa_bias_collate_fn)
    image, questions, answers, weights, n, bias = next(iter(loader))
    assert image.shape[0] == 3 and len(questions) == 3 and sum(n) == len(answers) == weights.numel() == bias.numel()
    plain = vqa_collate_fn([d[:4] for d in (data[0], data[1])])
    assert len(plain) == 5 and plain[4] == [len(data[0][2]), len(data[1][2])]

    opt = create_two_optimizer(types.SimpleNamespace(lr1=1e-3, lr2=1e-4, weight_decay=0.02), model)
    sch, _ = create_scheduler(types.SimpleNamespace(sched="cosine", lr=1e-3, epochs=8, min_lr=1e-6, decay_rate=1,
                                                    warmup_lr=1e-5, warmup_epochs=4, cooldown_epochs=0), opt)
    eng = MaskTrainEngine(model, opt, gradient_clipping=1.0, bf16=False)
    alphas = []
    fwd = model.forward

    def spy(*a, **k):
        alphas.append(k["alpha"])
        return fwd(*a, **k)

    model.forward = spy
    cfg = {"alpha": 0.4, "warm_up": True, "add_ocr": False}
    s0 = vqa_mplug.train(eng, loader, opt, tok, 0, 4, torch.device("cpu"), sch, cfg, do_two_optim=True)
    assert alphas == [0.0, 0.2] and eng.global_steps == 2                    # alpha * min(1, i / len(loader))
    assert set(s0) == {"loss", "lr1", "lr2"}
    assert opt.param_groups[0]["lr"] == pytest.approx(1e-5)                    # scheduler.step(0) at i == 0: warm-up start
    s1 = vqa_mplug.train(eng, loader, opt, tok, 1, 4, torch.device("cpu"), sch, cfg, do_two_optim=True)
    assert alphas[2:] == [0.4, 0.4] and eng.global_steps == 4
    assert float(s1["loss"]) > 0 and float(s1["lr1"]) == 0.0                 # "{:.3f}" of 1e-5, as the reference logs it
    model.forward = fwd
    q = tok(questions)
    a = tok(answers)
    mean = vqa_mplug.train_pretokenized(eng, [(image, q, a, 0.4, n, weights)] * 2, 2)
    assert eng.global_steps == 6 and mean > 0
    assert not torch.equal(model.text_decoder.cls.predictions.transform.dense.weight,
                           gold["online"]["text_decoder.cls.predictions.transform.dense.weight"])


def test_synthetic_main_runs_dense_on_cpu(capsys):
    """python mPLUG/vqa_mplug.py --tiny --no_mask: the whole driver flow on the CPU (dense, no CUDA ops)."""
    from mPLUG import vqa_mplug
    from mPLUG.dataset import WhitespaceTokenizer
    stats = vqa_mplug.main(["--tiny", "--no_mask", "--image_res", "32", "--batch_size", "2", "--steps", "2",
                            "--device", "cpu"])
    assert set(stats) == {"loss", "lr1", "lr2"} and float(stats["loss"]) > 0
    tok = WhitespaceTokenizer()
    enc = tok(["what color is the cat", "yes"])
    assert enc.input_ids[1].tolist() == [101, 110, 102, 0, 0, 0, 0] and enc.attention_mask[1].tolist() == [1, 1, 1, 0, 0, 0, 0]
    assert tok.decode(enc.input_ids[1]) == "[CLS] yes [SEP] [PAD] [PAD] [PAD] [PAD]"


def test_momentum_update_drops_the_twins_operand_caches(gold, oracle_backend):
    """The EMA writes the twins' weights through .data; their masked modules must not keep bf16 operands derived from
    the old values (matters only when the optional twin forward is switched on)."""
    from mPLUG import vqa_mplug
    from mPLUG.masking.mask_config import MaskConfigs
    model = build(gold)
    conf = MaskConfigs()
    conf.zero_rate = 0.5
    quiet(vqa_mplug.init_masker, conf, model, layers_to_mask=gold["layers_to_mask"])
    twin = model.text_encoder_m.encoder.layer[0].intermediate.dense
    online = model.text_encoder.encoder.layer[0].intermediate.dense
    twin._w16, twin._wm = torch.zeros(1), torch.zeros(1)
    online._w16 = torch.zeros(1)
    with torch.no_grad():
        online.weight_mask.add_(1.0)
    before = twin.weight_mask.detach().clone()
    model._momentum_update()
    assert twin._w16 is None and twin._wm is None and online._w16 is not None
    assert torch.allclose(twin.weight_mask, before * 0.995 + online.weight_mask * 0.005)


def test_batch_first_vit_attention_equals_the_sequence_first_module(monkeypatch):
    """The ViT tower runs batch-major by default (heads as strided views of the packed projection,
    ResidualAttentionBlock._self_attention_batch_first); with CRVQA_MPLUG_FUSED=0 it goes through nn.MultiheadAttention
    on [tokens, batch, width] as the reference does (mPLUG/models/clip/model.py:157-249).  Same function: outputs and
    every parameter gradient agree to fp32 rounding; the module tree / state_dict is untouched."""
    from mPLUG.models.clip.model import VisualTransformer
    torch.manual_seed(3)
    vit = VisualTransformer(input_resolution=32, patch_size=8, width=64, layers=3, heads=4, output_dim=32).eval()
    img = torch.randn(3, 3, 32, 32)
    params = [p for p in vit.parameters()]

    def run():
        out = vit(img, skip_last_layer=True)
        grads = torch.autograd.grad((out * torch.linspace(-1, 1, out.numel()).view_as(out)).sum(), params,
                                    allow_unused=True)
        return out.detach(), grads

    assert vit.transformer.batch_first_ok()
    out_fast, g_fast = run()
    monkeypatch.setenv("CRVQA_MPLUG_FUSED", "0")
    assert not vit.transformer.batch_first_ok()
    out_ref, g_ref = run()
    assert out_fast.shape == out_ref.shape == (3, 17, 64)
    assert torch.allclose(out_fast, out_ref, rtol=1e-5, atol=1e-5)
    for p, a, b in zip(params, g_fast, g_ref):
        assert (a is None) == (b is None)
        if a is not None:
            assert float((a - b).abs().max()) <= 1e-5 * float(b.abs().max()) + 1e-7
    # a tower with an attention mask keeps the sequence-first module
    monkeypatch.setenv("CRVQA_MPLUG_FUSED", "1")
    assert not vit.transformer.batch_first_ok(text_mask=torch.zeros(17, 17))
    assert "transformer.resblocks.0.attn.in_proj_weight" in vit.state_dict()
