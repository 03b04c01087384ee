"""Golden vectors for the mPLUG-VQA network (BASELINE config 5), from the UNMODIFIED reference classes
mPLUG/models/{model_vqa_mplug,modeling_mplug}.py, mPLUG/models/clip/model.py and mPLUG/masking/maskers.py run on the CPU
at a miniature size:

    python tests/golden/make_golden_mplug_model.py        # writes tests/golden/mplug_model_tiny.pt

Shims (all outside the reference tree; the reference pins transformers==4.14.1, this image has 5.x):
  * ``apply_chunking_to_forward`` / ``prune_linear_layer`` moved to transformers.pytorch_utils; ``get_head_mask`` is gone
    (the path passes head_mask=None) -> re-attached to the names the reference imports / calls;
  * ``PreTrainedModel.init_weights`` of 5.x needs ``post_init`` bookkeeping -> the 4.x behaviour: apply ``_init_weights``
    and tie the LM head's decoder to the word embeddings;
  * ``invert_attention_mask`` of 5.x multiplies by finfo.min instead of 4.x's -10000 -> 4.x form (identical softmax
    weights in fp32: exp of either underflows to exactly 0);
  * ``timm`` and ``ftfy`` are absent -> stub modules (``models.vit`` is imported but never used on this path);
  * checkpoints are not shipped -> ``from_pretrained`` builds from the config, ``initialize_clip`` builds the CLIP visual
    tower at the test resolution (both random init under the seed).
Recorded: the dense network's loss and per-parameter gradient norms (distill on: the loss must NOT depend on the momentum
twins), the closed-set ``rank_answer`` output, then the same network patched by the reference mPLUG masker: module census, trainable set, thresholds, kept
counts, masked loss and score-gradient norms.
"""
import contextlib
import io
import json
import logging
import os
import sys
import tempfile
import types

import torch
from torch import nn

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_mplug as mgm  # noqa: E402

REF = mgm.REF

BERT = dict(vocab_size=60, hidden_size=64, num_hidden_layers=3, num_attention_heads=4, intermediate_size=128,
            hidden_act="gelu", hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, max_position_embeddings=32,
            type_vocab_size=2, initializer_range=0.02, layer_norm_eps=1e-12, pad_token_id=0, encoder_width=64,
            add_cross_attention=False, use_cache=False, gradient_checkpointing=False, text_encoder_layers=1,
            fusion_layers=3, text_decode_layers=1, stride_layer=2)
CONFIG = dict(image_res=64, vision_width=64, distill=True, clip_name="ViT-B-16", use_checkpoint=False,
              clip_width=64, clip_layers=1, clip_heads=4, clip_output_dim=32, clip_patch_size=16,
              min_length=1, max_length=5, beam_size=2, add_ocr=False, add_object=False)
LAYERS = {"visual_encoder": [0], "text_encoder": [0], "fusion_encoder": [0, 1, 2], "text_decoder": [0]}


def load_reference_model_classes():
    sys.path.insert(0, os.path.join(REF, "mPLUG"))
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    for n in ("apply_chunking_to_forward", "find_pruneable_heads_and_indices", "prune_linear_layer"):
        if not hasattr(mu, n):
            setattr(mu, n, getattr(pu, n, None))
    sys.modules.setdefault("ftfy", types.ModuleType("ftfy"))
    vit = types.ModuleType("models.vit")
    vit.VisionTransformer = object
    import models  # noqa: F401  (the reference mPLUG/models package must be the one in sys.modules)
    sys.modules["models.vit"] = vit
    import models.modeling_mplug as mm
    import models.clip.model as cm
    import models.visual_transformers as vt

    def init_weights(self):
        self.apply(self._init_weights)
        out = self.get_output_embeddings() if hasattr(self, "get_output_embeddings") else None
        if out is not None and getattr(self.config, "tie_word_embeddings", True):
            out.weight = self.get_input_embeddings().weight

    def invert_attention_mask(self, m):
        e = m[:, None, :, :] if m.dim() == 3 else m[:, None, None, :]
        return (1.0 - e.to(self.dtype)) * -10000.0

    mm.BertPreTrainedModel.init_weights = init_weights
    mm.BertPreTrainedModel.invert_attention_mask = invert_attention_mask
    mm.BertPreTrainedModel.get_head_mask = lambda self, head_mask, n, *a: [None] * n
    mm.BertLMHeadModel.get_input_embeddings = lambda self: self.bert.embeddings.word_embeddings
    for cls in (mm.BertModel, mm.FusionModel, mm.BertLMHeadModel):
        cls.from_pretrained = classmethod(lambda c, name, config=None, **kw: c(config, **kw))

    class Shell(nn.Module):
        def __init__(self, visual):
            super().__init__()
            self.visual = visual

    def initialize_clip(config, num_patches=240):
        return Shell(cm.VisualTransformer(input_resolution=config["image_res"], patch_size=config["clip_patch_size"],
                                          width=config["clip_width"], layers=config["clip_layers"],
                                          heads=config["clip_heads"], output_dim=config["clip_output_dim"])), None

    vt.initialize_clip = initialize_clip
    import models.model_vqa_mplug as mv
    mv.initialize_clip = initialize_clip
    return mv


def batch(seed=5, B=4):
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(B, 3, CONFIG["image_res"], CONFIG["image_res"], generator=g)
    q_ids = torch.randint(1, BERT["vocab_size"], (B, 7), generator=g)
    q_att = torch.ones(B, 7, dtype=torch.long)
    q_att[1, 5:] = 0
    q_ids[1, 5:] = 0
    k = [2, 1, 3, 2]
    n = sum(k)
    a_ids = torch.randint(1, BERT["vocab_size"], (n, 5), generator=g)
    a_att = torch.ones(n, 5, dtype=torch.long)
    a_ids[0, 3:] = 0
    a_att[0, 3:] = 0
    a_ids[5, 4:] = 0
    a_att[5, 4:] = 0
    weights = torch.rand(n, generator=g) + 0.2
    bias = torch.rand(n, generator=g) * 0.5
    question = types.SimpleNamespace(input_ids=q_ids, attention_mask=q_att)
    answer = types.SimpleNamespace(input_ids=a_ids, attention_mask=a_att)
    return image, question, answer, k, weights, bias


def run(model, with_bias):
    image, question, answer, k, weights, bias = batch()
    for p in model.parameters():
        p.grad = None
    loss = model(image, question, answer, train=True, alpha=0.4, k=k, weights=weights, bias=bias if with_bias else None)
    loss.backward()
    norms = {n: float(p.grad.norm()) for n, p in model.named_parameters() if p.grad is not None}
    return float(loss.detach()), norms


def main():
    mv = load_reference_model_classes()
    M, SP = mgm.load_reference()
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(BERT, f)
    config = dict(CONFIG, bert_config=f.name, text_encoder="none", text_decoder="none")
    torch.manual_seed(13)
    model = mv.MPLUG(config=config, tokenizer=types.SimpleNamespace(pad_token_id=0))
    # the fusion twin is not in the reference's model_pairs (initialised separately, never updated): make it a copy so
    # the fixture needs to carry the online weights only
    model.fusion_encoder_m.load_state_dict(model.fusion_encoder.state_dict())
    model.eval()                                   # dropout off (the CLIP blocks hard-code dropout=0.1); grads still flow
    sd = model.state_dict()
    online = {k: v.clone() for k, v in sd.items() if not any(t in k for t in (
        "visual_encoder_m.", "text_encoder_m.", "fusion_encoder_m.", "text_decoder_m."))}
    gold = {"bert": BERT, "config": CONFIG, "layers_to_mask": LAYERS, "state_dict_keys": sorted(sd), "online": online,
            "tied": model.text_decoder.cls.predictions.decoder.weight is
            model.text_decoder.bert.embeddings.word_embeddings.weight}
    gold["dense_loss"], gold["dense_grad_norms"] = run(model, with_bias=False)
    gold["dense_loss_bias"], _ = run(model, with_bias=True)
    twin_after = {k: v.clone() for k, v in model.state_dict().items() if k.startswith("text_encoder_m.")}
    gold["twin_moved"] = any(not torch.equal(v, sd[k]) for k, v in twin_after.items())

    # ---- closed-set ranking on the dense network (rank_answer :188-245, tile :247-253)
    with torch.no_grad():
        image, question, answer, k, weights, bias = batch()
        image_embeds = model.visual_encoder.visual(image, skip_last_layer=True, use_checkpoint=False)
        image_atts = torch.ones(image_embeds.size()[:-1], dtype=torch.long)
        text = model.text_encoder(question.input_ids, attention_mask=question.attention_mask,
                                  return_dict=True).last_hidden_state
        img_out, q_out = model.fusion_encoder(encoder_embeds=text, attention_mask=question.attention_mask,
                                      encoder_hidden_states=image_embeds, encoder_attention_mask=image_atts,
                                      return_dict=False)
        states, atts = torch.cat([img_out, q_out], 1), torch.cat([image_atts, question.attention_mask], 1)
        cand_ids = answer.input_ids.clone()
        cand_ids[:, 0] = 7                                  # a shared [BOS] token, as the tokenizer would emit
        ids, probs = model.rank_answer(states, atts, cand_ids, answer.attention_mask, 3)
        gold["rank"] = {"bos": 7, "k": 3, "topk_ids": ids.clone(), "topk_probs": probs.clone(),
                        "states_norm": float(states.norm())}
        gold["tile"] = mv.tile(torch.arange(6).view(2, 3), 0, 3).clone()

    # ---- masked: the reference masker with init_masker's wiring (vqa_mplug.py:59-128) on a fresh copy of the weights
    torch.manual_seed(13)
    model = mv.MPLUG(config=config, tokenizer=types.SimpleNamespace(pad_token_id=0))
    model.fusion_encoder_m.load_state_dict(model.fusion_encoder.state_dict())
    model.eval()
    masker, sched = mgm.make_masker(M, SP, zero_rate=0.5, init_sparsity=None, final_epoch=1,
                                    controlled_init="magnitude_soft", global_prune=False)
    weight_types = {"visual_encoder": ["I_visual", "O_visual"], "text_encoder": ["K", "Q", "V", "AO", "I", "O"],
                    "fusion_encoder": ["SK", "SQ", "SV", "SAO", "CK", "CQ", "CV", "CAO", "I", "O"],
                    "text_decoder": ["SK", "SQ", "SV", "SAO", "CK", "CQ", "CV", "CAO", "I", "O"]}
    names = set()
    for tower, ab in weight_types.items():
        names.update(M.chain_module_names(tower, LAYERS[tower], ab))
    with contextlib.redirect_stdout(io.StringIO()):
        masker.patch_modules(model=model, names_tobe_masked=names, name_of_masker="MaskedLinear1")
    mods = mgm.masked(model)
    gold["masked"] = {
        "module_names": [n for n, _ in mods],
        "trainable": sorted(n for n, p in model.named_parameters() if p.requires_grad),
        "thresholds": mgm.thr_record(model),
        "kept": {n: int(m.get_masks()[0].float().sum()) for n, m in mods}}
    loss, norms = run(model, with_bias=True)
    gold["masked"]["loss"] = loss
    gold["masked"]["grad_norms"] = norms
    mean = M.reset_threshold(model, 0.7)
    gold["masked"]["reset_0.7"] = {"mean": mean, "thresholds": mgm.thr_record(model),
                                   "kept": {n: int(m.get_masks()[0].float().sum()) for n, m in mods}}
    loss, _ = run(model, with_bias=True)
    gold["masked"]["loss_after_reset"] = loss
    os.unlink(f.name)
    path = os.path.join(HERE, "mplug_model_tiny.pt")
    torch.save(gold, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB; tied", gold["tied"], "dense", gold["dense_loss"],
          gold["dense_loss_bias"], "twin moved", gold["twin_moved"], "masked", gold["masked"]["loss"],
          gold["masked"]["loss_after_reset"], len(mods), "modules;", len(gold["masked"]["trainable"]), "trainable")


if __name__ == "__main__":
    logging.disable(logging.WARNING)
    main()
