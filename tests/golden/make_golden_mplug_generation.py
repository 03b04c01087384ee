"""Golden vectors for mPLUG answer generation (``MPLUG.forward(train=False)`` -> predictor.TextGenerator beam search),
from the UNMODIFIED reference classes run on the CPU at a miniature size (shims: make_golden_mplug_model.py):

    python tests/golden/make_golden_mplug_generation.py        # writes tests/golden/mplug_generation_tiny.pt

The search hard-codes [CLS] = 101 and [SEP] = 102, so the vocabulary must exceed 102; the LM-head bias of the random
network is raised on [SEP] so that beams actually finish at different steps (early close, dropped questions, a beam that
ends while not being the best one).
"""
import json
import os
import sys
import tempfile
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_mplug_model as mm  # noqa: E402

BERT = dict(mm.BERT, vocab_size=110, hidden_size=32, num_attention_heads=2, intermediate_size=64, encoder_width=32,
            num_hidden_layers=2, fusion_layers=2, stride_layer=1)
CONFIG = dict(mm.CONFIG, image_res=32, vision_width=32, clip_width=32, clip_heads=2, clip_output_dim=16, distill=False,
              min_length=1, max_length=6, beam_size=3)


def batch(seed=9, B=5):
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(B, 3, CONFIG["image_res"], CONFIG["image_res"], generator=g)
    q_ids = torch.randint(1, 100, (B, 6), generator=g)
    q_att = torch.ones(B, 6, dtype=torch.long)
    q_att[2, 4:] = 0
    q_ids[2, 4:] = 0
    image = image * torch.arange(1, B + 1).view(-1, 1, 1, 1).float()
    return image, types.SimpleNamespace(input_ids=q_ids, attention_mask=q_att)


def main():
    mv = mm.load_reference_model_classes()
    with tempfile.NamedTemporaryFile("w", suffix=".json", delete=False) as f:
        json.dump(BERT, f)
    config = dict(CONFIG, bert_config=f.name, text_encoder="none", text_decoder="none")
    torch.manual_seed(21)
    model = mv.MPLUG(config=config, tokenizer=types.SimpleNamespace(pad_token_id=0))
    with torch.no_grad():          # make [SEP] competitive, the logits less flat and the question / image matter
        model.text_decoder.cls.predictions.bias[102] = 0.5
        model.text_decoder.bert.embeddings.word_embeddings.weight.mul_(5.0)
        for lyr in model.text_decoder.bert.encoder.layer:
            lyr.crossattention.self.value.weight.mul_(8.0)
            lyr.crossattention.output.dense.weight.mul_(8.0)
        for lyr in model.fusion_encoder.encoder.layer:
            lyr.crossattention.self.value.weight.mul_(8.0)
    model.eval()
    image, question = batch()
    runs = {}
    for beam, min_len, max_len in ((3, 1, 6), (2, 0, 4), (1, 2, 5)):
        model.beam_generator.beam_size, model.beam_generator.min_length, model.beam_generator.max_length = beam, min_len, max_len
        with torch.no_grad():
            ids, scores = model(image, question, None, train=False, k=None)
        runs[(beam, min_len, max_len)] = {"ids": [[t.clone() for t in q] for q in ids],
                                          "scores": [[float(s) for s in q] for q in scores]}
        if beam == 3:                                      # the full ranked list of every question
            with torch.no_grad():
                image_embeds = model.visual_encoder.visual(image, skip_last_layer=True, use_checkpoint=False)
                image_atts = torch.ones(image_embeds.size()[:-1], dtype=torch.long)
                text = model.text_encoder(question.input_ids, attention_mask=question.attention_mask,
                                          return_dict=True).last_hidden_state
                img_out, q_out = model.fusion_encoder(encoder_embeds=text, attention_mask=question.attention_mask,
                                                      encoder_hidden_states=image_embeds,
                                                      encoder_attention_mask=image_atts, return_dict=False)
                full_ids, full_scores = model.beam_generator.translate_batch(
                    [torch.cat([img_out, q_out], 1), torch.cat([image_atts, question.attention_mask], 1)], out_size=3)
            runs["ranked3"] = {"ids": [[t.clone() for t in q] for q in full_ids],
                               "scores": [[float(s) for s in q] for q in full_scores]}
        print((beam, min_len, max_len), [[t.tolist() for t in q] for q in ids], [[round(float(s), 4) for s in q] for q in scores])
    # a slightly stronger [SEP]: questions 1 and 4 now close early (their best beam ends at steps 4 / 5) and leave the
    # batch while the others run on to max_length
    with torch.no_grad():
        model.text_decoder.cls.predictions.bias[102] = 0.78
    runs["early_close_bias"] = 0.78
    model.beam_generator.beam_size, model.beam_generator.min_length, model.beam_generator.max_length = 3, 1, 6
    with torch.no_grad():
        early_ids, early_scores = model.beam_generator.translate_batch(
            [torch.cat([img_out, q_out], 1), torch.cat([image_atts, question.attention_mask], 1)], out_size=3)
    runs["early_close"] = {"ids": [[t.clone() for t in q] for q in early_ids],
                           "scores": [[float(s) for s in q] for q in early_scores]}
    print("early", [[len(t) for t in q] for q in early_ids])
    with torch.no_grad():
        model.text_decoder.cls.predictions.bias[102] = 0.5
    os.unlink(f.name)
    gold = {"bert": BERT, "config": CONFIG, "state_dict": {k: v.clone() for k, v in model.state_dict().items()},
            "runs": runs}
    path = os.path.join(HERE, "mplug_generation_tiny.pt")
    torch.save(gold, path)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
