"""Mask-training pieces of the reference's ``mPLUG/vqa_mplug.py``.

Kept with the reference's names and semantics: ``load_mask_and_prune`` (:44-52), ``encode_maskconfig`` (:54-57),
``init_masker`` (:59-128: scheduler wiring, the four towers' weight types / layers, ``_m`` twins), the mask-update
block of the training loop (:202-210, here ``update_masks``), ``train`` (:130-217, same signature; plus
``train_pretokenized`` for batches that already are model arguments), and the generation-side helpers ``evaluation`` (:219-246), ``evaluate`` (:248-290), ``cal_metric``
(:292-306) and ``save_result`` (:309-311).  The reference driver itself needs DeepSpeed, the CLIP / BERT checkpoints and the VQA image datasets (none
shipped); the engine of ``engine.py`` stands where the DeepSpeed engine stands.
"""
import argparse
import json
import logging
import os
import types

import torch

if __package__:                                     # imported as mPLUG.vqa_mplug (package root on sys.path)
    from . import param_parser
    from .masking import maskers
    from .masking import sparsity_control as sp_control
    from .masking.mask_config import MaskConfigs
    from .masking.pruned import custom_from_mask
else:                                               # the reference's layout: mPLUG/ itself is on sys.path
    import param_parser
    from masking import maskers
    from masking import sparsity_control as sp_control
    from masking.mask_config import MaskConfigs
    from masking.pruned import custom_from_mask

reset_threshold, save_model_mask, see_sparsity = maskers.reset_threshold, maskers.save_model_mask, maskers.see_sparsity

# which Linear layers of which tower are masked (:105-116)
WEIGHT_TYPES = {
    "visual_encoder": ["I_visual", "O_visual"],
    "text_encoder": ["K", "Q", "V", "AO", "I", "O"],
    "fusion_encoder": ["SK", "SQ", "SV", "SAO", "CK", "CQ", "CV", "CAO", "I", "O"],
    "text_decoder": ["SK", "SQ", "SV", "SAO", "CK", "CQ", "CV", "CAO", "I", "O"],
}
LAYERS_TO_MASK = {
    "visual_encoder": list(range(12)),
    "text_encoder": list(range(6)),
    "fusion_encoder": list(range(6, 12)),
    "text_decoder": list(range(12)),
}


def load_mask_and_prune(mask_dir, model):
    """Freeze a saved mask into the network: every ``<module>.weight`` entry of mask.pt reparametrises that module as
    ``weight_orig * weight_mask`` (the reference uses prune.CustomFromMask; same parameter / buffer names here)."""
    print("Loading mask from %s" % mask_dir)
    masks = torch.load(os.path.join(mask_dir, "mask.pt"))
    for k, m in masks.items():
        custom_from_mask(model, k.replace("module.", "").replace(".weight", ""), m.bool())
    return model


def encode_maskconfig(obj):
    return obj.__dict__ if isinstance(obj, MaskConfigs) else obj


def names_to_mask(conf, weight_types=None, layers_to_mask=None):
    weight_types = WEIGHT_TYPES if weight_types is None else weight_types
    layers_to_mask = LAYERS_TO_MASK if layers_to_mask is None else layers_to_mask
    names = set()
    for tower, abbres in weight_types.items():
        names.update(maskers.chain_module_names(tower, layers_to_mask[tower], abbres))
    if conf.mask_classifier:
        names.add("text_decoder_m.cls.predictions.transform.dense")
    return names


def init_masker(conf, model, weight_types=None, layers_to_mask=None):
    """Build the scheduler and the masker from a MaskConfigs and patch ``model`` in place.  ``weight_types`` /
    ``layers_to_mask`` default to the reference's tables (they are literals inside the reference function)."""
    mask_logger = logging.getLogger(__name__)
    conf.masking_scheduler_conf_ = (param_parser.dict_parser(conf.masking_scheduler_conf)
                                    if conf.masking_scheduler_conf is not None else None)
    conf.masking_scheduler_conf_["final_sparsity"] = conf.zero_rate
    conf.masking_scheduler_conf_["final_epoch"] = conf.final_sparsity_epoch
    if conf.init_sparsity is not None:
        conf.masking_scheduler_conf_["init_sparsity"] = conf.init_sparsity
    if conf.masking_scheduler_conf is not None:
        for k, v in conf.masking_scheduler_conf_.items():
            setattr(conf, f"masking_scheduler_{k}", v)
    conf.logger = mask_logger
    masker_scheduler = sp_control.MaskerScheduler(conf)

    assert not (conf.train_classifier and conf.mask_classifier), \
        "If the classifier is masked, don't train its weights!"
    masker = maskers.Masker(
        masker_scheduler=masker_scheduler, logger=mask_logger, mask_biases=conf.mask_biases,
        structured_masking_info={"structured_masking": conf.structured_masking,
                                 "structured_masking_types": conf.structured_masking_types,
                                 "force_masking": conf.force_masking},
        threshold=conf.threshold, init_scale=conf.init_scale, controlled_init=conf.controlled_init,
        train_classifier=conf.train_classifier, global_prune=conf.global_prune)
    masker.patch_modules(model=model, names_tobe_masked=names_to_mask(conf, weight_types, layers_to_mask),
                         name_of_masker=conf.name_of_masker)
    return masker


def update_masks(model, masker, epoch, output_dir=None):
    """The block the reference runs every ``masker_update_step`` optimiser steps (:202-210): advance the sparsity
    schedule, refresh every threshold for the new target, export the masks, report the sparsity."""
    _, target_sparsity, _ = masker.masker_scheduler.step(cur_epoch=epoch)
    mean_thresh = reset_threshold(model, target_sparsity)
    save_model_mask(model, is_save=False)
    if output_dir is not None:
        save_model_mask(model, is_save=True, output_dir=output_dir)
    print({"mean_thresh": mean_thresh})
    see_sparsity(model)
    return mean_thresh, target_sparsity


def _train_step(model, args, masker, masker_update_step, epoch, output_dir):
    loss = model(*args[0], **args[1])
    model.backward(loss)
    model.step()
    if masker is not None and model.global_steps % masker_update_step == 0:
        update_masks(model, masker, epoch, output_dir)
    return loss.detach()


def train_pretokenized(model, data_loader, epoch, masker=None, masker_update_step=5, output_dir=None, log=None):
    """One epoch on an engine (``engine.MaskTrainEngine``) over batches that are already the positional arguments of
    ``model(...)``: forward -> backward -> step, and the mask update whenever ``global_steps`` is a multiple of
    ``masker_update_step``.  Returns the mean loss."""
    model.train()
    total, count = None, 0
    for batch in data_loader:
        loss = _train_step(model, (batch, {}), masker, masker_update_step, epoch, output_dir)
        total = loss if total is None else total + loss
        count += 1
        if log is not None:
            log(model.global_steps, loss)
    return float(total / count) if count else float("nan")


def train(model, data_loader, optimizer, tokenizer, epoch, warmup_steps, device, scheduler, config, do_amp=False,
          do_two_optim=False, do_accum=False, accum_steps=1, masker=None, masker_update_step=5, output_dir=None,
          max_input_length=25):
    """One epoch of the reference loop with its own signature (:130-217).  Batches come from ``vqa_collate_fn`` /
    ``vqa_bias_collate_fn`` (image, questions, answers, weights, n[, bias]); questions and answers are tokenised here;
    the distillation weight ramps over epoch 0 when ``config['warm_up']``; during epoch 0 the scheduler is stepped every
    100 iterations up to ``warmup_steps * 100``; the masks are refreshed every ``masker_update_step`` optimiser steps.
    ``model`` is the engine (``model(...)``, ``.backward``, ``.step``, ``.global_steps``); ``do_amp`` / ``do_accum`` /
    ``accum_steps`` are accepted and, as in the reference (whose code for them is commented out), unused.  Returns the
    reference's stats dict: {"lr" | "lr1", "lr2", "loss"} as formatted averages."""
    model.train()
    step_size = 100
    warmup_iterations = warmup_steps * step_size
    losses, lr_log = [], []
    n_batches = len(data_loader)
    for i, data in enumerate(data_loader):
        if len(data) == 5:
            image, question, answer, weights, n = data
            bias = None
        else:
            image, question, answer, weights, n, bias = data
            bias = bias.to(device, non_blocking=True)
        image, weights = image.to(device, non_blocking=True), weights.to(device, non_blocking=True)
        question_input = tokenizer(question, padding="longest", truncation=True,
                                   max_length=max_input_length if config.get("add_ocr") else 25,
                                   return_tensors="pt").to(device)
        answer_input = tokenizer(answer, padding="longest", return_tensors="pt").to(device)
        if epoch > 0 or not config["warm_up"]:
            alpha = config["alpha"]
        else:
            alpha = config["alpha"] * min(1, i / n_batches)
        kwargs = dict(train=True, alpha=alpha, k=n, weights=weights, bias=bias)
        loss = _train_step(model, ((image, question_input, answer_input), kwargs), masker, masker_update_step, epoch,
                           output_dir)
        losses.append(loss)
        lr_log.append((optimizer.param_groups[0]["lr"], optimizer.param_groups[2]["lr"] if do_two_optim else None))
        if epoch == 0 and i % step_size == 0 and i <= warmup_iterations:
            scheduler.step(i // step_size)
    mean = float(torch.stack(losses).mean()) if losses else float("nan")
    stats = {"loss": "{:.3f}".format(mean)}
    if lr_log:
        if do_two_optim:
            stats["lr1"] = "{:.3f}".format(sum(a for a, _ in lr_log) / len(lr_log))
            stats["lr2"] = "{:.3f}".format(sum(b for _, b in lr_log) / len(lr_log))
        else:
            stats["lr"] = "{:.3f}".format(sum(a for a, _ in lr_log) / len(lr_log))
    return stats


# --------------------------------------------------------------------------- generation-side helpers
def _decode(tokenizer, ids):
    text = tokenizer.decode(ids)
    return text.replace("[SEP]", "").replace("[CLS]", "").replace("[PAD]", "").strip()


@torch.no_grad()
def evaluation(model, data_loader, tokenizer, device, config):
    """Generate one answer per question: [{"question_id", "answer"}] (beam search, best hypothesis)."""
    model.eval()
    result = []
    for image, question, question_id in data_loader:
        image = image.to(device, non_blocking=True)
        question_input = tokenizer(question, padding="longest", return_tensors="pt").to(device)
        topk_ids, _ = model(image, question_input, None, train=False, k=config["k_test"])
        for ques_id, topk_id in zip(question_id, topk_ids):
            result.append({"question_id": int(ques_id.item() if torch.is_tensor(ques_id) else ques_id),
                           "answer": _decode(tokenizer, topk_id[0])})
    return result


def cal_metric(vqa_result, val_file):
    """Mean soft VQA score of the generated answers against ``val_file[0]`` (a JSON list of {question_id, label})."""
    with open(val_file[0], "r") as f:
        id2datum = {each["question_id"]: each["label"] for each in json.load(f)}
    score = 0.0
    for each in vqa_result:
        label = id2datum[each["question_id"]]
        if each["answer"] in label:
            score += label[each["answer"]]
    return score / len(vqa_result)


def save_result(result, output_file):
    with open(output_file, "w") as f:
        json.dump(result, f)


@torch.no_grad()
def evaluate(model, data_loader, dataset, tokenizer, device, config, output_dir):
    """Generate, score every batch against the label file ``dataset`` and write ``vqa_answer.json``; returns
    ``{"acc": "<mean over questions, 4 decimals>"}`` like the reference's metric logger."""
    model.eval()
    results, total, count = [], 0.0, 0
    for image, question, question_id in data_loader:
        batch = evaluation(model, [(image, question, question_id)], tokenizer, device, config)
        results += batch
        total += cal_metric(batch, dataset) * len(batch)
        count += len(batch)
    os.makedirs(output_dir, exist_ok=True)
    save_result(results, os.path.join(output_dir, "vqa_answer.json"))
    return {"acc": "{:.4f}".format(total / max(count, 1))}


# --------------------------------------------------------------------------- synthetic entry point
TINY = dict(clip_width=64, clip_layers=2, clip_heads=4, clip_output_dim=32, clip_patch_size=16, vision_width=64,
            bert_config=dict(vocab_size=128, hidden_size=64, num_hidden_layers=4, num_attention_heads=4,
                             intermediate_size=128, max_position_embeddings=32, encoder_width=64, fusion_layers=2,
                             stride_layer=1, text_encoder_layers=2, text_decode_layers=2))
TINY_LAYERS = {"visual_encoder": [0, 1], "text_encoder": [0, 1], "fusion_encoder": [2, 3], "text_decoder": [0, 1]}


def main(argv=None):
    """The reference's main() flow (:313-460) on synthetic image-VQA data and random init: model -> init_masker ->
    create_two_optimizer -> create_scheduler -> engine (in place of deepspeed.initialize) -> train()."""
    if __package__:
        from .dataset import SyntheticVQAImageDataset, WhitespaceTokenizer, vqa_bias_collate_fn
        from .engine import MaskTrainEngine
        from .models.model_vqa_mplug import MPLUG
        from .optim import create_two_optimizer
        from .scheduler import create_scheduler
    else:
        from dataset import SyntheticVQAImageDataset, WhitespaceTokenizer, vqa_bias_collate_fn
        from engine import MaskTrainEngine
        from models.model_vqa_mplug import MPLUG
        from optim import create_two_optimizer
        from scheduler import create_scheduler
    ap = argparse.ArgumentParser(description="synthetic mPLUG masked training")
    ap.add_argument("--image_res", type=int, default=384)
    ap.add_argument("--batch_size", type=int, default=8)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--masker_update_step", type=int, default=2)
    ap.add_argument("--zero_rate", type=float, default=0.7)
    ap.add_argument("--tiny", action="store_true", help="a miniature network (smoke runs)")
    ap.add_argument("--no_mask", action="store_true", help="dense network, no masker (runs on a CPU)")
    ap.add_argument("--output_dir", default=None)
    ap.add_argument("--device", default="cuda" if torch.cuda.is_available() else "cpu")
    ap.add_argument("--seed", type=int, default=42)
    a = ap.parse_args(argv)
    torch.manual_seed(a.seed)
    device = torch.device(a.device)
    config = dict(image_res=a.image_res, vision_width=768, distill=True, clip_name="ViT-B-16", alpha=0.4, warm_up=True,
                  add_ocr=False, bert_config=dict(stride_layer=3, fusion_layers=6, text_encoder_layers=6,
                                                  text_decode_layers=12))
    if a.tiny:
        config.update(TINY)
    tokenizer = WhitespaceTokenizer()
    model = MPLUG(config=config, tokenizer=tokenizer).to(device)
    masker = None
    if not a.no_mask:
        conf = MaskConfigs()
        conf.zero_rate = a.zero_rate
        masker = init_masker(conf, model, layers_to_mask=TINY_LAYERS if a.tiny else None)
    optimizer = create_two_optimizer(types.SimpleNamespace(lr1=3e-5, lr2=5e-6, weight_decay=0.02), model)
    scheduler, _ = create_scheduler(types.SimpleNamespace(sched="cosine", lr=3e-5, epochs=8, min_lr=1e-6, decay_rate=1,
                                                          warmup_lr=1e-5, warmup_epochs=4, cooldown_epochs=0), optimizer)
    engine = MaskTrainEngine(model, optimizer, gradient_clipping=1.0, bf16=True)
    data = SyntheticVQAImageDataset(a.batch_size * a.steps, image_res=a.image_res, seed=a.seed)
    loader = torch.utils.data.DataLoader(data, batch_size=a.batch_size, collate_fn=vqa_bias_collate_fn)
    stats = train(engine, loader, optimizer, tokenizer, 0, 4, device, scheduler, config, do_two_optim=True,
                  masker=masker, masker_update_step=a.masker_update_step, output_dir=a.output_dir)
    print(stats)
    return stats


if __name__ == "__main__":
    main()
