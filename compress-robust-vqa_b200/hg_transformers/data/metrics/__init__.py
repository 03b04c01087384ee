"""VQA batch score (reference hg_transformers/data/metrics/__init__.py:90-104), without the host sync:
the reference moves the argmax to the CPU every step; this version stays on the logits' device."""
import torch


def compute_score_with_logits(task_name, logits, labels):
    if task_name != "vqa":
        raise KeyError(task_name)
    assert len(logits) == len(labels)
    assert logits.shape == labels.shape
    labels = labels.to(logits.device)
    am = torch.max(logits, 1)[1]
    return {"acc": labels.gather(1, am.view(-1, 1)).sum()}


def vqa_compute_metrics(p):
    """compute_metrics callable the drivers build (prune_debias_VQA.py: build_compute_metrics_fn)."""
    return compute_score_with_logits("vqa", p.predictions, p.label_ids)
