"""Oracle restatement of the answer-head losses (torch CPU fp32, gradients by autograd)."""
import torch
import torch.nn.functional as F


def bce_loss(logits, labels):
    """instance_bce_with_logits -- hg_transformers/modeling_lxmert.py:248-253: mean BCE times A."""
    return F.binary_cross_entropy_with_logits(logits, labels, reduction="mean") * labels.size(1)


def lpf_loss(logits, bias, max_label, gamma):
    """LPF_loss -- hg_transformers/mask_trainer_VQA.py:111-129."""
    p = torch.clamp_min(F.softmax(logits, dim=-1), 1.0e-7)     # torch.max(p, 1e-7 * ones)
    q = torch.clamp_min(bias, 1.0e-7)
    idx = max_label.view(-1, 1)
    logp = torch.log(p).gather(-1, idx).view(-1)
    logq = torch.log(q).gather(-1, idx).view(-1)
    return ((1 - torch.exp(logq)) ** gamma * (-logp)).mean()


def _log_sigmoid_pair(logits):
    """convert_sigmoid_logits_to_binary_logprobs -- vqa_debias_loss_functions.py:10-14."""
    lp = -F.softplus(-logits)
    return lp, -logits + lp


def _lse2(a, b):
    """elementwise_logsumexp -- vqa_debias_loss_functions.py:17-19."""
    return torch.max(a, b) + torch.log1p(torch.exp(-torch.abs(a - b)))


def lmh_loss(pooled, logits, bias, labels, lin_w, lin_b, smooth_param, w=0.36, constant_smooth=0.0):
    """LearnedMixin.forward -- vqa_debias_loss_functions.py:148-196 (lin_w [1,H], lin_b [1])."""
    factor = F.softplus(F.linear(pooled, lin_w, lin_b))                     # [B,1]
    b2 = torch.stack([bias, 1 - bias], 2) + constant_smooth
    b2 = b2 + torch.sigmoid(smooth_param).unsqueeze(1)
    b2 = torch.log(b2) * factor.unsqueeze(1)                                # [B,A,2]
    lp, l1p = _log_sigmoid_pair(logits)
    u = b2 + torch.stack([lp, l1p], 2)
    norm = _lse2(u[:, :, 0], u[:, :, 1])
    LP, L1P = u[:, :, 0] - norm, u[:, :, 1] - norm
    sum_prob = (LP * labels + (1 - labels) * L1P).sum(1)
    sum_prob = torch.where(torch.isnan(sum_prob), torch.zeros_like(sum_prob), sum_prob)
    loss = -sum_prob.mean(0)
    bias_logprob = b2 - _lse2(b2[:, :, 0], b2[:, :, 1]).unsqueeze(2)
    entropy = -(torch.exp(bias_logprob) * bias_logprob).sum(2).mean()
    return loss + w * entropy


def vqa_score(logits, labels):
    """compute_score_with_logits -- hg_transformers/data/metrics/__init__.py:90-104."""
    am = logits.max(1)[1]
    return labels.gather(1, am.view(-1, 1)).sum()
