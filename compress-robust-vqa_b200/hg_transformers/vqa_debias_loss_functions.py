"""Debias losses of the answer head (reference hg_transformers/vqa_debias_loss_functions.py).

``LearnedMixin`` (LMH), ``BiasProduct`` and ``Plain`` run as ONE fused forward+backward CUDA kernel over [B, A]
(libcrvqa: crv_vqa_loss_lmh / crv_vqa_loss_bce; BiasProduct is the LMH expression with factor 1 and no entropy
term) on CUDA tensors; ``ReweightByInvBias`` -- which no script selects -- stays a torch graph, and so does
``BiasProduct`` on CPU tensors (host-logic tests)."""
import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from crvqa import ops


def convert_sigmoid_logits_to_binary_logprobs(logits):
    log_prob = -F.softplus(-logits)
    return log_prob, -logits + log_prob


def elementwise_logsumexp(a, b):
    return torch.max(a, b) + torch.log1p(torch.exp(-torch.abs(a - b)))


def renormalize_binary_logits(a, b):
    norm = elementwise_logsumexp(a, b)
    return a - norm, b - norm


class DebiasLossFn(nn.Module):
    def forward(self, hidden, logits, bias, labels):
        raise NotImplementedError()


class Plain(DebiasLossFn):
    def forward(self, hidden, logits, bias, labels):
        return ops.vqa_loss_bce(logits, labels)[0]


class ReweightByInvBias(DebiasLossFn):
    def forward(self, hidden, logits, bias, labels):
        log_prob, log_one_minus_prob = convert_sigmoid_logits_to_binary_logprobs(logits)
        loss = -(log_prob * labels + (1 - labels) * log_one_minus_prob)
        weights = 1 - bias
        return (loss * weights).sum() / weights.sum()


class _Smoothed(DebiasLossFn):
    def __init__(self, smooth, smooth_init, constant_smooth):
        super().__init__()
        self.constant_smooth = constant_smooth
        self.smooth_init = smooth_init
        self.smooth = smooth
        self.smooth_param = (nn.Parameter(torch.from_numpy(np.full((1,), smooth_init, dtype=np.float32)))
                             if smooth else None)
        self._smooth_cache = None

    def smooth_value(self):
        """constant_smooth + sigmoid(smooth_param) as a host float.  The parameter is never handed to the
        optimiser by any driver (init_optimizer walks model.named_parameters() only), so the value is
        cached against the parameter's version counter instead of being read back every step."""
        s = float(self.constant_smooth)
        if self.smooth:
            key = (self.smooth_param._version, self.smooth_param.data_ptr())
            if self._smooth_cache is None or self._smooth_cache[0] != key:
                self._smooth_cache = (key, float(torch.sigmoid(self.smooth_param.detach().float().cpu())[0]))
            s += self._smooth_cache[1]
        return s


class BiasProduct(_Smoothed):
    def __init__(self, smooth=True, smooth_init=-1, constant_smooth=0.0):
        super().__init__(smooth, smooth_init, constant_smooth)

    def forward(self, hidden, logits, bias, labels):
        if logits.is_cuda:
            return ops.vqa_loss_bias_product(logits, bias, labels, self.smooth_value())[0]
        smooth = self.constant_smooth
        if self.smooth:
            smooth = smooth + torch.sigmoid(self.smooth_param)
        bias_lp, bias_l_inv = torch.log(bias + smooth), torch.log1p(-bias + smooth)
        log_prob, log_one_minus_prob = convert_sigmoid_logits_to_binary_logprobs(logits)
        log_prob, log_one_minus_prob = renormalize_binary_logits(log_prob + bias_lp, log_one_minus_prob + bias_l_inv)
        return -(log_prob * labels + (1 - labels) * log_one_minus_prob).sum(1).mean(0)


class LearnedMixin(_Smoothed):
    """LMH: loss = -mean_b sum_a [renormalised two-class log-probs of logsigmoid(+-logit) + factor *
    log(bias + smooth)] + w * mean entropy of the scaled bias (reference :125-196).  `last_score` holds
    the batch VQA score the fused kernel computes alongside."""

    def __init__(self, w, smooth=True, smooth_init=-1, constant_smooth=0.0, hidden_size=768):
        super().__init__(smooth, smooth_init, constant_smooth)
        self.w = w
        self.bias_lin = torch.nn.Linear(hidden_size, 1)
        self.last_score = None

    def forward(self, hidden, logits, bias, labels, device=None):
        factor_pre = self.bias_lin(hidden)  # [B, 1]; softplus and its gradient are inside the kernel
        loss, score = ops.vqa_loss_lmh(logits, bias, labels, factor_pre, self.smooth_value(), self.w)
        self.last_score = score
        return loss
