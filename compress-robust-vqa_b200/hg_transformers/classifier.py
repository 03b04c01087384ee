"""Answer head (reference: hg_transformers/classifier.py:5-22, fc.py:6-19).

Weight-normed Linear(in,hid) -> ReLU -> Dropout -> weight-normed Linear(hid,out).  It is the only
trainable *weight* in stage 2 and is never masked, so it stays a stock PyTorch module; the
state-dict keys (main.0.weight_g / weight_v / bias, main.3.*) match the reference so
``classifier4masker.bin`` files interchange.
"""
import warnings

import torch.nn as nn
import torch.nn.functional as F


def _weight_norm(module):
    # the reference uses the legacy hook-based weight_norm with dim=None (one scalar g)
    from torch.nn.utils import weight_norm
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return weight_norm(module, dim=None)


class SimpleClassifier(nn.Module):
    def __init__(self, in_dim, hid_dim, out_dim, norm="weight", act="ReLU", dropout=0.5):
        super().__init__()
        if norm != "weight" or act != "ReLU":
            raise NotImplementedError("only the weight-norm / ReLU head of the VQA path is provided")
        self.main = nn.Sequential(
            _weight_norm(nn.Linear(in_dim, hid_dim)),
            nn.ReLU(),
            nn.Dropout(dropout, inplace=False),
            _weight_norm(nn.Linear(hid_dim, out_dim)),
        )

    def forward(self, x):
        last = self.main[3]
        n = last.out_features
        if not x.is_cuda or n % 16 == 0:
            return self.main(x)
        # The VQA head has 3129 (or 2274) answers: an odd leading dimension sends cuBLAS to its unaligned sm_80 kernels
        # (43 us per GEMM at batch 256, 57 TFLOP/s).  Same math on zero-padded operands -- weight rows / bias entries
        # n .. n_pad-1 are zeros and the extra logits are sliced away -- runs the aligned sm_100 kernels instead; the
        # first n columns are untouched dot products over the hidden dimension.
        h = self.main[2](self.main[1](self.main[0](x)))
        for hook in last._forward_pre_hooks.values():      # legacy weight_norm: rebuild .weight from g, v
            hook(last, (h,))
        pad = (-n) % 16
        w = F.pad(last.weight, (0, 0, 0, pad))
        b = F.pad(last.bias, (0, pad)) if last.bias is not None else None
        return F.linear(h, w, b)[:, :n]
