"""Answer head (reference: hg_transformers/classifier.py:5-22, fc.py:6-19).

Weight-normed Linear(in,hid) -> ReLU -> Dropout -> weight-normed Linear(hid,out).  It is the only
trainable *weight* in stage 2 and is never masked, so it stays a stock PyTorch module; the
state-dict keys (main.0.weight_g / weight_v / bias, main.3.*) match the reference so
``classifier4masker.bin`` files interchange.
"""
import warnings

import torch.nn as nn


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
        return self.main(x)
