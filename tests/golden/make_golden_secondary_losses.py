"""Golden values for the secondary debias losses, from the UNMODIFIED reference:
RUBI_loss (hg_transformers/mask_trainer_VQA.py:131-135) and BiasProduct
(hg_transformers/vqa_debias_loss_functions.py:83-122).

    python tests/golden/make_golden_secondary_losses.py      # writes tests/golden/secondary_losses.pt
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402


def main():
    R = mg.load_reference()
    g = torch.Generator().manual_seed(11)
    B, A, H = 6, 40, 32
    logits = torch.randn(B, A, generator=g).requires_grad_(True)
    bias = torch.rand(B, A, generator=g) * 0.5
    labels = (torch.rand(B, A, generator=g) > 0.9).float() * torch.rand(B, A, generator=g)
    hidden = torch.randn(B, H, generator=g)
    max_label = labels.argmax(1)
    out = {"logits": logits.detach().clone(), "bias": bias, "labels": labels, "hidden": hidden, "max_label": max_label}
    loss = R.trainer_base.RUBI_loss(logits, bias, max_label)
    loss.backward()
    out["rubi"], out["rubi_dlogits"] = loss.detach().clone(), logits.grad.clone()
    logits.grad = None
    torch.manual_seed(3)
    bp = R.loss.BiasProduct()
    out["bp_smooth_param"] = bp.smooth_param.detach().clone() if hasattr(bp, "smooth_param") else None
    loss = bp(hidden, logits, bias, labels)
    loss.backward()
    out["bp"], out["bp_dlogits"] = loss.detach().clone(), logits.grad.clone()
    torch.save(out, os.path.join(HERE, "secondary_losses.pt"))
    print("rubi", float(out["rubi"]), "bp", float(out["bp"]), "smooth", out["bp_smooth_param"])


if __name__ == "__main__":
    main()
