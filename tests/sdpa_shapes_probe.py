"""Measurement aid (not a test): time torch's scaled_dot_product_attention per backend on the attention shapes of one
mPLUG-base training step (batch 32, 577 image tokens, 16 question / 6 answer tokens, 2 answers per question), with the
additive masks and dropout the model passes.  Prints one JSON line per (shape, backend): forward and forward+backward
microseconds.   python tests/sdpa_shapes_probe.py
"""
import json

import torch
import torch.nn.functional as F
from torch.nn.attention import SDPBackend, sdpa_kernel

SHAPES = [   # name, calls per step, B, H, Lq, Lk, mask ("row" = [B,1,1,Lk], "causal" = [B,1,Lq,Lk], None), q/k/v layout
    ("vit_self", 12, 32, 12, 577, 577, None),
    ("text_self", 6 + 4, 32, 12, 16, 16, "row"),
    ("fusion_cross", 4, 32, 12, 16, 577, "row"),
    ("fusion_joint_self", 2, 32, 12, 593, 593, "row"),
    ("decoder_self", 12, 64, 12, 6, 6, "causal"),
    ("decoder_cross", 12, 64, 12, 6, 593, "row"),
]
BACKENDS = {"default": None, "cudnn": SDPBackend.CUDNN_ATTENTION, "efficient": SDPBackend.EFFICIENT_ATTENTION,
            "math": SDPBackend.MATH}


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def main():
    torch.manual_seed(0)
    d = 64
    for name, calls, B, H, Lq, Lk, mk in SHAPES:
        # heads as strided views of [B, L, H*d] projections, as the model builds them
        q = torch.randn(B, Lq, H, d, device="cuda", dtype=torch.bfloat16).transpose(1, 2).requires_grad_(True)
        k = torch.randn(B, Lk, H, d, device="cuda", dtype=torch.bfloat16).transpose(1, 2).requires_grad_(True)
        v = torch.randn(B, Lk, H, d, device="cuda", dtype=torch.bfloat16).transpose(1, 2).requires_grad_(True)
        mask = None
        if mk == "row":
            mask = torch.zeros(B, 1, 1, Lk, device="cuda", dtype=torch.bfloat16)
            mask[1, :, :, Lk - 2:] = -10000.0
        elif mk == "causal":
            mask = torch.zeros(B, 1, Lq, Lk, device="cuda", dtype=torch.bfloat16)
            mask.masked_fill_(torch.ones(Lq, Lk, device="cuda").triu(1).bool(), -10000.0)
        for bname, backend in BACKENDS.items():
            def fwd():
                return F.scaled_dot_product_attention(q, k, v, attn_mask=mask, dropout_p=0.1)

            def fwd_bwd():
                out = fwd()
                out.backward(torch.ones_like(out))
                q.grad = k.grad = v.grad = None

            try:
                if backend is None:
                    t_f, t_fb = timed(fwd), timed(fwd_bwd)
                else:
                    with sdpa_kernel([backend]):
                        t_f, t_fb = timed(fwd), timed(fwd_bwd)
                rec = {"fwd_us": round(t_f, 1), "fwd_bwd_us": round(t_fb, 1)}
            except RuntimeError as e:
                rec = {"error": str(e).splitlines()[0][:80]}
            print(json.dumps({"shape": name, "calls_per_step": calls, "B": B, "H": H, "Lq": Lq, "Lk": Lk, "mask": mk,
                              "backend": bname, **rec}), flush=True)


if __name__ == "__main__":
    main()
