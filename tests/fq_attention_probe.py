"""Measurement aid (not a test): the few-query attention kernels (crv_fq_attention_fwd / _bwd) on the text-side
attention shapes of one mPLUG-base training step, same inputs as tests/sdpa_shapes_probe.py times torch's SDPA on.
One JSON line per shape: forward and forward + backward microseconds.   python tests/fq_attention_probe.py
"""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "compress-robust-vqa_b200"))

SHAPES = [("text_self", 10, 32, 12, 16, 16, "row"), ("fusion_cross", 4, 32, 12, 16, 577, "row"),
          ("decoder_self", 12, 64, 12, 6, 6, "causal"), ("decoder_cross", 12, 64, 12, 6, 593, "row")]


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
    from crvqa import fused
    torch.manual_seed(0)
    for name, calls, B, H, Lq, Lk, mk in SHAPES:
        q = torch.randn(B, Lq, H * 64, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        k = torch.randn(B, Lk, H * 64, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        v = torch.randn(B, Lk, H * 64, device="cuda", dtype=torch.bfloat16, requires_grad=True)
        if mk == "row":
            mask = torch.zeros(B, 1, 1, Lk, device="cuda")
            mask[1, :, :, Lk - 2:] = -10000.0
        else:
            mask = torch.zeros(B, 1, Lq, Lk, device="cuda").masked_fill_(
                torch.ones(Lq, Lk, device="cuda").triu(1).bool(), -10000.0)
        site = fused.RngState.new_site()
        dout = torch.ones(B, Lq, H * 64, device="cuda", dtype=torch.bfloat16)

        def fwd():
            return fused.few_query_attention(q, k, v, mask, H, 0.1, site, True)

        def fwd_bwd():
            fwd().backward(dout)
            q.grad = k.grad = v.grad = None

        t_f, t_fb = timed(fwd), timed(fwd_bwd)
        kv_mb = 2 * B * Lk * H * 64 * 2 / 1e6
        print(json.dumps({"shape": name, "calls_per_step": calls, "B": B, "H": H, "Lq": Lq, "Lk": Lk, "mask": mk,
                          "fwd_us": round(t_f, 1), "fwd_bwd_us": round(t_fb, 1), "kv_megabytes": round(kv_mb, 1),
                          "fwd_gb_per_s": round(kv_mb / t_f * 1e3, 1)}), flush=True)


if __name__ == "__main__":
    main()
