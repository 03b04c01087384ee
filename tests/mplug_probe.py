"""BASELINE config 5 probe: mPLUG-base (ViT-B/16 at 384 px + 6/6/12 BERT layers) masked training on one GPU, synthetic
batch, zero rate 0.7: ms/step and samples/s of forward + backward + engine step.  Not a test; prints one JSON line.
    python tests/mplug_probe.py [batch] [steps]
"""
import json
import os
import sys
import time
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "compress-robust-vqa_b200"))


def main():
    from mPLUG import vqa_mplug
    from mPLUG.engine import MaskTrainEngine
    from mPLUG.masking.mask_config import MaskConfigs
    from mPLUG.models.model_vqa_mplug import MPLUG
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
    profile_to = sys.argv[3] if len(sys.argv) > 3 else None      # path for a torch.profiler kernel table of 2 steps
    hold = os.environ.get("CRVQA_MPLUG_HOLD_MASKS", "1") != "0"
    torch.manual_seed(49)
    torch.backends.cuda.matmul.allow_tf32 = True      # the UNMASKED torch GEMMs (ViT qkv, heads); the reference runs bf16
    config = dict(image_res=384, vision_width=768, distill=True, clip_name="ViT-B-16",
                  bert_config=dict(stride_layer=3, fusion_layers=6, text_encoder_layers=6, text_decode_layers=12))
    t0 = time.time()
    with torch.device("cuda"):
        model = MPLUG(config=config, tokenizer=types.SimpleNamespace(pad_token_id=0))
    conf = MaskConfigs()
    conf.zero_rate = 0.7
    import contextlib
    import io
    with contextlib.redirect_stdout(io.StringIO()):
        masker = vqa_mplug.init_masker(conf, model)
    n_scores = sum(p.numel() for n, p in model.named_parameters() if p.requires_grad and n.endswith("weight_mask"))
    opt = torch.optim.AdamW([p for p in model.parameters() if p.requires_grad], lr=3e-5, weight_decay=0.02)
    eng = MaskTrainEngine(model, opt, gradient_clipping=1.0, bf16=True, hold_masks=hold)
    model.train()
    g = torch.Generator(device="cuda").manual_seed(49)
    image = torch.randn(B, 3, 384, 384, device="cuda", generator=g)
    q = types.SimpleNamespace(input_ids=torch.randint(1, 30522, (B, 16), device="cuda", generator=g),
                              attention_mask=torch.ones(B, 16, dtype=torch.long, device="cuda"))
    k = [2] * B
    a = types.SimpleNamespace(input_ids=torch.randint(1, 30522, (2 * B, 6), device="cuda", generator=g),
                              attention_mask=torch.ones(2 * B, 6, dtype=torch.long, device="cuda"))
    w = torch.rand(2 * B, device="cuda", generator=g)
    setup_s = time.time() - t0

    # CRVQA_PROBE_SDPA=flash|cudnn|efficient|math pins torch's scaled_dot_product_attention backend for the probe
    # (default: torch's own choice); a measurement aid, the model code does not read it
    pin = os.environ.get("CRVQA_PROBE_SDPA")
    if pin:
        from torch.nn.attention import SDPBackend, sdpa_kernel
        backend = {"flash": SDPBackend.FLASH_ATTENTION, "cudnn": SDPBackend.CUDNN_ATTENTION,
                   "efficient": SDPBackend.EFFICIENT_ATTENTION, "math": SDPBackend.MATH}[pin]
        ctx = sdpa_kernel([backend])
        ctx.__enter__()

    def step():
        loss = eng(image, q, a, train=True, alpha=0.4, k=k, weights=w)
        eng.backward(loss)
        eng.step()
        return loss

    for _ in range(3):
        loss = step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    busy = None
    if profile_to:
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
            for _ in range(2):
                step()
            torch.cuda.synchronize()
        ka = prof.key_averages()
        busy = sum(e.self_device_time_total for e in ka) / 2 / 1e3
        with open(profile_to, "w") as f:
            f.write(f"# mplug_probe batch {B}: {ms:.1f} ms/step unprofiled; GPU-busy {busy:.1f} ms/step (sum of kernel "
                    f"time over 2 profiled steps / 2)\n")
            f.write(ka.table(sort_by="self_cuda_time_total", row_limit=45, max_name_column_width=70))
    with contextlib.redirect_stdout(io.StringIO()):
        t1 = time.time()
        vqa_mplug.update_masks(eng, masker, 0)
        torch.cuda.synchronize()
        upd = time.time() - t1
    print(json.dumps({"workload": "mPLUG-base masked training, 384 px (577 image tokens), 16 question / 6 answer "
                                  "tokens, 2 answers per question, distill twins updated, zero rate 0.7",
                      "batch": B, "steps": steps, "ms_per_step": ms, "hold_masks": hold, "bf16_activations": bool(model.bf16_activations),
                      "gpu_busy_ms_per_step": busy, "samples_per_s": B / ms * 1e3,
                      "fused_optimizer_step": bool(eng._fused and eng._fused.plan is not None
                                                   and eng._fused.plan.uniform_steps),
                      "fused_env": os.environ.get("CRVQA_MPLUG_FUSED", "1"), "sdpa_backend": pin or "torch default",
                      "masked_modules": len([1 for _, m in model.named_modules() if hasattr(m, "threshold")]),
                      "trainable_scores": n_scores, "loss": float(loss), "setup_s": setup_s,
                      "mask_update_s": upd, "max_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))


if __name__ == "__main__":
    main()
