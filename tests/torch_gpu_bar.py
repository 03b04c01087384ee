"""Same-box bar: the reference's stage-2 step as STOCK PyTorch executes it on this GPU -- the
oracle's restatement of the reference formula (binarise -> W * M -> F.linear per call, autograd backward, clip, AdamW
loop over tensors) run with torch's own CUDA kernels, fp32 as the reference, and again with TF32 matmuls allowed.
Not a test and not the product: prints one JSON line to compare bench.py's `value` with.
    python tests/torch_gpu_bar.py [batch] [steps] [device]
"""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "compress-robust-vqa_b200"))
sys.path.insert(0, ROOT)

RATES = {"Lang": 0.7, "Vis": 0.7, "Fus": 0.7, "P": 0.7}


def main():
    from hg_transformers.modeling_lxmert import LxmertConfig, LxmertForMultipleChoice
    from oracle import lxmert_oracle as lxo
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    dev = torch.device(sys.argv[3] if len(sys.argv) > 3 else "cuda")
    A = 3129
    torch.manual_seed(49)
    model = LxmertForMultipleChoice(LxmertConfig(ans_num=A))
    params = {k: v.detach().to(dev) for k, v in model.state_dict().items()}
    del model
    for k in params:
        params[k].requires_grad_(k.startswith("classifier."))
    scores, thr, modal = {}, {}, {}
    for name, m in lxo.module_names():                       # magnitude init with torch on the device
        w = params[name + ".weight"]
        k = int(w.numel() * RATES[m])
        cut = w.abs().flatten().kthvalue(k).values
        scores[name] = torch.where(w.abs() > cut, 0.02, 0.0).to(w.dtype).requires_grad_(True)
        thr[name] = 1e-2
    ctx = lxo.Ctx(params, scores, thr, operand="fp32", train=True)
    data = {k: v.to(dev) for k, v in lxo.synthetic_batch(B, A).items()}
    out = {"workload": f"LXMERT stage-2 lpf step, batch {B}, A={A}, stock PyTorch kernels on {dev}", "batch": B}
    for label, tf32 in (("fp32", False), ("tf32_matmul", True)):
        torch.backends.cuda.matmul.allow_tf32 = tf32
        opt_state = {}
        for _ in range(2):
            lxo.training_step(ctx, data, "lpf", opt_state=opt_state)
        if dev.type == "cuda":
            torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(steps):
            r = lxo.training_step(ctx, data, "lpf", opt_state=opt_state)
        if dev.type == "cuda":
            torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) / steps * 1e3
        out[label] = {"ms_per_step": ms, "samples_per_s": B / ms * 1e3, "loss": float(r["loss"])}
    if dev.type == "cuda":
        out["max_mem_gb"] = torch.cuda.max_memory_allocated() / 2 ** 30
    print(json.dumps(out))


if __name__ == "__main__":
    main()
