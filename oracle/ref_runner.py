"""Runs the reference's OWN stage-2 step (the unmodified modules staged in oracle/_ref by oracle/build_ref.py, or
/root/reference when it exists) for bench.py's reference arms.

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- nothing in compress-robust-vqa_b200/ imports it.

One step = what Trainer.train / _training_step do per batch (hg_transformers/mask_trainer_Robust_VQA.py:640-680,
801-886): model.train(); forward; LPF_loss / BCE / LearnedMixin; backward; clip_grad_norm_(1.0); root
optimization.AdamW.step(); model.zero_grad().  Devices: the host CPU (bench.py --impl reference) or cuda:0 under stock
PyTorch (bench.py --impl torch-gpu: fp32 as the reference runs it, or under bf16 autocast) -- none of this
repository's kernels, modules or engine are on that path.
"""
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
STAGED = os.path.join(HERE, "_ref")


def reference_root():
    """Where the reference's modules can be imported from on this machine, or None."""
    for cand in (os.environ.get("CRVQA_REFERENCE_ROOT"), "/root/reference", STAGED):
        if cand and os.path.isdir(os.path.join(cand, "masking")) and os.path.isdir(os.path.join(cand, "hg_transformers")):
            return cand
    return None


def load():
    root = reference_root()
    if root is None:
        raise RuntimeError("reference modules not found (run `python oracle/build_ref.py` where /root/reference exists)")
    os.environ["CRVQA_REFERENCE_ROOT"] = root
    os.environ.setdefault("WANDB_MODE", "disabled")
    os.environ.setdefault("WANDB_SILENT", "true")
    gold = os.path.join(ROOT, "tests", "golden")
    if gold not in sys.path:
        sys.path.insert(0, gold)
    # the product package mirrors the reference's module names (masking, hg_transformers, optimization): make sure
    # the reference's own resolve first in this process
    for name in [n for n in sys.modules if n.split(".")[0] in ("masking", "hg_transformers", "optimization", "utils")]:
        del sys.modules[name]
    pkg = os.path.join(ROOT, "compress-robust-vqa_b200")
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != pkg]
    import ref_shims
    ref_shims.REF_ROOT = root
    import make_golden as mg
    return mg, mg.load_reference(), root


class ReferenceStage2:
    """LXMERT 9/5/5 + reference Masker (rates 0.3/0.3/0.3, zero rate 0.7, magnitude init) + reference AdamW."""

    def __init__(self, ans_num, device="cpu", seed=49, lr=5e-5):
        import torch
        self.torch = torch
        self.mg, self.R, self.root = load()
        R = self.R
        torch.manual_seed(seed)
        self.model = R.lx.LxmertForMultipleChoice(R.cfg.LxmertConfig(ans_num=ans_num))
        self.masker = self.mg.make_masker(R, self.model)        # on the CPU, as the reference driver patches
        self.device = torch.device(device)
        self.model.to(self.device)
        params = [p for _, p in self.model.named_parameters() if p.requires_grad]
        self.opt = R.optim.AdamW([{"params": [p]} for p in params], lr=lr, eps=1e-8)
        self.ans_num = ans_num

    def batch(self, B, seed=49):
        b = self.mg.synthetic_batch(B, self.ans_num, seed=seed)
        return {k: v.to(self.device) for k, v in b.items()}

    def step(self, b, loss_kind="lpf", autocast_bf16=False):
        torch, R, model = self.torch, self.R, self.model
        model.train()
        ctx = torch.autocast("cuda", dtype=torch.bfloat16) if autocast_bf16 else _null()
        with ctx:
            loss, logits, pooled = model(b["ids"], b["feats"], b["pos"], labels=b["target"])
            if loss_kind == "lpf":
                loss = R.trainer.LPF_loss(logits, b["bias"], b["max_label"], self.device, 5)
            elif loss_kind == "lmh":
                if not hasattr(self, "lmh"):
                    self.lmh = R.loss.LearnedMixin(0.36).to(self.device)
                loss = self.lmh(pooled, logits, b["bias"], b["target"], self.device)
        loss.backward()
        torch.nn.utils.clip_grad_norm_(model.parameters(), 1.0)
        self.opt.step()
        model.zero_grad()
        return loss.detach()

    def timed(self, B, steps, warmup, loss_kind="lpf", autocast_bf16=False):
        torch = self.torch
        b = self.batch(B)
        cuda = self.device.type == "cuda"
        for _ in range(warmup):
            self.step(b, loss_kind, autocast_bf16)
        if cuda:
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        t0 = time.perf_counter()
        for _ in range(steps):
            loss = self.step(b, loss_kind, autocast_bf16)
        if cuda:
            e1.record()
            torch.cuda.synchronize()
            dt = e0.elapsed_time(e1) * 1e-3
        else:
            dt = time.perf_counter() - t0
        return dt, float(loss)


class _null:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False
