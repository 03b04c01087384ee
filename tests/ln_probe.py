"""Forward + backward of the fused dropout + residual + LayerNorm and of the bf16 GELU at the LXMERT vision shape
(M = 9216 rows, H = 768 / 3072 columns), for `ncu --set full -k regex:ln_|gelu_` (round-2 target: ln_bwd sits on the
128-register cap).  Prints achieved GB/s from CUDA events as well.
    python tests/ln_probe.py [M] [H]
"""
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "compress-robust-vqa_b200"))
import torch  # noqa: E402

from crvqa import fused  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 9216
H = int(sys.argv[2]) if len(sys.argv) > 2 else 768
dev = "cuda"
g = (torch.randn(M, H, device=dev) * 0.5).bfloat16().requires_grad_(True)
res = torch.randn(M, H, device=dev, requires_grad=True)
ln = types.SimpleNamespace(weight=torch.rand(H, device=dev) + 0.5, bias=torch.randn(H, device=dev), eps=1e-12)
u = (torch.randn(M, 4 * H, device=dev)).bfloat16().requires_grad_(True)
fused.RngState.get(torch.device(dev)).advance()


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e-3


y32, y16 = fused.drop_add_layernorm(g, res, ln, 0.1, 7, True)
dy32, dy16 = torch.randn_like(y32), torch.randn_like(y16)
t_f = timed(lambda: fused.drop_add_layernorm(g, res, ln, 0.1, 7, True))
t_b = timed(lambda: torch.autograd.grad((y32, y16), (g, res), (dy32, dy16), retain_graph=True))
# algorithmic bytes per element: forward reads g (2) + res (4), writes y32 (4) + y16 (2); backward reads dy32 (4) +
# dy16 (2) + g (2) + res (4), writes dg (2) + dres (4)
print(f"ln_fwd  M={M} H={H}: {t_f * 1e6:.1f} us  {12 * M * H / t_f / 1e9:.0f} GB/s")
print(f"ln_bwd  M={M} H={H}: {t_b * 1e6:.1f} us  {18 * M * H / t_b / 1e9:.0f} GB/s")
yg = fused.gelu_bf16(u)
dyg = torch.randn_like(yg)
t_gf = timed(lambda: fused.gelu_bf16(u))
t_gb = timed(lambda: torch.autograd.grad(yg, u, dyg, retain_graph=True))
print(f"gelu_fwd n={u.numel()}: {t_gf * 1e6:.1f} us  {4 * u.numel() / t_gf / 1e9:.0f} GB/s")
print(f"gelu_bwd n={u.numel()}: {t_gb * 1e6:.1f} us  {6 * u.numel() / t_gb / 1e9:.0f} GB/s")
