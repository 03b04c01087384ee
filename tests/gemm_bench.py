"""Micro-benchmark of the masked-GEMM family at the BASELINE shapes (B=256): TFLOP/s per variant."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'compress-robust-vqa_b200'))
import torch
from crvqa import ops
dev = 'cuda'
torch.manual_seed(0)
SHAPES = [(9216, 768, 768), (5120, 768, 768), (9216, 3072, 768), (9216, 768, 3072), (9216, 768, 2048)]
if len(sys.argv) > 1 and sys.argv[1] == 'mplug':   # mPLUG-base, batch 32: ViT MLPs (577 tokens), cross-attention K/V, connected fusion layer
    SHAPES = [(18464, 3072, 768), (18464, 768, 3072), (37952, 768, 768), (18976, 768, 768)]
elif len(sys.argv) > 1: SHAPES = SHAPES[:int(sys.argv[1])]
def timeit(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for (M, N, K) in SHAPES:
    x = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.02).bfloat16(); w32 = w.float()
    s = torch.rand(N, K, device=dev); thr = torch.tensor(0.7, device=dev); dy = torch.randn(M, N, device=dev).bfloat16()
    b = torch.randn(N, device=dev); ds = torch.zeros(N, K, device=dev)
    fl = 2.0 * M * N * K
    res = {}
    res['fwd masked f32'] = timeit(lambda: ops.masked_linear_fwd(x, w, s, thr, b))
    res['fwd masked bf16'] = timeit(lambda: ops.masked_linear_fwd(x, w, s, thr, b, torch.bfloat16))
    res['fwd plain f32'] = timeit(lambda: ops.masked_linear_fwd(x, w, None, thr, b))
    res['fwd plain bf16'] = timeit(lambda: ops.masked_linear_fwd(x, w, None, thr, b, torch.bfloat16))
    res['dx masked f32'] = timeit(lambda: ops.masked_linear_bwd_dx(dy, w, s, thr))
    res['dx masked bf16'] = timeit(lambda: ops.masked_linear_bwd_dx(dy, w, s, thr, torch.bfloat16))
    res['dx plain f32'] = timeit(lambda: ops.masked_linear_bwd_dx(dy, w, None, thr))
    res['dx plain bf16'] = timeit(lambda: ops.masked_linear_bwd_dx(dy, w, None, thr, torch.bfloat16))
    res['ds store'] = timeit(lambda: ops.masked_linear_bwd_ds(dy, x, w32, out=ds, accumulate=False))
    res['ds accum'] = timeit(lambda: ops.masked_linear_bwd_ds(dy, x, w32, out=ds, accumulate=True))
    res['torch bf16 mm'] = timeit(lambda: torch.matmul(x, w.t()))
    print(f'M={M} N={N} K={K}: ' + ' | '.join(f'{k} {v*1e3:.0f}us {fl/v/1e9:.0f}TF' for k, v in res.items()), flush=True)
