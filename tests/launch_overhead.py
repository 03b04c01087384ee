import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'compress-robust-vqa_b200'))
import torch
from crvqa import ops
dev = 'cuda'
def timeit(fn, iters=200):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3
for (M, N, K) in [(128, 256, 64), (256, 256, 64), (256, 256, 768), (5120, 768, 768)]:
    x = torch.randn(M, K, device=dev).bfloat16(); w = torch.randn(N, K, device=dev).bfloat16()
    g = torch.cuda.CUDAGraph()
    y = ops.masked_linear_fwd(x, w, None, None, None, torch.bfloat16)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): ops.masked_linear_fwd(x, w, None, None, None, torch.bfloat16)
    torch.cuda.synchronize()
    with torch.cuda.graph(g, stream=s):
        for _ in range(50): ops.masked_linear_fwd(x, w, None, None, None, torch.bfloat16)
    t_graph = timeit(lambda: g.replay(), 20) / 50
    t_eager = timeit(lambda: ops.masked_linear_fwd(x, w, None, None, None, torch.bfloat16))
    # alternate with a no-smem kernel (elementwise) to see the reconfiguration cost
    z = torch.zeros(1024, device=dev)
    g2 = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g2, stream=s):
        for _ in range(50):
            ops.masked_linear_fwd(x, w, None, None, None, torch.bfloat16); z.add_(1.0)
    t_alt = timeit(lambda: g2.replay(), 20) / 50
    print(f'{M}x{N}x{K}: per launch eager {t_eager:.1f} us, in-graph back-to-back {t_graph:.1f} us, in-graph alternating with tiny elementwise {t_alt:.1f} us (pair)')
