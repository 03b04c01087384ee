"""The three GEMM instantiations of the step at their largest stage-2 shape, for `ncu --set full -k regex:masked_gemm2`."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'compress-robust-vqa_b200'))
import torch
from crvqa import ops
dev = 'cuda'
M, N, K = 9216, 3072, 768
x = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.02).bfloat16(); w32 = w.float()
dy = torch.randn(M, N, device=dev).bfloat16(); b = torch.randn(N, device=dev); ds = torch.zeros(N, K, device=dev)
for _ in range(int(os.environ.get('REPS', '4'))):
    ops.masked_linear_fwd(x, w, None, None, b, torch.bfloat16)
    ops.masked_linear_bwd_dx(dy, w, None, None, torch.bfloat16)
    ops.masked_linear_bwd_ds(dy, x, w32, out=ds, accumulate=False)
torch.cuda.synchronize()
print('ok')
