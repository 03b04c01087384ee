"""North-star kernel (1) on its 2-CTA variant (scores -> bit mask -> transform warps -> tcgen05.mma.cta_group::2) at
the largest stage-2 shape and at mPLUG's ViT MLP shape, for
    ncu --set full --clock-control none --import-source on -k regex:"masked_gemm2|binarize_bits" python tests/xform_ncu_probe.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'compress-robust-vqa_b200'))
import torch
from crvqa import ops
dev = 'cuda'
for (M, N, K) in [(9216, 3072, 768), (18464, 3072, 768)]:
    x = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.02).bfloat16()
    s = torch.rand(N, K, device=dev); thr = torch.tensor(0.7, device=dev)
    dy = torch.randn(M, N, device=dev).bfloat16(); b = torch.randn(N, device=dev)
    for _ in range(int(os.environ.get('REPS', '3'))):
        ops.masked_linear_fwd(x, w, s, thr, b, torch.bfloat16)
        ops.masked_linear_bwd_dx(dy, w, s, thr, torch.bfloat16)
torch.cuda.synchronize()
print('ok')
