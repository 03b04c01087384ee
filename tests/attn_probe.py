"""One forward + backward of the small-sequence attention at the LXMERT vision shape (for ncu -k regex:attn_)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'compress-robust-vqa_b200'))
import torch
from crvqa import fused
qkv = (torch.randn(256, 36, 2304, device='cuda') * 0.5).bfloat16().requires_grad_(True)
fused.RngState.get(qkv.device).advance()
for _ in range(3):
    o = fused.small_attention(0, 12, None, 0.1, 5, True, qkv)
    o.backward(torch.randn_like(o))
torch.cuda.synchronize()
print('ok')
