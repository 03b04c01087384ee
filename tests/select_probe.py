"""One batched select over LXMERT-shaped score tensors (for an ncu launch list of the select pipeline)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'compress-robust-vqa_b200')); sys.path.insert(0, ROOT)
import torch
from crvqa import ops
dev = torch.device('cuda')
shapes = [(30522, 768), (768, 2048), (768, 4)] + [(768, 768)] * 125 + [(3072, 768)] * 20 + [(768, 3072)] * 20
g = torch.Generator(device='cuda').manual_seed(1)
scores = [torch.where(torch.rand(s, device=dev, generator=g) < 0.7, 0.0, 0.02) + torch.randn(s, device=dev, generator=g) * 3e-3 for s in shapes]
ks = [max(1, int(s.numel() * 0.7)) for s in scores]
plan = ops.KthPlan(scores)
for _ in range(int(os.environ.get("REPS", "2"))):
    thr = plan(ks)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); thr = plan(ks); e1.record(); torch.cuda.synchronize()
print("select ms", e0.elapsed_time(e1), flush=True)
ref = torch.stack([torch.kthvalue(s.reshape(-1), k).values for s, k in zip(scores, ks)])
print('exact', bool((ref == thr).all()))
import time
t0 = time.perf_counter(); thr = plan(ks); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print('host enqueue ms', (t1 - t0) * 1e3, 'total ms', (t2 - t0) * 1e3)
ties = [torch.where(torch.rand(s, device=dev, generator=g) < 0.7, 0.0, 0.02) for s in shapes]
plan2 = ops.KthPlan(ties); plan2(ks); torch.cuda.synchronize()
e0.record(); thr2 = plan2(ks); e1.record(); torch.cuda.synchronize()
print('ties ms', e0.elapsed_time(e1), 'exact', bool((torch.stack([torch.kthvalue(s.reshape(-1), k).values for s, k in zip(ties, ks)]) == thr2).all()))
