"""Per-CTA phase timestamps of the GEMM kernel (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'compress-robust-vqa_b200'))
import ctypes, torch
from crvqa import ops, lib
dev = 'cuda'
dbg = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
def run(name, fn):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    lib.crv_gemm_debug_timestamps(ctypes.c_void_p(dbg.data_ptr()))
    dbg.zero_(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    lib.crv_gemm_debug_timestamps(ctypes.c_void_p(0))
    d = dbg.view(148, 8).cpu()
    d = d[d[:, 0] > 0]
    t0 = d[:, 0].min()
    entry = (d[:, 7] - t0).float() / 1000.0
    rel = (d[:, :6] - t0).float() / 1000.0
    print(f'{name}: entry {entry.mean():.2f} (min {entry.min():.2f}) event {e0.elapsed_time(e1)*1e3:.1f}us ctas={len(d)} tiles/cta={d[:,6].float().mean():.2f} | setup_done {rel[:,0].mean():.2f} (max {rel[:,0].max():.2f}) first_tma {rel[:,1].mean():.2f} stage0_landed {rel[:,2].mean():.2f} acc0_ready {rel[:,3].mean():.2f} epi0_drained {rel[:,4].mean():.2f} all_done {rel[:,5].mean():.2f} (max {rel[:,5].max():.2f}) us')
import os
for (M, N, K) in [(5120, 768, 768), (9216, 768, 768)]:
    x = torch.randn(M, K, device=dev).bfloat16(); w = (torch.randn(N, K, device=dev) * 0.02).bfloat16(); w32 = w.float()
    dy = torch.randn(M, N, device=dev).bfloat16(); b = torch.randn(N, device=dev); ds = torch.zeros(N, K, device=dev)
    run(f'fwd bf16 {M}x{N}x{K}', lambda: ops.masked_linear_fwd(x, w, None, None, b, torch.bfloat16))
    run(f'fwd nobias {M}x{N}x{K}', lambda: ops.masked_linear_fwd(x, w, None, None, None, torch.bfloat16))
    run(f'dx  bf16 {M}x{N}x{K}', lambda: ops.masked_linear_bwd_dx(dy, w, None, None, torch.bfloat16))
    run(f'ds       {M}x{N}x{K}', lambda: ops.masked_linear_bwd_ds(dy, x, w32, out=ds, accumulate=False))
