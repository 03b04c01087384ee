"""In-graph timing of the masked-GEMM family at the LXMERT step's shapes (batch 256): one launch per GEMM (round-1
path) against grouped launches.  Every variant is captured as a CUDA graph of REPS back-to-back launches (no host
time between them, programmatic dependent launch as in the step) and timed with CUDA events.

    python tests/gemm_group_probe.py [out.json]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "compress-robust-vqa_b200"))
import torch  # noqa: E402

from crvqa import ops  # noqa: E402

dev = "cuda"
REPS = 20
torch.manual_seed(0)


def graph_time(fn, reps=REPS):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (3 * reps) * 1e3   # us per fn()


class Mod:
    def __init__(self, M, N, K):
        self.M, self.N, self.K = M, N, K
        self.x = torch.randn(M, K, device=dev).bfloat16()
        self.w32 = torch.randn(N, K, device=dev) * 0.02
        self.w = self.w32.bfloat16()
        self.b = torch.randn(N, device=dev)
        self.dy = torch.randn(M, N, device=dev).bfloat16()
        self.y = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        self.u = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        self.dx = torch.empty(M, K, dtype=torch.bfloat16, device=dev)
        self.ds = torch.zeros(N, K, device=dev)
        self.flop = 2.0 * M * N * K

    def fwd(self):
        return ops.gemm_problem(ops.GEMM_FWD, self.x, self.w, self.y, bias=self.b)

    def fwd_gelu(self):
        return ops.gemm_problem(ops.GEMM_FWD, self.x, self.w, self.y, bias=self.b, aux=self.u, act=ops.ACT_GELU)

    def dxp(self, u=None):
        return ops.gemm_problem(ops.GEMM_DX, self.dy, self.w, self.dx, aux=u, act=ops.ACT_GELU if u is not None else 0)

    def dsp(self):
        return ops.gemm_problem(ops.GEMM_DS, self.dy, self.x, self.ds, w_f32=self.w32)


out = {}


def rec(name, us, flop):
    out[name] = {"us": us, "tflops": flop / us / 1e6}
    print(f"{name:58s} {us:8.1f} us  {flop / us / 1e6:7.0f} TFLOP/s", flush=True)


for tag, Ml, Mv in (("B256", 5120, 9216),):
    for N, K in ((768, 768), (2304, 768), (3072, 768), (768, 3072)):
        l, v = Mod(Ml, N, K), Mod(Mv, N, K)
        for side, m in (("lang", l), ("visn", v)):
            t_f = graph_time(lambda: ops.masked_linear_fwd(m.x, m.w, None, None, m.b, torch.bfloat16))
            t_x = graph_time(lambda: ops.masked_linear_bwd_dx(m.dy, m.w, None, None, torch.bfloat16))
            t_s = graph_time(lambda: ops.masked_linear_bwd_ds(m.dy, m.x, m.w32, out=m.ds, accumulate=False))
            rec(f"{side} {m.M}x{N}x{K} single fwd", t_f, m.flop)
            rec(f"{side} {m.M}x{N}x{K} single dx", t_x, m.flop)
            rec(f"{side} {m.M}x{N}x{K} single ds", t_s, m.flop)
            rec(f"{side} {m.M}x{N}x{K} single dx+ds (sum)", t_x + t_s, 2 * m.flop)
            pf, pxs = [m.fwd()], [m.dxp(), m.dsp()]
            rec(f"{side} {m.M}x{N}x{K} grouped fwd(1)", graph_time(lambda: ops.gemm_grouped(pf)), m.flop)
            rec(f"{side} {m.M}x{N}x{K} grouped dx+ds", graph_time(lambda: ops.gemm_grouped(pxs)), 2 * m.flop)
        pf2 = [l.fwd(), v.fwd()]
        rec(f"both  {N}x{K} grouped fwd lang+visn", graph_time(lambda: ops.gemm_grouped(pf2)), l.flop + v.flop)
        pb4 = [l.dxp(), l.dsp(), v.dxp(), v.dsp()]
        rec(f"both  {N}x{K} grouped dx+ds lang+visn", graph_time(lambda: ops.gemm_grouped(pb4)), 2 * (l.flop + v.flop))
        if N == 3072:
            pg = [l.fwd_gelu(), v.fwd_gelu()]
            rec(f"both  {N}x{K} grouped fwd+GELU lang+visn", graph_time(lambda: ops.gemm_grouped(pg)), l.flop + v.flop)
        if K == 3072:
            ul = torch.randn(Ml, K, device=dev).bfloat16()
            uv = torch.randn(Mv, K, device=dev).bfloat16()
            pgg = [l.dxp(ul), l.dsp(), v.dxp(uv), v.dsp()]
            rec(f"both  {N}x{K} grouped dx*gelu'+ds lang+visn", graph_time(lambda: ops.gemm_grouped(pgg)),
                2 * (l.flop + v.flop))
        del l, v
        torch.cuda.empty_cache()

if len(sys.argv) > 1:
    with open(sys.argv[1], "w") as f:
        json.dump(out, f, indent=1)
