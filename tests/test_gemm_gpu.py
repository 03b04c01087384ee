"""Kernel-level parity of the three tcgen05 masked GEMMs against fp32 torch math on the SAME
bf16-rounded operands (so the only difference is fp32 accumulation order; tolerance 2e-3 relative to
the output scale, as BASELINE.json's north_star states)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(M, N, K, seed, rate=0.7):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) * 0.02
    s = torch.rand(N, K, generator=g) * 0.02
    s[torch.rand(N, K, generator=g) < 0.3] = 0.0          # ties at 0, like the magnitude init
    thr = torch.tensor(0.02 * rate * 0.5)
    b = torch.randn(N, generator=g)
    dy = torch.randn(M, N, generator=g)
    dev = torch.device("cuda")
    return [t.to(dev) for t in (x, w, s, thr, b, dy)]


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


SHAPES = [(128, 128, 64), (256, 256, 128), (640, 768, 768), (1152, 3072, 768), (1152, 768, 3072),
          (1152, 768, 2048), (32, 768, 768), (200, 136, 72), (5120, 768, 768)]


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("masked", [True, False])
def test_fwd(M, N, K, masked):
    from crvqa import ops
    x, w, s, thr, b, _ = _mk(M, N, K, 1)
    xb, wb = x.bfloat16(), w.bfloat16()
    mask = (s > thr).float()
    wm = wb.float() * mask if masked else wb.float()
    ref = xb.float() @ wm.t() + b
    y = ops.masked_linear_fwd(xb, wb, s if masked else None, thr, b)
    torch.cuda.synchronize()
    assert _rel(y, ref) < 2e-3, _rel(y, ref)
    yb = ops.masked_linear_fwd(xb, wb, s if masked else None, thr, b, torch.bfloat16)
    assert _rel(yb.float(), ref) < 1e-2


@pytest.mark.parametrize("M,N,K", SHAPES)
@pytest.mark.parametrize("masked", [True, False])
def test_dx(M, N, K, masked):
    from crvqa import ops
    _, w, s, thr, _, dy = _mk(M, N, K, 2)
    dyb, wb = dy.bfloat16(), w.bfloat16()
    mask = (s > thr).float()
    wm = wb.float() * mask if masked else wb.float()
    ref = dyb.float() @ wm
    dx = ops.masked_linear_bwd_dx(dyb, wb, s if masked else None, thr)
    torch.cuda.synchronize()
    assert _rel(dx, ref) < 2e-3, _rel(dx, ref)


@pytest.mark.parametrize("M,N,K", SHAPES)
def test_ds(M, N, K):
    """dS = (dY^T X) (.) W with the UN-ROUNDED fp32 W, as the reference multiplies (masking/maskers.py:337-339 +
    autograd of `self.weight * M_w`): only the MMA operands dY and X are bf16.  A bf16-rounded multiplier would sit
    2^-9 = 2e-3 away element-wise, which the element-wise check below would catch."""
    from crvqa import ops
    x, w, _, _, _, dy = _mk(M, N, K, 3)
    xb, dyb = x.bfloat16(), dy.bfloat16()
    acc = dyb.float().t() @ xb.float()
    ref = acc * w
    ds = ops.masked_linear_bwd_ds(dyb, xb, w)
    torch.cuda.synchronize()
    assert _rel(ds, ref) < 2e-3, _rel(ds, ref)
    # element-wise in units of the element's own magnitude: the multiplier is exact fp32, so what remains is the
    # fp32 summation order of the accumulator (|acc| can cancel to ~0, hence the absolute floor of 1e-3 max|dS|)
    floor = 1e-3 * float(ref.abs().max())
    elem = ((ds - ref).abs() / (ref.abs() + floor)).max()
    assert float(elem) < 5e-4, float(elem)
    rounded = acc * w.bfloat16().float()
    assert float(((rounded - ref).abs() / (ref.abs() + floor)).max()) > 1e-3   # the check can see a bf16 multiplier
    # accumulate: second invocation of a shared module adds
    ds2 = ops.masked_linear_bwd_ds(dyb, xb, w, out=ds.clone(), accumulate=True)
    assert _rel(ds2, 2 * ref) < 2e-3


@pytest.mark.parametrize("N,K", [(768, 768), (3072, 768), (768, 3072), (768, 2048), (256, 128)])
def test_in_kernel_mask_transform_is_bit_exact(N, K):
    """North-star kernel (1) on its fast variant (2-CTA tcgen05 GEMM, scores binarised to a bit mask, transform warps AND
    the W tile in shared memory before the MMA; csrc/gemm_sm100.cu Smem2T<true>): with identity activations the output
    IS the masked weight, so every mask bit's position is checked exactly -- forward Y = I . (W (.) M)^T and
    dX = I . (W (.) M), bf16 products with one non-zero term are exact.  Scores sit at, just below and just above the
    threshold (strict `>`, reference masking/maskers.py:325-339)."""
    from crvqa import ops
    g = torch.Generator().manual_seed(N + K)
    w = (torch.randn(N, K, generator=g) * 0.02).bfloat16().cuda()
    thr = torch.tensor(0.01, device="cuda")
    s = torch.rand(N, K, generator=g) * 0.02
    pick = torch.rand(N, K, generator=g)
    s[pick < 0.2] = 0.01                                    # ties: masked out
    s[(pick >= 0.2) & (pick < 0.3)] = torch.nextafter(torch.tensor(0.01), torch.tensor(1.0))
    s[(pick >= 0.3) & (pick < 0.4)] = torch.nextafter(torch.tensor(0.01), torch.tensor(0.0))
    s = s.cuda()
    wm = torch.where(s > thr, w.float(), torch.zeros((), device="cuda"))
    eye_k = torch.eye(K, device="cuda").bfloat16()          # M = K rows (>= 256 where the 2-CTA variant applies)
    y = ops.masked_linear_fwd(eye_k, w, s, thr, None)
    assert torch.equal(y, wm.t())
    eye_n = torch.eye(N, device="cuda").bfloat16()
    dx = ops.masked_linear_bwd_dx(eye_n, w, s, thr)
    assert torch.equal(dx, wm)
    # ragged M (partial last tile) and a second call with other scores on the same stream (the bit scratch is reused)
    s2 = torch.rand(N, K, generator=g).cuda() * 0.02
    x = torch.randn(300, K, generator=g).bfloat16().cuda()
    y2 = ops.masked_linear_fwd(x, w, s2, thr, None)
    ref2 = x.float() @ torch.where(s2 > thr, w.float(), torch.zeros((), device="cuda")).t()
    assert _rel(y2, ref2) < 2e-3
