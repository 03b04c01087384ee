"""Torch-facing wrappers of the C ABI: raw ops on CUDA tensors + the autograd Functions that replace
``_Binarizer1`` + ``weight * M_w`` + ``F.linear`` / ``F.embedding`` (masking/maskers.py:325-366) and the
loss graphs of the trainers.  Every op runs on torch's current CUDA stream."""
import ctypes
import math
import os

import torch

from ._lib import GemmProblem, check, lib

DT_F32, DT_BF16 = 0, 1

# bench.py sets this to a list to collect (kind, M, N, K, start_event, end_event) for every masked-GEMM
# launch (CUDA events on the launching stream); None = no instrumentation.
PROFILE = None
# bench.py sets this to a list to collect re-issuable (kind, M, N, K, thunk) records of every masked-GEMM
# launch of a step, so the GEMM family can be replayed back to back (as one CUDA graph) and timed alone.
RECORD = None


class _Timed:
    def __init__(self, kind, M, N, K):
        self.rec = None
        if PROFILE is not None:
            self.rec = (kind, M, N, K, torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))

    def __enter__(self):
        if self.rec is not None:
            self.rec[4].record()

    def __exit__(self, *exc):
        if self.rec is not None:
            self.rec[5].record()
            PROFILE.append(self.rec)


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _launch(kind, M, N, K, name, fn, *args):
    """One library GEMM call; the CUDA-event bracket only when bench.py asked for per-launch timings."""
    if PROFILE is None:
        check(fn(*args), name)
    else:
        with _Timed(kind, M, N, K):
            check(fn(*args), name)


def _stream():
    """torch's current CUDA stream as an integer handle (every entry point declares void* argtypes, so ctypes takes
    plain ints / None: no c_void_p objects on the per-launch path -- an eager mPLUG step makes ~1 100 library calls)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _p(t):
    return t.data_ptr() if t is not None else None


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("crvqa ops need CUDA tensors: the masked kernels have no CPU fallback")


def _stage(t):
    """Set-up-time ops (magnitude init, mask export, threshold refresh) accept CPU tensors when a GPU
    exists -- the reference patches the model on the CPU and moves it afterwards -- by staging them on
    the current CUDA device.  There is still no CPU implementation: without a GPU this raises."""
    if t.is_cuda:
        return t
    if not torch.cuda.is_available():
        raise RuntimeError("crvqa ops need a CUDA device: the masked kernels have no CPU fallback")
    return t.cuda()


def to_bf16(x):
    """fp32 -> bf16 (RNE) with the library's cast kernel; bf16 input is returned as is."""
    if x.dtype == torch.bfloat16:
        return x.contiguous()
    _need_cuda(x)
    x = x.contiguous()
    if x.dtype != torch.float32:
        x = x.float()
    out = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
    check(lib.crv_cast_f32_to_bf16(_p(x), _p(out), x.numel(), _stream()), "crv_cast_f32_to_bf16")
    return out


def mul_cast_bf16(w, mask, out=None):
    """bf16(w * mask), product in fp32 (stage-3 pruned operand, crv_mul_cast_bf16)."""
    _need_cuda(w)
    w = w.detach().contiguous()
    mask = mask.contiguous()
    if w.dtype != torch.float32 or mask.dtype != torch.float32 or w.shape != mask.shape:
        raise ValueError("mul_cast_bf16 needs fp32 weight and fp32 0/1 mask of the same shape")
    if out is None:
        out = torch.empty(w.shape, dtype=torch.bfloat16, device=w.device)
    elif out.dtype != torch.bfloat16 or out.numel() != w.numel() or not out.is_contiguous():
        raise ValueError("mul_cast_bf16: out must be a contiguous bf16 tensor of the same size")
    check(lib.crv_mul_cast_bf16(_p(w), _p(mask), _p(out), w.numel(), _stream()), "crv_mul_cast_bf16")
    return out


def as_thr(thr, device):
    """The reference keeps thresholds as 0-dim tensors (possibly on CPU); the kernels read a device float."""
    if not torch.is_tensor(thr):
        return torch.tensor(float(thr), dtype=torch.float32, device=device)
    if thr.device != device or thr.dtype != torch.float32:
        thr = thr.to(device=device, dtype=torch.float32)
    return thr


def binarize(scores, thr, want_count=False, as_bool=False):
    """binarizer_fn1: 1.0 where scores > thr else 0.0 (masking/maskers.py:325-329)."""
    home = scores.device
    s = _stage(scores.detach()).contiguous().float()
    thr = as_thr(thr, s.device)
    mf = None if as_bool else torch.empty_like(s)
    mb = torch.empty(s.shape, dtype=torch.uint8, device=s.device) if as_bool else None
    cnt = torch.zeros((), dtype=torch.int64, device=s.device) if want_count else None
    check(lib.crv_binarize(_p(s), _p(thr), _p(mf), _p(mb), _p(cnt), s.numel(), _stream()), "crv_binarize")
    out = (mb.view(torch.bool) if as_bool else mf).to(home)
    return (out, cnt) if want_count else out


def apply_mask_bf16(w_bf16, scores, thr):
    _need_cuda(w_bf16, scores)
    out = torch.empty_like(w_bf16)
    thr = as_thr(thr, scores.device)
    check(lib.crv_apply_mask_bf16(_p(w_bf16), _p(scores), _p(thr), _p(out), w_bf16.numel(), _stream()),
          "crv_apply_mask_bf16")
    return out


def apply_mask_segmented(w_flat, s_flat, thr_vec, chunks, wm_flat):
    """Refresh a whole mask cache: wm = w (.) (s > thr_vec[segment]) in one launch (chunks: int32 [n,4])."""
    _need_cuda(w_flat, s_flat, thr_vec, chunks, wm_flat)
    check(lib.crv_apply_mask_segmented(_p(w_flat), _p(s_flat), _p(thr_vec), _p(chunks), chunks.shape[0],
                                       _p(wm_flat), _stream()), "crv_apply_mask_segmented")


def masked_linear_fwd(x_bf16, w_bf16, scores, thr, bias, out_dtype=torch.float32):
    """Y[M,N] = X[M,K] . (W (.) (S > thr))^T + b; scores None => W taken as is."""
    _need_cuda(x_bf16, w_bf16)
    M, K = x_bf16.shape
    N = w_bf16.shape[0]
    y = torch.empty((M, N), dtype=out_dtype, device=x_bf16.device)
    thr_t = as_thr(thr, x_bf16.device) if scores is not None else None
    if RECORD is not None:
        RECORD.append(("fwd", M, N, K, lambda: masked_linear_fwd(x_bf16, w_bf16, scores, thr, bias, out_dtype)))
    _launch("fwd", M, N, K, "crv_masked_linear_fwd", lib.crv_masked_linear_fwd, _p(x_bf16), _p(w_bf16), _p(scores),
            _p(thr_t), _p(bias), _p(y), DT_F32 if out_dtype == torch.float32 else DT_BF16, M, N, K, _stream())
    return y


def masked_linear_bwd_dx(dy_bf16, w_bf16, scores, thr, out_dtype=torch.float32):
    _need_cuda(dy_bf16, w_bf16)
    M, N = dy_bf16.shape
    K = w_bf16.shape[1]
    dx = torch.empty((M, K), dtype=out_dtype, device=dy_bf16.device)
    thr_t = as_thr(thr, dy_bf16.device) if scores is not None else None
    if RECORD is not None:
        RECORD.append(("dx", M, N, K, lambda: masked_linear_bwd_dx(dy_bf16, w_bf16, scores, thr, out_dtype)))
    _launch("dx", M, N, K, "crv_masked_linear_bwd_dx", lib.crv_masked_linear_bwd_dx, _p(dy_bf16), _p(w_bf16),
            _p(scores), _p(thr_t), _p(dx), DT_F32 if out_dtype == torch.float32 else DT_BF16, M, N, K, _stream())
    return dx


def masked_linear_bwd_ds(dy_bf16, x_bf16, w_f32, out=None, accumulate=False):
    """dS[N,K] (+)= (dY^T X) (.) W with W in fp32 (the reference multiplies by its fp32 Parameter)."""
    _need_cuda(dy_bf16, x_bf16, w_f32)
    M, N = dy_bf16.shape
    K = x_bf16.shape[1]
    if w_f32.dtype != torch.float32 or not w_f32.is_contiguous() or w_f32.shape != (N, K):
        raise ValueError("masked_linear_bwd_ds multiplies by a contiguous fp32 [N, K] tensor")
    if out is None:
        out = torch.empty((N, K), dtype=torch.float32, device=dy_bf16.device)
        accumulate = False
    if RECORD is not None:
        scratch = torch.empty_like(out)
        RECORD.append(("ds", M, N, K, lambda: masked_linear_bwd_ds(dy_bf16, x_bf16, w_f32, out=scratch,
                                                                     accumulate=accumulate)))
    _launch("ds", M, N, K, "crv_masked_linear_bwd_ds", lib.crv_masked_linear_bwd_ds, _p(dy_bf16), _p(x_bf16),
            _p(w_f32), _p(out), int(accumulate), M, N, K, _stream())
    return out


GEMM_FWD, GEMM_DX, GEMM_DS = 0, 1, 2
ACT_NONE, ACT_GELU = 0, 1


def grouped_2cta_ok(MM, NN):
    """Shapes the grouped 2-CTA kernel takes (generic output extents MM x NN): whole 256-column tiles, >= 256 rows."""
    return NN % 256 == 0 and MM >= 256 and os.environ.get("CRVQA_2CTA", "1") != "0"


def gemm_problem(kind, a, b, out, *, bias=None, w_f32=None, aux=None, act=ACT_NONE, accumulate=False):
    """One entry of a grouped launch.  kind FWD: a = X [M,K], b = Wm [N,K]; DX: a = dY [M,N], b = Wm [N,K];
    DS: a = dY [M,N], b = X [M,K], w_f32 = fp32 multiplier [N,K].  Returns (struct, flop, keep-alive tensors)."""
    if kind == GEMM_FWD:
        (M, K), N = a.shape, b.shape[0]
    elif kind == GEMM_DX:
        (M, N), K = a.shape, b.shape[1]
    else:
        (M, N), K = a.shape, b.shape[1]
    pr = GemmProblem(kind, act, DT_BF16 if out.dtype == torch.bfloat16 else DT_F32, int(accumulate), M, N, K, 0,
                     a.data_ptr(), b.data_ptr(), bias.data_ptr() if bias is not None else None,
                     w_f32.data_ptr() if w_f32 is not None else None, out.data_ptr(),
                     aux.data_ptr() if aux is not None else None)
    return pr, ("fwd", "dx", "ds")[kind], M, N, K, (a, b, out, bias, w_f32, aux)


def gemm_grouped(problems):
    """Launch a list of gemm_problem() entries as grouped 2-CTA kernels (crv_masked_gemm_grouped)."""
    n = len(problems)
    arr = (GemmProblem * n)(*[p[0] for p in problems])
    if RECORD is not None:
        for p in problems:
            RECORD.append((p[1], p[2], p[3], p[4], None))
        keep = [p[5] for p in problems]          # the replay reads / writes these buffers: keep them allocated
        RECORD.append(("group", 0, 0, 0, lambda: (keep, check(lib.crv_masked_gemm_grouped(arr, n, _stream()),
                                                              "crv_masked_gemm_grouped"))))
    check(lib.crv_masked_gemm_grouped(arr, n, _stream()), "crv_masked_gemm_grouped")


_kth_ws = {}


class KthPlan:
    """The fixed half of a batched select -- segment pointers, sizes and the workspace -- built once for a set of
    tensors whose storage does not move (the score arena), so a reset_threshold call only marshals the ranks."""

    def __init__(self, tensors):
        self.count = len(tensors)
        if self.count == 0:
            raise ValueError("no segments")
        self.home = tensors[0].device
        staged = [_stage(t.detach()) for t in tensors]
        self.dev = staged[0].device
        self.flat = [t.reshape(-1) for t in staged]   # keeps the storages alive
        for t in self.flat:
            if t.dtype != torch.float32 or not t.is_contiguous():
                raise ValueError("kth_value_batched needs contiguous float32 tensors")
        self.sizes = [t.numel() for t in self.flat]
        self.ptrs = (ctypes.c_void_p * self.count)(*[t.data_ptr() for t in self.flat])
        self.ns = (ctypes.c_longlong * self.count)(*self.sizes)
        self.nbytes = lib.crv_kth_value_workspace_bytes_for(self.ns, self.count)
        key = (self.dev, self.count)
        ws = _kth_ws.get(key)
        if ws is None or ws.numel() < self.nbytes:
            ws = torch.empty(self.nbytes, dtype=torch.uint8, device=self.dev)
            _kth_ws[key] = ws
        self.ws = ws

    def __call__(self, ks, use_abs=False):
        kk = (ctypes.c_longlong * self.count)(*ks)
        out = torch.empty(self.count, dtype=torch.float32, device=self.dev)
        check(lib.crv_kth_value_batched(self.ptrs, self.ns, kk, self.count, int(bool(use_abs)), _p(out), _p(self.ws),
                                        self.nbytes, _stream()), "crv_kth_value_batched")
        return out if self.home == self.dev else out.to(self.home)


def kth_value_batched(tensors, ks, use_abs=False):
    """Exact k-th smallest (1-based) of each fp32 CUDA tensor; returns a float32 device vector."""
    return KthPlan(tensors)([int(k) for k in ks], use_abs)


def magnitude_init(weight, w_thr, hi, lo):
    home = weight.device
    w = _stage(weight.detach()).contiguous()
    out = torch.empty_like(w)
    check(lib.crv_magnitude_init(_p(w), _p(as_thr(w_thr, w.device)), float(hi), float(lo), _p(out), w.numel(),
                                 _stream()), "crv_magnitude_init")
    return out.to(home)


def _sink_grad(sink):
    """Arena gradient view of a masked module (hg_transformers._engine.ScoreArena), or None."""
    return getattr(sink, "_arena_grad", None) if sink is not None else None


DS_OVERWRITE, DS_ADD, DS_ZEROED = 0, 1, 3


def ds_mode(*modules):
    """`accumulate` of the next score-gradient GEMM into these modules' arena gradient: add when an earlier invocation
    of this step already wrote it, ZEROED when the optimiser pass cleared it (no memset before a split reduction),
    overwrite otherwise.  Consumes the zeroed state."""
    if any(m._grad_dirty for m in modules):
        return DS_ADD
    zeroed = all(getattr(m, "_grad_zero", False) for m in modules)
    for m in modules:
        m._grad_zero = False
    return DS_ZEROED if zeroed else DS_OVERWRITE


def _sink_skipped(sink):
    """A forward invocation of `sink` whose output nothing read: its backward is owed no gradient, but the bucket
    bookkeeping of the gradient exchange still has to see the invocation finish."""
    sync = getattr(sink, "_sync", None)
    if sync is not None:
        sync.module_backward_done(sink)


def _sink_done(sink):
    sink._grad_dirty = True
    sync = getattr(sink, "_sync", None)
    if sync is not None:
        sync.module_backward_done(sink)


class _DsLane:
    """Second stream for the score-gradient GEMMs of the backward pass.

    dX feeds the next layer's backward; dS is only read by the gradient exchange / optimiser at the end of the
    step.  The stage-2 GEMMs are short (one wave of 256 x 256 tiles, 10 - 40 us) and more than half of such a
    launch is pipeline fill, epilogue and drain, so dS runs on its own stream and its fixed costs overlap the dX
    chain instead of extending it.  fork() orders the lane after dY's producer; the operands are held until
    join(), which the arena calls before anything reads the gradients (ScoreArena.finalize_grads,
    GradSync._launch).  Inside a CUDA-graph capture this is an ordinary fork / join of the capturing stream.
    CRVQA_DS_STREAM=0 keeps everything on one stream."""

    def __init__(self, device):
        self.device = device
        self.stream = torch.cuda.Stream(device)
        self.held = []
        self.open = False

    def fork(self):
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        self.open = True

    def hold(self, *tensors):
        self.held.extend(tensors)
        if len(self.held) > 4096:        # nobody joined for ~10 backward passes (arena gradients read without
            self.join()                  # finalize_grads): join here rather than pin activations without bound

    def join(self):
        if self.open:
            torch.cuda.current_stream(self.device).wait_stream(self.stream)
            self.open = False
        self.held.clear()


_ds_lanes = {}


def ds_lane(device):
    """The dS lane of `device`, or None when disabled."""
    if os.environ.get("CRVQA_DS_STREAM", "1") == "0":
        return None
    key = torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()
    lane = _ds_lanes.get(key)
    if lane is None:
        lane = _ds_lanes[key] = _DsLane(torch.device("cuda", key))
    return lane


def ds_lane_join():
    """Make the current stream wait for every outstanding dS launch (cheap no-op when none is)."""
    for lane in _ds_lanes.values():
        lane.join()


def operand_mode():
    """'bf16' (product path) or 'split': every MMA operand is carried as hi + lo bf16 halves and each GEMM is issued
    three times (hi.hi + lo.hi + hi.lo, fp32 accumulate), i.e. ~16 mantissa bits per operand through the SAME tcgen05
    kernels.  A debugging / parity aid (CRVQA_OPERAND=split): it shows how much of the end-to-end gap to the fp32
    reference is bf16 operand rounding and how much is the kernels (tests/test_parity_precise_gpu.py)."""
    return os.environ.get("CRVQA_OPERAND", "bf16")


def split_bf16(x32):
    """x ~= hi + lo with hi = bf16(x), lo = bf16(x - hi): relative error ~2^-17."""
    hi = to_bf16(x32)
    lo = to_bf16(x32 - hi.float())
    return hi, lo


class MaskedLinearFn(torch.autograd.Function):
    """y = F.linear(x, weight * binarize(scores, thr), bias) with the straight-through score gradient.

    forward : bf16 tcgen05 GEMM with the mask applied in the TMA-fed prologue (fp32 accumulate)
    backward: dX = dY . (W (.) M)   and   dS = (dY^T . X) (.) W   (no dW, no db: weights are frozen,
              masking/maskers.py:564-569,594-596); the (.) W of dS uses the fp32 weight
    `sink` is the owning module when its score gradient lives in a ScoreArena: dS is then accumulated
    in place by the GEMM epilogue and autograd receives no tensor for `scores`.
    """

    @staticmethod
    def forward(ctx, x, scores, w_bf16, thr, bias, sink=None, wm_bf16=None, w_f32=None):
        shp = x.shape
        thr_t = as_thr(thr, x.device)
        if w_f32 is None:
            w_f32 = w_bf16.float()
        if w_f32.requires_grad:
            w_f32 = w_f32.detach()
        ctx.split = operand_mode() == "split" and x.dtype == torch.float32
        if ctx.split:
            return MaskedLinearFn._forward_split(ctx, x, scores, w_bf16, thr_t, bias, sink, wm_bf16, w_f32)
        # a producer that already wrote the bf16 copy of x in the same launch (crvqa.fused.drop_add_layernorm /
        # ln_avg_drop attach it as x._crv_bf16) saves the cast here; gradients still flow through x
        x16 = getattr(x, "_crv_bf16", None)
        if x16 is not None and x16.shape == shp and x16.dtype == torch.bfloat16 and x16.is_contiguous():
            x2 = x16.reshape(-1, shp[-1])
        else:
            x2 = to_bf16(x.reshape(-1, shp[-1]))
        # bf16 activations (a bf16 input, or a caller running under bf16 autocast, where nn.Linear would answer in bf16
        # too) -> bf16 out; dX always comes back in the input's own dtype: no fp32 round trips between bf16 layers
        autocast16 = (x.is_cuda and torch.is_autocast_enabled("cuda")
                      and torch.get_autocast_dtype("cuda") == torch.bfloat16)
        io = torch.bfloat16 if (x.dtype == torch.bfloat16 or autocast16) else torch.float32
        if wm_bf16 is not None:   # mask cache: W (.) M was materialised for this (scores, threshold) state
            y = masked_linear_fwd(x2, wm_bf16, None, None, bias, io)
        else:
            y = masked_linear_fwd(x2, w_bf16, scores.detach(), thr_t, bias, io)
        ctx.save_for_backward(x2, scores, w_bf16, thr_t, w_f32)
        ctx.wm = wm_bf16
        ctx.x_shape = shp
        ctx.dx_dtype = torch.bfloat16 if x.dtype == torch.bfloat16 else torch.float32
        ctx.need_dx = x.requires_grad
        ctx.sink = sink
        return y.view(*shp[:-1], w_bf16.shape[0])

    # -- CRVQA_OPERAND=split: the same three GEMMs, each as hi.hi + lo.hi + hi.lo ------------------------------
    @staticmethod
    def _forward_split(ctx, x, scores, w_bf16, thr_t, bias, sink, wm_bf16, w_f32):
        shp = x.shape
        xh, xl = split_bf16(x.reshape(-1, shp[-1]).contiguous())
        wl = to_bf16(w_f32 - w_bf16.float())
        if wm_bf16 is not None:      # mask cache: plain (2-CTA) kernels on materialised masked halves
            wh, wl = wm_bf16, apply_mask_bf16(wl, scores.detach(), thr_t)
            sc = th = None
        else:                        # no cache: the in-kernel mask transform kernels derive the mask per call
            wh, sc, th = w_bf16, scores.detach(), thr_t
        y = masked_linear_fwd(xh, wh, sc, th, bias, torch.float32)
        y += masked_linear_fwd(xl, wh, sc, th, None, torch.float32)
        y += masked_linear_fwd(xh, wl, sc, th, None, torch.float32)
        ctx.save_for_backward(xh, xl, wh, wl, scores, w_f32, thr_t)
        ctx.masked_in_kernel = sc is not None
        ctx.x_shape, ctx.need_dx, ctx.sink = shp, x.requires_grad, sink
        return y.view(*shp[:-1], w_bf16.shape[0])

    @staticmethod
    def _backward_split(ctx, dy):
        xh, xl, wh, wl, scores, w_f32, thr_t = ctx.saved_tensors
        sc, th = (scores.detach(), thr_t) if ctx.masked_in_kernel else (None, None)
        dh, dl = split_bf16(dy.reshape(-1, dy.shape[-1]).contiguous().float())
        dx = None
        if ctx.need_dx:
            dx = masked_linear_bwd_dx(dh, wh, sc, th, torch.float32)
            dx += masked_linear_bwd_dx(dl, wh, sc, th, torch.float32)
            dx += masked_linear_bwd_dx(dh, wl, sc, th, torch.float32)
            dx = dx.view(ctx.x_shape)
        ds = None
        if ctx.needs_input_grad[1]:
            sink_grad = _sink_grad(ctx.sink)
            if sink_grad is not None:
                out, acc = sink_grad, ds_mode(ctx.sink)
            else:
                out, acc = torch.empty_like(w_f32), False
            masked_linear_bwd_ds(dh, xh, w_f32, out=out, accumulate=acc)
            masked_linear_bwd_ds(dl, xh, w_f32, out=out, accumulate=True)
            masked_linear_bwd_ds(dh, xl, w_f32, out=out, accumulate=True)
            if sink_grad is not None:
                _sink_done(ctx.sink)
            else:
                ds = out
        return dx, ds, None, None, None, None, None, None

    @staticmethod
    def backward(ctx, dy):
        if ctx.split:
            return MaskedLinearFn._backward_split(ctx, dy)
        x2, scores, w_bf16, thr_t, w_f32 = ctx.saved_tensors
        dy2 = to_bf16(dy.reshape(-1, dy.shape[-1]))
        sink_grad = _sink_grad(ctx.sink) if ctx.needs_input_grad[1] else None
        lane = ds_lane(dy.device) if (sink_grad is not None and ctx.need_dx) else None
        if lane is not None:
            lane.fork()                      # dY is ready here; dS need not wait for dX
        dx = None
        if ctx.need_dx:
            if ctx.wm is not None:
                dx = masked_linear_bwd_dx(dy2, ctx.wm, None, None, ctx.dx_dtype).view(ctx.x_shape)
            else:
                dx = masked_linear_bwd_dx(dy2, w_bf16, scores.detach(), thr_t, ctx.dx_dtype).view(ctx.x_shape)
        ds = None
        if ctx.needs_input_grad[1]:
            if sink_grad is not None:
                if lane is not None:
                    with torch.cuda.stream(lane.stream):
                        masked_linear_bwd_ds(dy2, x2, w_f32, out=sink_grad, accumulate=ds_mode(ctx.sink))
                    lane.hold(dy2, x2)
                else:
                    masked_linear_bwd_ds(dy2, x2, w_f32, out=sink_grad, accumulate=ds_mode(ctx.sink))
                _sink_done(ctx.sink)
            else:
                ds = masked_linear_bwd_ds(dy2, x2, w_f32)
        return dx, ds, None, None, None, None, None, None


class MaskedLinearSmallKFn(torch.autograd.Function):
    """Same contract for inner dimensions TMA cannot take (box_fc, K = 4): exact fp32 SIMT kernels."""

    @staticmethod
    def forward(ctx, x, scores, weight, thr, bias, sink=None):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).contiguous().float()
        thr_t = as_thr(thr, x.device)
        N, K = weight.shape
        y = torch.empty((x2.shape[0], N), dtype=torch.float32, device=x.device)
        check(lib.crv_masked_linear_small_k_fwd(_p(x2), _p(weight), _p(scores.detach()), _p(thr_t), _p(bias), _p(y),
                                                x2.shape[0], N, K, _stream()), "crv_masked_linear_small_k_fwd")
        ctx.save_for_backward(x2, scores, weight, thr_t)
        ctx.x_shape = shp
        ctx.need_dx = x.requires_grad
        ctx.sink = sink
        return y.view(*shp[:-1], N)

    @staticmethod
    def backward(ctx, dy):
        x2, scores, weight, thr_t = ctx.saved_tensors
        N, K = weight.shape
        dy2 = dy.reshape(-1, N).contiguous().float()
        dx = torch.empty_like(x2) if ctx.need_dx else None
        sink_grad = _sink_grad(ctx.sink)
        acc = 0
        if sink_grad is not None:
            ds, acc = sink_grad, int(ctx.sink._grad_dirty)
        else:
            ds = torch.empty((N, K), dtype=torch.float32, device=dy.device)
        check(lib.crv_masked_linear_small_k_bwd(_p(dy2), _p(x2), _p(weight), _p(scores.detach()), _p(thr_t), _p(dx),
                                                _p(ds), acc, x2.shape[0], N, K, _stream()),
              "crv_masked_linear_small_k_bwd")
        if sink_grad is not None:
            _sink_done(ctx.sink)
            ds = None
        return (dx.view(ctx.x_shape) if dx is not None else None), ds, None, None, None, None


class MaskedEmbeddingFn(torch.autograd.Function):
    """F.embedding(ids, weight * binarize(scores, thr), padding_idx) without masking the whole table."""

    @staticmethod
    def forward(ctx, ids, scores, weight, thr, padding_idx, sink=None):
        _need_cuda(ids, scores, weight)
        ids_c = ids.contiguous()
        thr_t = as_thr(thr, scores.device)
        vocab, dim = weight.shape
        out = torch.empty((*ids.shape, dim), dtype=torch.float32, device=weight.device)
        check(lib.crv_masked_embedding_fwd(_p(ids_c), _p(weight), _p(scores.detach()), _p(thr_t), _p(out),
                                           ids_c.numel(), vocab, dim, _stream()), "crv_masked_embedding_fwd")
        ctx.save_for_backward(ids_c, weight)
        ctx.padding_idx = -1 if padding_idx is None else int(padding_idx)
        ctx.sink = sink
        return out

    @staticmethod
    def backward(ctx, dout):
        ids_c, weight = ctx.saved_tensors
        vocab, dim = weight.shape
        sink_grad = _sink_grad(ctx.sink)
        if sink_grad is not None:
            ds = sink_grad
            if ds_mode(ctx.sink) == DS_OVERWRITE:
                ds.zero_()
        else:
            ds = torch.zeros((vocab, dim), dtype=torch.float32, device=weight.device)
        d = dout.contiguous().float()
        check(lib.crv_masked_embedding_bwd(_p(ids_c), _p(d), _p(weight), _p(ds), ids_c.numel(), vocab, dim,
                                           ctx.padding_idx, _stream()), "crv_masked_embedding_bwd")
        if sink_grad is not None:
            _sink_done(ctx.sink)
            ds = None
        return None, ds, None, None, None, None


# ----------------------------------------------------------------------------- losses
def _loss_ws(B, device):
    return torch.empty(3 * B, dtype=torch.float32, device=device)


class _FusedLoss(torch.autograd.Function):
    """Common shell: forward runs the fused fwd+bwd kernel and stashes the gradients."""

    @staticmethod
    def forward(ctx, kind, logits, *args):
        _need_cuda(logits)
        B, A = logits.shape
        lg = logits.detach().contiguous().float()
        out = torch.empty(2, dtype=torch.float32, device=lg.device)
        dl = torch.empty_like(lg)
        ws = _loss_ws(B, lg.device)
        dfac = None
        if kind == "bce":
            (labels,) = args
            check(lib.crv_vqa_loss_bce(_p(lg), _p(labels.contiguous()), _p(out), _p(dl), B, A, _p(ws), _stream()),
                  "crv_vqa_loss_bce")
        elif kind == "lpf":
            bias, max_label, gamma, labels = args
            check(lib.crv_vqa_loss_lpf(_p(lg), _p(bias.contiguous()), _p(max_label.contiguous()), float(gamma),
                                       _p(out), _p(labels.contiguous() if labels is not None else None), _p(dl),
                                       B, A, _p(ws), _stream()), "crv_vqa_loss_lpf")
        elif kind == "rubi":
            bias, max_label, labels = args
            check(lib.crv_vqa_loss_rubi(_p(lg), _p(bias.contiguous()), _p(max_label.contiguous()), _p(out),
                                        _p(labels.contiguous() if labels is not None else None), _p(dl), B, A, _p(ws),
                                        _stream()), "crv_vqa_loss_rubi")
        elif kind == "lmh":
            bias, labels, factor_pre, smooth, w = args
            fp = factor_pre.detach().contiguous().float().view(-1)
            dfac = torch.empty_like(fp)
            check(lib.crv_vqa_loss_lmh(_p(lg), _p(bias.contiguous()), _p(labels.contiguous()), _p(fp), float(smooth),
                                       float(w), _p(out), _p(dl), _p(dfac), B, A, _p(ws), _stream()),
                  "crv_vqa_loss_lmh")
            ctx.fshape = factor_pre.shape
        else:
            raise ValueError(kind)
        ctx.kind = kind
        ctx.nargs = len(args)
        ctx.save_for_backward(dl, dfac) if dfac is not None else ctx.save_for_backward(dl)
        ctx.mark_non_differentiable(out[1:])
        loss = out[0]
        score = out[1].detach()
        return loss, score

    @staticmethod
    def backward(ctx, gloss, gscore):
        saved = ctx.saved_tensors
        dl = saved[0] * gloss
        grads = [None, dl] + [None] * ctx.nargs
        if ctx.kind == "lmh":
            grads[2 + 2] = (saved[1] * gloss).view(ctx.fshape)  # factor_pre is the 3rd extra argument
        return tuple(grads)


def vqa_loss_bce(logits, labels):
    """(loss, batch_score): instance_bce_with_logits (modeling_lxmert.py:248-253) + VQA score."""
    return _FusedLoss.apply("bce", logits, labels)


def vqa_loss_lpf(logits, bias, max_label, gamma, labels=None):
    """LPF_loss (mask_trainer_VQA.py:111-129) + VQA score (needs labels)."""
    return _FusedLoss.apply("lpf", logits, bias, max_label, gamma, labels)


def vqa_loss_rubi(logits, bias, max_label, labels=None):
    """RUBI_loss (mask_trainer_VQA.py:131-135) + VQA score (needs labels)."""
    return _FusedLoss.apply("rubi", logits, bias, max_label, labels)


def vqa_loss_bias_product(logits, bias, labels, smooth):
    """BiasProduct (vqa_debias_loss_functions.py:83-122) = the LearnedMixin expression with factor 1 and no entropy
    term: the LMH kernel with a constant pre-softplus factor log(e - 1)."""
    fp = torch.full((logits.shape[0], 1), 0.5413248546129181, dtype=torch.float32, device=logits.device)
    return _FusedLoss.apply("lmh", logits, bias, labels, fp, smooth, 0.0)


def vqa_loss_lmh(logits, bias, labels, factor_pre, smooth, w):
    """LearnedMixin (vqa_debias_loss_functions.py:148-196); factor_pre = bias_lin(pooled), pre-softplus."""
    return _FusedLoss.apply("lmh", logits, bias, labels, factor_pre, smooth, w)


# ----------------------------------------------------------------------------- optimiser pieces
_sumsq_ws = {}


def _sumsq_workspace(device):
    """Per-device scratch of the deterministic sum of squares (zeroed once; the kernels leave it zeroed)."""
    key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
    ws = _sumsq_ws.get(key)
    if ws is None:
        ws = _sumsq_ws[key] = torch.zeros(lib.crv_sumsq_workspace_bytes() // 4, dtype=torch.float32, device=device)
    return ws


def sumsq_into(x, acc):
    check(lib.crv_sumsq(_p(x), x.numel(), _p(acc), _p(_sumsq_workspace(x.device)), _stream()), "crv_sumsq")


def sumsq_segmented_into(x, chunks, acc):
    if chunks.shape[0]:
        check(lib.crv_sumsq_segmented(_p(x), _p(chunks), chunks.shape[0], _p(acc), _p(_sumsq_workspace(x.device)),
                                      _stream()), "crv_sumsq_segmented")


def adam_step_size(lr, step, beta1, beta2, correct_bias=True):
    if not correct_bias:
        return lr
    return lr * math.sqrt(1.0 - beta2 ** step) / (1.0 - beta1 ** step)


ADAM_REFERENCE, ADAM_TORCH = 0, 1     # root optimization.AdamW (stage 2) | torch.optim.Adam (stage 3)


def adam_hyper(mode, lr, step, beta1, beta2, correct_bias=True):
    """(lr, step_size, inv_bc2_sqrt) of crv_adamw_step / crv_adamw_segmented for optimiser step `step` (1-based)."""
    if mode == ADAM_REFERENCE:
        return lr, adam_step_size(lr, step, beta1, beta2, correct_bias), 1.0
    return lr, lr / (1.0 - beta1 ** step), 1.0 / math.sqrt(1.0 - beta2 ** step)


def adamw_step_flat(p, g, m, v, s, lr, step, beta1, beta2, eps, weight_decay, total_sumsq=None, max_norm=1.0,
                    correct_bias=True, hyper=None, mode=ADAM_REFERENCE):
    """hyper: optional device tensor {lr, step_size, inv_bc2_sqrt}; when given it overrides lr / step (graph replay)."""
    lr, step_size, inv_bc2 = adam_hyper(mode, lr, step, beta1, beta2, correct_bias)
    check(lib.crv_adamw_step(_p(p), _p(g), _p(m), _p(v), _p(s), p.numel(), float(lr), float(step_size),
                             float(beta1), float(beta2), float(eps), float(weight_decay), _p(total_sumsq),
                             float(max_norm), _p(hyper), int(mode), float(inv_bc2), _stream()), "crv_adamw_step")


def adamw_segmented(p, g, m, v, s, chunks, thr_vec, w16, wm, lr, step, beta1, beta2, eps, weight_decay,
                    total_sumsq=None, max_norm=1.0, correct_bias=True, hyper=None, zero_grad=False,
                    mode=ADAM_REFERENCE):
    """crv_adamw_segmented: clip + Adam over an arena + masked-operand refresh (+ gradient clearing)."""
    lr, step_size, inv_bc2 = adam_hyper(mode, lr, step, beta1, beta2, correct_bias)
    check(lib.crv_adamw_segmented(_p(p), _p(g), _p(m), _p(v), _p(s), _p(chunks), chunks.shape[0], _p(thr_vec), _p(w16),
                                  _p(wm), float(lr), float(step_size), float(beta1), float(beta2), float(eps),
                                  float(weight_decay), _p(total_sumsq), float(max_norm), _p(hyper), int(bool(zero_grad)),
                                  int(mode), float(inv_bc2), _stream()), "crv_adamw_segmented")


_colsum_ws = {}


def colsum_bf16(x, out, accumulate=False):
    """out[n] (+)= sum_m x[m, n] for a bf16 [M, N] matrix (bias gradient of a linear layer), deterministic."""
    M, N = x.shape
    key = (x.device.index, N)
    ws = _colsum_ws.get(key)
    if ws is None:
        ws = _colsum_ws[key] = torch.empty(lib.crv_colsum_workspace_bytes(N) // 4, dtype=torch.float32, device=x.device)
    check(lib.crv_colsum_bf16(_p(x), M, N, _p(out), int(bool(accumulate)), _p(ws), _stream()), "crv_colsum_bf16")


def partial_reduce(part, out, accumulate=False):
    """out[j] (+)= sum_i part[i, j]."""
    check(lib.crv_partial_reduce(_p(part), part.shape[0], part.shape[1], _p(out), int(bool(accumulate)), _stream()),
          "crv_partial_reduce")


class MomentumPlan:
    """Pointer / chunk tables of a fixed list of (online, twin) fp32 tensor pairs for crv_momentum_update: built once
    (the pairs of mPLUG's towers never change), rebuilt when a tensor was re-allocated (its data_ptr moved)."""
    CHUNK = 16384       # elements per row

    def __init__(self, online, twins):
        dev = twins[0].device
        self.key = tuple(t.data_ptr() for t in online) + tuple(t.data_ptr() for t in twins)
        rows = []
        for i, (a, b) in enumerate(zip(online, twins)):
            if (a.dtype != torch.float32 or b.dtype != torch.float32 or a.numel() != b.numel() or not a.is_contiguous()
                    or not b.is_contiguous() or a.data_ptr() % 16 or b.data_ptr() % 16 or a.device != dev or b.device != dev):
                raise ValueError("momentum pairs must be contiguous, 16-byte aligned fp32 tensors of equal size on one device")
            n = a.numel()
            for c0 in range(0, n, self.CHUNK):
                ln = min(self.CHUNK, n - c0)
                rows.append((i, c0 // 4, ln // 4, ln % 4))
        self.online = torch.tensor([t.data_ptr() for t in online], dtype=torch.int64, device=dev)
        self.twins = torch.tensor([t.data_ptr() for t in twins], dtype=torch.int64, device=dev)
        self.rows = torch.tensor(rows, dtype=torch.int32, device=dev).contiguous()

    @staticmethod
    def key_of(online, twins):
        return tuple(t.data_ptr() for t in online) + tuple(t.data_ptr() for t in twins)

    def run(self, momentum):
        m = float(torch.tensor(momentum, dtype=torch.float32))
        om = float(torch.tensor(1.0 - momentum, dtype=torch.float32))      # torch rounds the Python scalar 1 - m to fp32
        check(lib.crv_momentum_update(_p(self.online), _p(self.twins), _p(self.rows), self.rows.shape[0], m, om,
                                      _stream()), "crv_momentum_update")


# ----------------------------------------------------------------------------- multi-tensor optimiser pieces (mPLUG engine)
MULTI_CHUNK = 16384      # elements per row of the multi-tensor kernels (a multiple of 8: rows start 16-byte aligned)


def multi_rows(numels, flags=None, chunk=MULTI_CHUNK):
    """Host row table {tensor, first element / 8, elements, flags} cutting tensor i (numels[i] elements) into chunks."""
    rows = []
    for i, n in enumerate(numels):
        f = int(flags[i]) if flags is not None else 0
        for c0 in range(0, int(n), chunk):
            rows.append((i, c0 // 8, min(chunk, int(n) - c0), f))
    return rows


def upload(values, dtype, device, out=None):
    """Host list -> device tensor without a host-side wait: staged in pinned memory the caching host allocator keeps
    alive until the copy has run (a reused pinned buffer could be overwritten while its copy is still queued)."""
    host = torch.tensor(values, dtype=dtype).pin_memory()
    if out is None:
        out = torch.empty(host.shape, dtype=dtype, device=device)
    out.copy_(host, non_blocking=True)
    return out


def sumsq_multi(ptrs, rows, acc):
    """acc += sum of squares over the rows of the tensors whose addresses `ptrs` (device int64) lists."""
    if rows.shape[0]:
        check(lib.crv_sumsq_multi(_p(ptrs), _p(rows), rows.shape[0], _p(acc), _p(_sumsq_workspace(acc.device)),
                                  _stream()), "crv_sumsq_multi")


def adamw_multi(p, g, m, v, w16, wm, thr, rows, lr, step, beta1, beta2, eps, weight_decay, total_sumsq=None,
                max_norm=1.0):
    """crv_adamw_multi: clip + torch.optim.AdamW step `step` (1-based) over the rows of many tensors (device int64
    address tables), refreshing W (.) (S_new > thr) for the tensors whose rows carry flag 1."""
    if rows.shape[0]:
        check(lib.crv_adamw_multi(_p(p), _p(g), _p(m), _p(v), _p(w16), _p(wm), _p(thr), _p(rows), rows.shape[0],
                                  float(lr), int(step), float(beta1), float(beta2), float(eps), float(weight_decay),
                                  _p(total_sumsq), float(max_norm), _stream()), "crv_adamw_multi")
