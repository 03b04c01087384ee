"""B200 engine pieces under the stage-2 trainers: HBM layout of the trainable state and the
data-parallel gradient exchange.  None of this exists in the reference (which wraps the model in
nn.DataParallel / DDP, hg_transformers/mask_trainer_VQA.py:533-543); it is what makes the reference's
loop cheap on one 8 x B200 box.

ScoreArena
    All score tensors (`weight_mask`, 207 M fp32 for LXMERT) live in ONE flat buffer, their gradients
    in a second one, Adam's exp_avg / exp_avg_sq / sum in three more (5 x 829 MB of 180 GB HBM).
    Module parameters are views, so the reference's per-tensor API keeps working, while
      * the dS GEMM epilogue accumulates straight into the gradient buffer (no autograd copy),
      * clip + AdamW is one streaming launch over the arena,
      * the threshold refresh reads one pointer table,
      * the gradient all-reduce runs on large contiguous slices with no flatten / unflatten copies.

GradSync
    Data parallel = all-reduce(mean) of the score gradients and the classifier gradients only (the
    frozen weights have none).  The arena is cut into buckets in forward order; as soon as the
    backward pass has produced every module of a bucket, its slice is all-reduced asynchronously (NCCL
    over NVLink / NVSwitch) while the remaining dS / dX GEMMs run.
"""
import contextlib
import os

import torch
import torch.distributed as dist

from crvqa import ops

_ALIGN = 64  # floats: 256-byte alignment of every module slice (TMA needs 16 B; keep sector alignment)


def masked_modules_of(model):
    """(name, module) for every masked module, found the way the reference trainers find them: by the
    presence of a `threshold` attribute (mask_trainer_VQA.py:473,935)."""
    out = []
    for name, module in model.named_modules():
        if hasattr(module, "threshold") and hasattr(module, "weight_mask"):
            out.append((name[7:] if name.startswith("module.") else name, module))
    return out


def execution_order(named_modules):
    """Arena (and therefore bucket) order = the order the fused LXMERT forward executes the modules: the visual
    feature encoder and the word embeddings, then language layer i and vision layer i in lockstep, the remaining
    language layers, the cross layers, the pooler (modeling_lxmert._forward_fast).  The backward pass completes
    gradients in the reverse order.  Other models (VisualBERT: one stack) keep named_modules order."""
    import re

    def key(item):
        name = item[0]
        if "visn_fc" in name or "embeddings" in name:
            return (0, 0, 0)
        m = re.search(r"\.(layer|r_layers)\.(\d+)\.", name)
        if m and ".x_layers." not in name:
            return (1, int(m.group(2)), 0 if m.group(1) == "layer" else 1)
        if ".x_layers." in name:
            return (2, 0, 0)
        if "pooler" in name:
            return (3, 0, 0)
        return (1, 0, 0)
    if not any(".r_layers." in n for n, _ in named_modules):
        return list(named_modules)
    return sorted(named_modules, key=key)          # stable: named_modules order inside each group


class ScoreArena:
    def __init__(self, named_modules, device=None):
        self.names = [n for n, _ in named_modules]
        self.modules = [m for _, m in named_modules]
        if not self.modules:
            raise ValueError("no masked modules to place in the arena")
        device = device or self.modules[0].weight_mask.device
        self.offsets, off = [], 0
        for m in self.modules:
            self.offsets.append(off)
            off += (m.weight_mask.numel() + _ALIGN - 1) // _ALIGN * _ALIGN
        self.total = off
        self.scores = torch.zeros(off, dtype=torch.float32, device=device)
        self.grads = torch.zeros(off, dtype=torch.float32, device=device)
        self.exp_avg = self.exp_avg_sq = self.sum = None
        self._index = {}
        for i, m in enumerate(self.modules):
            p = m.weight_mask
            view = self._view(self.scores, i)
            view.copy_(p.data)
            p.data = view
            p.grad = self._view(self.grads, i)
            m._arena_grad = p.grad
            m._grad_dirty = False
            m._grad_zero = True        # the gradient buffer starts as zeros
            m._arena = self
            self._index[id(p)] = i
        # thresholds of all modules in one device vector; module.threshold stays a 0-dim tensor (a view)
        self.thr_vec = torch.tensor([float(m.threshold) for m in self.modules], dtype=torch.float32, device=device)
        for i, m in enumerate(self.modules):
            m.threshold = self.thr_vec[i]
        # mask cache (enable_mask_cache): bf16 weights and W (.) M, same element offsets as the scores
        self.cache_on = False
        # CRVQA_KEEP_GRADS=1 (parity tests that read the gradients after a full step): the optimiser pass does not
        # clear the gradient arena; split score-gradient GEMMs then clear their output with a memset per module
        self.keep_grads = os.environ.get("CRVQA_KEEP_GRADS", "0") == "1"
        self.step_chunks = None
        self.shard = None
        self.epoch = 0
        self.w16 = self.wm = self.w32 = self.chunks = None
        if self.scores.is_cuda:
            self._step_chunk_table()

    def _view(self, flat, i):
        p = self.modules[i].weight_mask
        return flat[self.offsets[i]: self.offsets[i] + p.numel()].view(p.shape)

    def owns(self, p):
        return id(p) in self._index

    # -- what crvqa.fused.ProjectionGroup asks of an arena (hg_transformers._engine_ft.WeightArena answers the same) ----
    def index_of(self, m):
        return self._index[id(m.weight_mask)]

    def weight_shape(self, m):
        return tuple(m.weight.shape)

    def anchor(self, m):
        return m.weight_mask

    def module_ready(self, m):
        return m.weight_mask.is_cuda

    def group_bias(self, modules):
        """(bias, bias gradient): the biases are frozen in stage 2 -- one private fp32 copy, no gradient."""
        biases = [m.bias for m in modules]
        if any(b is None for b in biases):
            return None, None
        return torch.cat([b.detach().float() for b in biases]).contiguous(), None

    def loose_grads(self):
        return []

    def _ensure_state(self):
        if self.exp_avg is None:
            self.exp_avg = torch.zeros_like(self.scores)
            self.exp_avg_sq = torch.zeros_like(self.scores)
            self.sum = torch.zeros_like(self.scores)

    def state_views(self, p):
        i = self._index.get(id(p))
        if i is None:
            return None
        self._ensure_state()
        return {"sum": self._view(self.sum, i), "exp_avg": self._view(self.exp_avg, i),
                "exp_avg_sq": self._view(self.exp_avg_sq, i)}

    # -- mask cache ------------------------------------------------------------------------------
    def set_thresholds(self, thr):
        """New thresholds (device vector, arena order) from reset_threshold; invalidates the mask cache."""
        self.thr_vec.copy_(thr)
        for i, m in enumerate(self.modules):
            m.threshold = self.thr_vec[i]
        self.epoch += 1

    def _gemm_module(self, m):
        return "embedding" not in m.name and m.weight.dim() == 2 and m.weight.shape[1] % 8 == 0

    def enable_mask_cache(self):
        """Scores change only at the optimiser step and thresholds only at reset_threshold, so the masked
        bf16 weight of every module is materialised once per step (ONE launch over the arena) and the
        forward / dX GEMMs read it as a plain operand instead of re-deriving the mask per call."""
        if self.cache_on:
            return
        dev = self.scores.device
        self.w16 = torch.zeros(self.total, dtype=torch.bfloat16, device=dev)
        self.wm = torch.zeros(self.total, dtype=torch.bfloat16, device=dev)
        # the frozen fp32 weights move into a flat buffer with the scores' element offsets too (the Parameters become
        # views, no second copy): the dS epilogue multiplies by fp32 W, and a projection group (query|key|value) needs
        # its three weights as one contiguous [sum N, K] multiplier
        self.w32 = torch.zeros(self.total, dtype=torch.float32, device=dev)
        rows = []
        for i, m in enumerate(self.modules):
            if not self._gemm_module(m):
                continue
            n, off = m.weight_mask.numel(), self.offsets[i]
            if m.weight.dtype == torch.float32:
                w32 = self.w32[off: off + n].view(m.weight.shape)
                w32.copy_(m.weight.detach())
                m.weight.data = w32
            self.w16[off: off + n].view(m.weight.shape).copy_(ops.to_bf16(m.weight.detach()))
            m._w16 = self.w16[off: off + n].view(m.weight.shape)
            m._w16_key = (m.weight.data_ptr(), m.weight._version, m.weight.device)
            m._wm = self.wm[off: off + n].view(m.weight.shape)
            for c0 in range(0, n, 8192):
                rows.append(((off + c0) // 8, min(8192, n - c0), i, 0))
        self.chunks = torch.tensor(rows, dtype=torch.int32, device=dev).contiguous()
        self.cache_on = True
        self.step_chunks = None
        self._step_chunk_table()    # built now: a CUDA-graph capture of the step must not create it
        self.refresh_masked()

    def _chunk_rows(self, ranges=None, modules=None, extra_flags=0):
        """{start / 8, length, segment, flags} rows (flags bit 0: the segment has a bf16 operand) over whole modules
        (`modules`: indices, default all) or over the parts of the modules inside `ranges` (element intervals whose
        bounds are multiples of 8)."""
        rows = []
        for i, m in enumerate(self.modules):
            if modules is not None and i not in modules:
                continue
            n, off = m.weight_mask.numel(), self.offsets[i]
            flag = 1 if (self.cache_on and self._gemm_module(m)) else 0
            spans = [(off, off + n)] if ranges is None else [(max(off, lo), min(off + n, hi)) for lo, hi in ranges]
            for lo, hi in spans:
                for c0 in range(lo, hi, 8192):
                    rows.append((c0 // 8, min(8192, hi - c0), i, flag | extra_flags))
        if not rows:
            return torch.zeros((0, 4), dtype=torch.int32, device=self.scores.device)
        return torch.tensor(rows, dtype=torch.int32, device=self.scores.device).contiguous()

    def _step_chunk_table(self):
        """Rows the optimiser pass of THIS rank covers: every module, or (sharded optimiser, GradSync) the element
        ranges this rank owns plus the replicated modules."""
        if self.step_chunks is None:
            if self.shard is None:
                self.step_chunks = self._chunk_rows()
            else:
                own = self._chunk_rows(ranges=self.shard["own"], modules=self.shard["sharded_modules"])
                rep = self._chunk_rows(modules=self.shard["replicated_modules"])
                self.shard["own_chunks"], self.shard["rep_chunks"] = own, rep
                # slices other ranks own: clear-only rows (their gradient holds reduce-scatter leftovers)
                other = self._chunk_rows(ranges=self.shard["other"], modules=self.shard["sharded_modules"], extra_flags=2)
                self.step_chunks = torch.cat([own, rep, other]).contiguous()
        return self.step_chunks

    def install_shard(self, own_ranges, other_ranges, sharded_modules, replicated_modules, sync):
        """GradSync (sharded mode): this rank updates only `own_ranges` of the sharded modules; their masked operands
        reach the other ranks by all-gather before first use (wait_ready), their scores on demand (sync.sync_scores)."""
        self.shard = {"own": own_ranges, "other": other_ranges, "sharded_modules": set(sharded_modules),
                      "replicated_modules": set(replicated_modules), "sync": sync}
        self.step_chunks = None
        self._step_chunk_table()

    def wait_ready(self, m):
        """Make the current stream wait until module `m`'s masked operand is complete on this rank."""
        if self.shard is not None:
            self.shard["sync"].wait_operand(m)

    def refresh_masked(self):
        if not self.cache_on:
            return
        if self.shard is not None:
            self.shard["sync"].sync_scores()          # the whole arena is re-masked here: every score must be current
        ops.apply_mask_segmented(self.w16, self.scores, self.thr_vec, self.chunks, self.wm)
        self._mark_cache_valid()
        if self.shard is not None:
            self.shard["sync"].operands_are_current()

    def _mark_cache_valid(self):
        for m in self.modules:
            m._wm_epoch = self.epoch
            m._wm_sver = m.weight_mask._version
            m._wm_thr_ptr = m.threshold.data_ptr()

    def cached_masked_weight(self, m):
        """The module's W (.) M if it is valid for the current scores and threshold, else None."""
        if (self.cache_on and getattr(m, "_wm_epoch", -1) == self.epoch and m.weight_mask._version == m._wm_sver
                and torch.is_tensor(m.threshold) and m.threshold.data_ptr() == m._wm_thr_ptr):
            return m._wm          # callers that READ it on a stream call wait_ready(m) first (sharded optimiser)
        return None

    # -- per-step protocol ---------------------------------------------------------------------
    def begin_step(self):
        """Replaces zero_grad for the arena: nothing is memset; the first dS of a module overwrites."""
        for m in self.modules:
            m._grad_dirty = False
            if m.weight_mask.grad is None:  # someone called zero_grad(set_to_none=True)
                m.weight_mask.grad = m._arena_grad

    def finalize_grads(self):
        """Modules that received no gradient this step (e.g. the vision side of the last cross layer,
        which nothing downstream reads) must contribute zeros.  Also the point where the dS lane (ops._DsLane)
        rejoins the current stream: every reader of the gradients comes after this call."""
        ops.ds_lane_join()
        for m in self.modules:
            if not m._grad_dirty:
                if not getattr(m, "_grad_zero", False):
                    m._arena_grad.zero_()
                m._grad_dirty = True

    def grad_sumsq_into(self, acc):
        ops.sumsq_into(self.grads, acc)

    def adamw_step(self, lr, step, beta1, beta2, eps, weight_decay, correct_bias, clip_sumsq, max_norm,
                   with_sum=True, hyper=None):
        """Clip + AdamW over the arena, the masked-operand refresh and (unless keep_grads) the gradient clearing of
        the reference's model.zero_grad() -- ONE launch (crv_adamw_segmented)."""
        self._ensure_state()
        self.epoch += 1            # scores move: the mask cache follows in the same pass
        zero = not self.keep_grads
        ops.adamw_segmented(self.scores, self.grads, self.exp_avg, self.exp_avg_sq, self.sum if with_sum else None,
                            self._step_chunk_table(), self.thr_vec, self.w16 if self.cache_on else None,
                            self.wm if self.cache_on else None, lr, step, beta1, beta2, eps, weight_decay, clip_sumsq,
                            max_norm, correct_bias, hyper, zero)
        if self.cache_on:
            self._mark_cache_valid()
        if zero:
            for m in self.modules:
                m._grad_zero = True
        if self.shard is not None:
            self.shard["sync"].after_optimizer_step()

    def release(self):
        """Give the parameters their own storage back (used when a trainer is torn down)."""
        for m in self.modules:
            p = m.weight_mask
            p.data = p.data.clone()
            p.grad = None
            m._arena_grad = None
            m._arena = None
            if self.w32 is not None and self._gemm_module(m):
                m.weight.data = m.weight.data.clone()


class GradSync:
    """Data-parallel exchange of an arena's gradients (+ a few loose tensors), bucketed and asynchronous.

    Buckets are runs of consecutive modules in arena (= execution) order; as soon as the backward pass has produced
    every module of a bucket, its collective is issued on a side stream while the remaining GEMMs run.

    Two modes (CRVQA_DP=sharded | allreduce; sharded needs NCCL and 2, 4 or 8 ranks):

    allreduce  every bucket is all-reduced (mean) and every rank runs the whole optimiser pass (the reference's DDP
               semantics, hg_transformers/mask_trainer_VQA.py:537-543, literally).
    sharded    same result, less traffic and less HBM work: a bucket of GEMM modules is REDUCE-SCATTERED, rank r keeps
               the mean gradient of slice r only, runs clip + AdamW + masked-operand refresh on that slice (1/N of the
               44 B per score), and the bf16 masked operands W (.) M -- all the forward / dX GEMMs read -- are
               ALL-GATHERED at the start of the next step, bucket by bucket in execution order, overlapping the forward
               pass (a module's first use waits for its bucket only).  Per step and score that is 4 B (fp32 reduce-
               scatter) + 2 B (bf16 all-gather) on the wire instead of 4 B + 4 B.  Modules whose scores are read
               directly by a kernel (word embeddings, box_fc) stay replicated: all-reduce + full update.  The fp32
               scores of non-owned slices go stale between threshold refreshes; sync_scores() all-gathers them
               (reset_threshold, save_model_mask, evaluation, end of training call it)."""

    def __init__(self, arena, bucket_bytes=32 << 20, group=None, mode=None):
        self.arena = arena
        self.group = group
        active = dist.is_available() and dist.is_initialized()
        self.world = dist.get_world_size(group) if active else 1
        self.rank = dist.get_rank(group) if active else 0
        self.enabled = self.world > 1
        mode = mode or os.environ.get("CRVQA_DP", "sharded")
        self.sharded = (self.enabled and mode == "sharded" and self.world in (2, 4, 8) and arena.scores.is_cuda
                        and arena.cache_on and dist.get_backend(group) == "nccl")
        # buckets = runs of consecutive modules (arena order == execution order).  The FIRST buckets of the arena are
        # the last ones the backward pass completes, and only their exchange is exposed at the end of the step, so
        # they start small (bucket_bytes / 4, / 2, then bucket_bytes).  Sharded mode: a bucket holds either GEMM
        # modules only (reduce-scatter) or replicated modules only (all-reduce).
        n = len(arena.modules)
        shardable = [self.sharded and arena._gemm_module(m) for m in arena.modules]
        self.bucket_of, self.bucket_ranges, self.bucket_members, self.bucket_sharded = [], [], [], []
        start_mod, start_off = 0, 0
        for i in range(n):
            end = arena.offsets[i + 1] if i + 1 < n else arena.total
            self.bucket_of.append(len(self.bucket_ranges))
            limit = (bucket_bytes >> max(0, 2 - len(self.bucket_ranges))) // 4
            last = i == n - 1
            if end - start_off >= limit or last or shardable[i + 1] != shardable[i]:
                self.bucket_ranges.append((start_off, end))
                self.bucket_members.append(list(range(start_mod, i + 1)))
                self.bucket_sharded.append(shardable[i])
                start_mod, start_off = i + 1, end
        self._pending = None
        self._handles = []
        self._sent = None
        self.defer = False  # gradient accumulation: exchange only after the last micro-batch
        for i, m in enumerate(arena.modules):
            m._sync = self
            m._sync_index = i
            m._calls_outstanding = 0
        self._wm_stale = False        # non-owned slices of the masked operands are out of date (sharded mode)
        self._scores_stale = False
        self._gather = None           # per-bucket handles of the running operand all-gather
        # Replicated buckets (the 94 MB word-embedding gradient) and the loose tensors travel on a SECOND communicator:
        # the embedding gradient is complete only when the backward pass is almost over, and on one communicator the
        # small reduce-scatters of the first layers would queue behind its all-reduce at the very end of the step.
        self.group_rep = dist.new_group(backend="nccl") if self.sharded else group
        self._loose, self._loose_left, self._loose_work = [], 0, None
        if self.sharded:
            own, other = [], []
            for (lo, hi), sh in zip(self.bucket_ranges, self.bucket_sharded):
                if sh:
                    assert (hi - lo) % (8 * self.world) == 0, "bucket does not split into 8-element aligned shards"
                    s = (hi - lo) // self.world
                    olo, ohi = lo + self.rank * s, lo + (self.rank + 1) * s
                    own.append((olo, ohi))
                    other += [r for r in ((lo, olo), (ohi, hi)) if r[1] > r[0]]
            arena.install_shard(own, other, [i for i in range(n) if shardable[i]],
                                [i for i in range(n) if not shardable[i]], self)

    # called by MaskedLinear1.forward in training mode: one more backward invocation is owed
    @staticmethod
    def note_forward(module):
        module._calls_outstanding = getattr(module, "_calls_outstanding", 0) + 1

    def _own(self, b):
        lo, hi = self.bucket_ranges[b]
        s = (hi - lo) // self.world
        return lo + self.rank * s, lo + (self.rank + 1) * s

    def attach_loose(self, params):
        """Loose trainable tensors (the answer head): their gradients are complete at the START of the backward pass,
        so their all-reduce is issued from a post-accumulate hook and overlaps everything that follows."""
        if not (self.sharded and hasattr(torch.Tensor, "register_post_accumulate_grad_hook")):
            return
        self._loose = [p for p in params]

        def hook(_p):
            if self._pending is None or self.defer:
                return
            self._loose_left -= 1
            if self._loose_left == 0:
                self._launch_loose()
        for p in self._loose:
            p.register_post_accumulate_grad_hook(hook)

    def _launch_loose(self):
        grads = [p.grad for p in self._loose if p.grad is not None]
        if not grads:
            return
        dev = grads[0].device
        side = self._side_stream(dev)
        ctx = torch.cuda.stream(side) if side is not None else contextlib.nullcontext()
        with ctx:
            flat = torch.cat([g.reshape(-1) for g in grads])
            work = dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=self.group_rep, async_op=True)
        self._loose_work = (work, flat, grads)

    def begin_step(self):
        self._pending = [len(b) for b in self.bucket_members]
        self._sent = [False] * len(self.bucket_ranges)
        self._handles = []
        self._loose_left, self._loose_work = len(self._loose), None
        for m in self.arena.modules:
            m._calls_outstanding = 0
        # a captured step graph must always contain the gather (replays follow optimiser passes), whatever the
        # Python-side flag says at capture time
        if self.sharded and self._gather is None and (self._wm_stale or torch.cuda.is_current_stream_capturing()):
            self._start_operand_gather()

    # -- sharded mode: masked operands ---------------------------------------------------------------------
    def after_optimizer_step(self):
        self._wm_stale = True
        self._scores_stale = True

    def operands_are_current(self):
        """The whole masked-operand arena was just recomputed locally from current scores (refresh_masked)."""
        self._wm_stale = False
        self._gather = None

    def _side_stream(self, device):
        lane = ops.ds_lane(device)
        if lane is None:
            return None
        lane.stream.wait_stream(torch.cuda.current_stream(device))
        lane.open = True
        return lane.stream

    def _start_operand_gather(self):
        """All-gather W (.) M of every sharded bucket, in execution order, on the side stream: the forward pass that
        follows waits bucket by bucket (wait_operand)."""
        wm = self.arena.wm
        side = self._side_stream(wm.device)
        ctx = torch.cuda.stream(side) if side is not None else contextlib.nullcontext()
        self._gather = {}
        with ctx:
            for b, sh in enumerate(self.bucket_sharded):
                if not sh:
                    continue
                lo, hi = self.bucket_ranges[b]
                olo, ohi = self._own(b)
                self._gather[b] = dist.all_gather_into_tensor(wm[lo:hi], wm[olo:ohi], group=self.group, async_op=True)

    def wait_operand(self, module):
        if self._gather is None:
            if self._wm_stale:         # a forward outside the step protocol (evaluation): gather now
                self._start_operand_gather()
            else:
                return
        b = self.bucket_of[module._sync_index]
        pending = [k for k in self._gather if k <= b]
        for k in pending:              # handles complete in issue order; waiting marks the dependency on this stream
            self._gather.pop(k).wait()
        if not self._gather:
            self._gather = None
            self._wm_stale = False

    def finish_operand_gather(self):
        if self._gather is not None:
            for k in sorted(self._gather):
                self._gather.pop(k).wait()
            self._gather = None
            self._wm_stale = False

    def sync_scores(self):
        """All-gather the fp32 scores of the sharded buckets so that every rank holds every current score."""
        if not (self.sharded and self._scores_stale):
            return
        sc = self.arena.scores
        for b, sh in enumerate(self.bucket_sharded):
            if sh:
                lo, hi = self.bucket_ranges[b]
                olo, ohi = self._own(b)
                dist.all_gather_into_tensor(sc[lo:hi], sc[olo:ohi], group=self.group)
        self._scores_stale = False

    def make_consistent(self):
        """Everything another consumer (evaluation, checkpoint, mask export) may read is current on this rank."""
        if self.sharded:
            if self._wm_stale and self._gather is None:
                self._start_operand_gather()
            self.finish_operand_gather()
            self.sync_scores()

    def global_sumsq(self, loose_grads):
        """Sum of squares of the full mean gradient (clip_grad_norm_): own slices summed over ranks + replicated
        modules + loose tensors (identical on every rank)."""
        a = self.arena
        acc = torch.zeros((), dtype=torch.float32, device=a.grads.device)
        a._step_chunk_table()
        ops.sumsq_segmented_into(a.grads, a.shard["own_chunks"], acc)
        dist.all_reduce(acc, op=dist.ReduceOp.SUM, group=self.group)
        ops.sumsq_segmented_into(a.grads, a.shard["rep_chunks"], acc)
        for g in loose_grads:
            if g is not None:
                ops.sumsq_into(g.contiguous(), acc)
        return acc

    # -- gradient exchange -------------------------------------------------------------------------------------
    def _launch(self, b):
        if self._sent[b]:
            return
        self._sent[b] = True
        if not self.enabled:
            return
        lo, hi = self.bucket_ranges[b]
        view = self.arena.grads[lo:hi]
        # The bucket's dS GEMMs may run on the dS lane.  Issue the collective FROM the lane (ordered after the current
        # stream as well): NCCL then waits for exactly the work that produced the bucket, and the dX chain on the
        # current stream is not held up at bucket boundaries.
        lane = ops.ds_lane(view.device) if view.is_cuda else None
        if lane is not None:
            lane.stream.wait_stream(torch.cuda.current_stream(view.device))
            ctx = torch.cuda.stream(lane.stream)
        else:
            ctx = contextlib.nullcontext()
        with ctx:
            if self.sharded and self.bucket_sharded[b]:
                olo, ohi = self._own(b)
                self._handles.append(dist.reduce_scatter_tensor(self.arena.grads[olo:ohi], view, op=dist.ReduceOp.AVG,
                                                                group=self.group, async_op=True))
            elif dist.get_backend(self.group) == "nccl":
                self._handles.append(dist.all_reduce(view, op=dist.ReduceOp.AVG, group=self.group_rep, async_op=True))
            else:
                h = dist.all_reduce(view, op=dist.ReduceOp.SUM, group=self.group, async_op=True)
                self._handles.append((h, view))
        if lane is not None:
            lane.open = True           # the lane holds work the current stream has to join before reading gradients

    def module_backward_done(self, module):
        """One backward invocation of `module` finished writing its dS."""
        if self._pending is None:
            return
        module._calls_outstanding -= 1
        if module._calls_outstanding > 0:
            return
        b = self.bucket_of[module._sync_index]
        self._pending[b] -= 1
        if self._pending[b] == 0 and not self.defer:
            self._launch(b)

    def finish(self, loose=()):
        """After backward: zero-fill untouched modules, send what is left, all-reduce the loose
        (classifier) gradients as one flat message, and wait for everything."""
        self.finish_operand_gather()     # a forward that skipped modules leaves their buckets' gathers un-awaited
        self.arena.finalize_grads()
        if self._sent is None:
            self.begin_step()
        for b in range(len(self.bucket_ranges)):
            self._launch(b)
        if self.enabled:
            loose = [g for g in loose if g is not None]
            if self._loose_work is not None:      # issued from the gradient hooks at the start of the backward pass
                work, flat, grads = self._loose_work
                work.wait()
                off = 0
                for g in grads:
                    g.copy_(flat[off: off + g.numel()].view_as(g))
                    off += g.numel()
                self._loose_work = None
            elif loose:
                flat = torch.cat([g.reshape(-1) for g in loose])
                dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.group)
                flat.div_(self.world)
                off = 0
                for g in loose:
                    g.copy_(flat[off: off + g.numel()].view_as(g))
                    off += g.numel()
            for h in self._handles:
                if isinstance(h, tuple):
                    h[0].wait()
                    h[1].div_(self.world)
                else:
                    h.wait()
            # (the reduce-scatter leaves partial sums in the slices this rank does not own; the optimiser pass clears
            # them through its clear-only chunk rows, ScoreArena._step_chunk_table)
        self._handles = []
        self._pending = None


class GraphedStep:
    """One whole optimisation step -- forward (188 masked-module calls), loss, backward, gradient exchange,
    clip + AdamW, mask-cache refresh -- captured as ONE CUDA graph and replayed per batch.

    The reference's loop is ~3000 small kernels issued from Python per step; once the kernels are fast
    the host cannot keep up (measured: 52 ms of Python per step against 43 ms of GPU work at batch 256).
    Replay needs (a) static input buffers, (b) the learning rate and Adam's bias-corrected step size in
    device memory (`optimizer.use_device_hyper`), because kernel arguments are frozen at capture, and
    (c) the host-side bookkeeping (scheduler, step counters) done outside the graph.
    """

    TENSOR_SLOTS = (0, 1, 2, 3, 6, 7)  # ids, feats, pos, target, bias, max_label of the 8-tuple batch
    HYPER_SLOTS = 8

    def __init__(self, trainer, model, optimizer, scheduler, warmup_steps=3):
        self.trainer, self.model, self.optimizer, self.scheduler = trainer, model, optimizer, scheduler
        self.warmup_steps = warmup_steps
        self.seen = 0
        self.graph = None
        self.static_inputs = None
        self.static_out = None
        dev = trainer.args.device
        self.hyper = torch.zeros(3, dtype=torch.float32, device=dev)    # {lr, step_size, 1 / sqrt(1 - b2^t)}
        # {lr, step_size} of step n travel through a RING of pinned buffers: the host runs many steps ahead of the GPU
        # (no sync between logging steps), so one reused buffer would be overwritten with step n+k's values while the
        # copy for step n is still queued.  A slot is rewritten only after the event behind its last copy completed.
        self.hyper_ring = [torch.zeros(3, dtype=torch.float32).pin_memory() for _ in range(self.HYPER_SLOTS)]
        self.hyper_events = [None] * self.HYPER_SLOTS
        self.hyper_turn = 0
        self.shapes = None
        # warm-up steps and the capture share one side stream, so the AccumulateGrad nodes of the loose
        # (classifier) parameters live on the capture stream, as torch.cuda.graph requires
        self.stream = torch.cuda.Stream(device=dev)

    def _next_step_index(self):
        p0 = self.optimizer.param_groups[0]["params"][0]
        return self.optimizer.state[p0]["step"] + 1

    def _upload_hyper(self):
        vals = self.optimizer.hyper_values(self._next_step_index())
        i = self.hyper_turn
        self.hyper_turn = (i + 1) % self.HYPER_SLOTS
        if self.hyper_events[i] is not None:
            self.hyper_events[i].synchronize()      # returns at once unless the GPU is HYPER_SLOTS steps behind
        buf = self.hyper_ring[i]
        buf[0], buf[1], buf[2] = vals[0], vals[1], (vals[2] if len(vals) > 2 else 1.0)
        self.hyper.copy_(buf, non_blocking=True)
        ev = self.hyper_events[i] or torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.hyper_events[i] = ev

    def _shape_key(self, inputs):
        return tuple((tuple(inputs[i].shape), inputs[i].dtype) for i in self.TENSOR_SLOTS)

    def step(self, inputs):
        """Returns (loss, score) as 0-dim device tensors, like Trainer._training_step."""
        t = self.trainer
        if self.graph is not None and self._shape_key(inputs) != self.shapes:
            return self.eager_step(inputs)  # odd-sized last batch of an epoch
        if self.graph is None and self.seen < self.warmup_steps:
            self.seen += 1
            cur = torch.cuda.current_stream()
            self.stream.wait_stream(cur)
            with torch.cuda.stream(self.stream):
                out = self._eager(inputs)
            cur.wait_stream(self.stream)
            return out
        dev = t.args.device
        if self.graph is None:
            # nothing may be CREATED inside the capture that has to survive it: a lazily built Adam state or dropout
            # counter would be allocated from the graph's pool and re-initialised by every replay
            if hasattr(self.optimizer, "ensure_state"):
                self.optimizer.ensure_state()
            from crvqa.fused import RngState
            RngState.get(dev)
            ops._sumsq_workspace(torch.device(dev))
            self.shapes = self._shape_key(inputs)
            self.static_inputs = list(inputs)
            for i in self.TENSOR_SLOTS:
                self.static_inputs[i] = inputs[i].to(dev, copy=True)
            self.optimizer.use_device_hyper(self.hyper)
            self._upload_hyper()
            torch.cuda.synchronize()
            self.graph = torch.cuda.CUDAGraph()
            steps_before = self._next_step_index()
            with torch.cuda.graph(self.graph, stream=self.stream):
                loss, score = t._device_step(self.model, self.static_inputs, self.optimizer)
                self.static_out = (loss, score)
            # capture ran the Python side of optimizer.step() once without executing any kernel
            self.optimizer.advance_steps(steps_before - self._next_step_index())
        else:
            for i in self.TENSOR_SLOTS:
                self.static_inputs[i].copy_(inputs[i], non_blocking=True)
        self._upload_hyper()
        self.graph.replay()
        self.optimizer.advance_steps(1)
        self.scheduler.step()
        return self.static_out

    def eager_step(self, inputs):
        """One eager step on the capture stream (profiling, odd-sized batches): the AccumulateGrad nodes of the loose
        parameters were created on that stream, and autograd warns when their producer runs on another one."""
        cur = torch.cuda.current_stream()
        self.stream.wait_stream(cur)
        with torch.cuda.stream(self.stream):
            out = self._eager(inputs)
        cur.wait_stream(self.stream)
        return out

    def _eager(self, inputs):
        self.optimizer.use_device_hyper(None)
        out = self.trainer._device_step(self.model, inputs, self.optimizer)
        self.scheduler.step()
        if self.graph is not None:
            self.optimizer.use_device_hyper(self.hyper)
        return out


class InputPrefetcher:
    """Host -> device copy of the NEXT batch on a side stream while the current step runs.

    `stage(batch)` starts the copies from (pinned) host memory into a device staging set and returns a
    handle; `take(handle)` makes the compute stream wait for them and hands the device tensors to the step
    (GraphedStep then copies them device-to-device into its static buffers, ~25 us for an 82 MB batch).
    Two staging sets alternate so a batch is never overwritten while a step still reads it."""

    def __init__(self, device):
        self.device = device
        self.stream = torch.cuda.Stream(device=device)
        self.slots = [None, None]
        self.events = [torch.cuda.Event(), torch.cuda.Event()]
        self.done = [None, None]
        self.turn = 0

    def stage(self, batch):
        i = self.turn
        self.turn ^= 1
        if self.done[i] is not None:
            self.stream.wait_event(self.done[i])      # the step that consumed this slot has finished
        with torch.cuda.stream(self.stream):
            out = []
            for j, t in enumerate(batch):
                if not torch.is_tensor(t):
                    out.append(t)
                    continue
                slot = self.slots[i][j] if self.slots[i] is not None and j < len(self.slots[i]) else None
                if slot is None or slot.shape != t.shape or slot.dtype != t.dtype:
                    slot = torch.empty(t.shape, dtype=t.dtype, device=self.device)
                slot.copy_(t, non_blocking=True)
                out.append(slot)
            self.slots[i] = out
            self.events[i].record(self.stream)
        return i

    def take(self, handle):
        cur = torch.cuda.current_stream()
        cur.wait_event(self.events[handle])
        return self.slots[handle]

    def release(self, handle):
        """Call after the step that used `handle` has been enqueued."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.done[handle] = ev
