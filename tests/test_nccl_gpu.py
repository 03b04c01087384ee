"""Multi-GPU correctness on NCCL (layer 4): data-parallel stage-2 training over 2 B200s must be
the reference's DDP semantics (hg_transformers/mask_trainer_VQA.py:537-543: gradients averaged over ranks, every rank
holds the same scores, thresholds and masks).  Skipped when fewer than two GPUs are visible (the driver's 1-GPU test
tier); run with `gpurun --gpus 2 -- python -m pytest tests/test_nccl_gpu.py -m gpu -q`."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

CFG = dict(vocab_size=1000, hidden_size=256, num_attention_heads=4, intermediate_size=512, l_layers=2, x_layers=2,
           r_layers=1, visual_feat_dim=128, max_position_embeddings=32, hidden_dropout_prob=0.0,
           attention_probs_dropout_prob=0.0)
A, B_GLOBAL, STEPS = 96, 32, 3


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _make_trainer(tmp, local_rank, world):
    from hg_transformers.data.data_collator import TrimCollator
    from hg_transformers.data.metrics import vqa_compute_metrics
    from hg_transformers.mask_trainer_Robust_VQA import Trainer
    from hg_transformers.optimization import get_constant_schedule
    from hg_transformers.training_args import TrainingArguments
    from optimization import AdamW
    from prune_debias_VQA import build_stage2
    targs = TrainingArguments(output_dir=tmp, per_gpu_train_batch_size=B_GLOBAL // world, logging_steps=1000, seed=49,
                              Masker_type="lpf", training_type="Masker", save_steps=0, dataloader_num_workers=0,
                              local_rank=local_rank if world > 1 else -1)
    model, masker, margs = build_stage2(A, device=targs.device, seed=49, config_kwargs=CFG)
    model.classifier.main[2].p = 0.0
    opt = AdamW([{"params": [p]} for p in model.parameters() if p.requires_grad], lr=2e-3, eps=1e-8)
    sched = get_constant_schedule(opt)
    trainer = Trainer(model=model, args=targs, model_args=margs, data_collator=TrimCollator(), train_dataset=None,
                      compute_metrics=vqa_compute_metrics, optimizers=(opt, sched), masker=masker)
    trainer._setup_engine(opt)
    trainer.global_step = 0
    trainer._zero_grad(opt)
    return trainer, model, opt, sched


def _global_batch(step):
    from prune_debias_VQA import batch_tuple, synthetic_batch
    return batch_tuple(synthetic_batch(B_GLOBAL, A, seed=100 + step, tokens=10, regions=8, feat_dim=128, vocab=1000))


def _shard(batch, rank, world):
    n = B_GLOBAL // world
    return [t[rank * n:(rank + 1) * n].cuda() for t in batch]


def _run(trainer, model, opt, sched, rank, world, graph):
    """STEPS optimiser steps + reset_threshold; returns (first-step gradient arena, scores, thresholds) on the CPU."""
    gs = trainer._make_graphed_step(model, opt, sched) if graph else None
    first_grads = None
    for step in range(STEPS):
        inputs = _shard(_global_batch(step), rank, world)
        if gs is not None:
            gs.step(inputs)
        else:
            trainer._device_step(model, inputs, opt)
            sched.step()
        if step == 0:
            torch.cuda.synchronize()
            first_grads = _assembled_grads(trainer)
    trainer.reset_threshold(model, 0.7)
    if trainer.grad_sync is not None:
        trainer.grad_sync.make_consistent()
    torch.cuda.synchronize()
    arena = trainer.arena
    return first_grads, arena.scores.detach().cpu().clone(), arena.thr_vec.detach().cpu().clone()


def _assembled_grads(trainer):
    """The exchanged (mean) gradient of the whole arena on this rank.  Sharded optimiser: a rank holds the mean only
    for the slices it owns (reduce-scatter), so the owned slices are all-gathered into a copy."""
    import torch.distributed as dist
    g = trainer.arena.grads.detach().clone()
    sync = trainer.grad_sync
    if sync is not None and sync.sharded:
        for b, sh in enumerate(sync.bucket_sharded):
            if sh:
                lo, hi = sync.bucket_ranges[b]
                olo, ohi = sync._own(b)
                dist.all_gather_into_tensor(g[lo:hi], g[olo:ohi].clone())
    return g.cpu()


def _worker(rank, world, port, tmp, out_path, graph):
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.path.join(root, "compress-robust-vqa_b200"), root):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank), CRVQA_CUDA_GRAPH="1" if graph else "0", CRVQA_KEEP_GRADS="1")
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        trainer, model, opt, sched = _make_trainer(os.path.join(tmp, f"r{rank}"), rank, world)
        # local (un-exchanged) first-step gradient of this rank, from an engine whose exchange is switched off
        sync = trainer.grad_sync
        sync.enabled = False
        trainer._training_step(model, _shard(_global_batch(0), rank, world), opt)
        sync.finish([])
        torch.cuda.synchronize()
        local = trainer.arena.grads.detach().clone()
        for p in trainer._loose_params():
            p.grad = None
        trainer._zero_grad(opt)
        sync.enabled = True
        grads, scores, thr = _run(trainer, model, opt, sched, rank, world, graph)
        # mean over ranks of the local gradients, by an independent collective
        dist.all_reduce(local, op=dist.ReduceOp.SUM)
        local /= world
        torch.save({"grads": grads, "scores": scores, "thr": thr, "mean_local": local.cpu()}, f"{out_path}.{rank}")
        dist.barrier()
    finally:
        torch.cuda.synchronize()
        os._exit(0) if graph else dist.destroy_process_group()


@pytest.mark.parametrize("graph,mode", [(False, "sharded"), (True, "sharded"), (False, "allreduce")])
def test_two_gpu_nccl_matches_ddp_semantics(tmp_path, graph, mode, monkeypatch):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    monkeypatch.setenv("CRVQA_DP", mode)          # inherited by the spawned ranks
    out = str(tmp_path / "res")
    ctx = mp.get_context("spawn")
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, str(tmp_path), out, graph)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
        assert p.exitcode == 0, p.exitcode
    r0, r1 = (torch.load(f"{out}.{r}") for r in range(2))
    # every rank ends with the same scores, thresholds (hence masks): bit for bit
    assert torch.equal(r0["scores"], r1["scores"])
    assert torch.equal(r0["thr"], r1["thr"])
    assert torch.equal(r0["grads"], r1["grads"])
    # exchanged gradient == mean of the per-rank gradients (split-K reduce order differs run to run: 1e-5 norm-wise)
    num = float((r0["grads"] - r0["mean_local"]).double().norm())
    den = float(r0["mean_local"].double().norm())
    print(f"[nccl graph={graph} {mode}] exchanged vs mean-of-local gradient: {num / den:.3e}")
    assert num / den < 1e-4
    # 1 GPU on the same GLOBAL batch: same thresholds / masks up to bf16 tiling noise of the gradients
    os.environ["CRVQA_CUDA_GRAPH"] = "0"
    os.environ["CRVQA_KEEP_GRADS"] = "1"
    try:
        trainer, model, opt, sched = _make_trainer(str(tmp_path / "single"), 0, 1)
        g1, s1, t1 = _run(trainer, model, opt, sched, 0, 1, False)
    finally:
        os.environ.pop("CRVQA_CUDA_GRAPH", None)
        os.environ.pop("CRVQA_KEEP_GRADS", None)
    rel = float((g1 - r0["grads"]).double().norm() / g1.double().norm())
    print(f"[nccl graph={graph}] 1-GPU vs 2-GPU first-step gradient (same global batch): {rel:.3e}")
    assert rel < 3e-2
    offs = trainer.arena.offsets
    agree = total = 0
    for i, m in enumerate(trainer.arena.modules):
        n = m.weight_mask.numel()
        a = s1[offs[i]: offs[i] + n] > t1[i]
        b = r0["scores"][offs[i]: offs[i] + n] > r0["thr"][i]
        agree += int((a == b).sum())
        total += n
    print(f"[nccl graph={graph}] masks after {STEPS} steps, 1 GPU vs 2 GPUs: {agree}/{total} equal")
    assert agree / total > 0.995
