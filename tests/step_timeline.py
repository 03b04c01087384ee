"""Device timeline of the replayed training step (BASELINE config 2, batch 256 / GPU) on 1 .. 8 GPUs: which stream
runs what, how long the NCCL kernels run, and how much of them is EXPOSED (no compute kernel of ours running beside
them on this rank).  Kineto / CUPTI trace of two graph replays on rank 0 -- timestamps come from the device.

    python tests/step_timeline.py out_prefix                      # 1 GPU
    python -m torch.distributed.run --nproc-per-node N ... tests/step_timeline.py out_prefix

Writes <out_prefix>_N<world>.json (summary) and <out_prefix>_N<world>_kernels.csv (start_us, dur_us, stream, name of
every kernel of the second profiled step).
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "compress-robust-vqa_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402


def main():
    from hg_transformers.data.data_collator import TrimCollator
    from hg_transformers.data.metrics import vqa_compute_metrics
    from hg_transformers.mask_trainer_Robust_VQA import Trainer
    from hg_transformers.training_args import TrainingArguments
    from prune_debias_VQA import batch_tuple, build_stage2, init_optimizer, synthetic_batch
    prefix = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "timeline")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, A = 256, 3129
    targs = TrainingArguments(output_dir=os.path.join(ROOT, "gpurun_out", "timeline_out"), per_gpu_train_batch_size=B,
                              logging_steps=100, seed=49, Masker_type="lpf", training_type="Masker", save_steps=0,
                              local_rank=local if world > 1 else -1, dataloader_num_workers=0)
    model, masker, margs = build_stage2(A, device=dev, seed=49)
    optimizer, scheduler = init_optimizer(model, targs, num_train_data=B * world * 10000)
    trainer = Trainer(model=model, args=targs, model_args=margs, data_collator=TrimCollator(), train_dataset=None,
                      compute_metrics=vqa_compute_metrics, optimizers=(optimizer, scheduler), masker=masker)
    trainer._setup_engine(optimizer)
    trainer.global_step = 0
    trainer._zero_grad(optimizer)
    inputs = [t.to(dev) for t in batch_tuple(synthetic_batch(B, A, seed=49 + rank))]
    gs = trainer._make_graphed_step(model, optimizer, scheduler)
    while gs.graph is None:
        gs.step(inputs)
    for _ in range(5):
        gs.step(inputs)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(3):
            gs.step(inputs)
        torch.cuda.synchronize()
    if rank == 0:
        trace = f"{prefix}_N{world}_trace.json"
        prof.export_chrome_trace(trace)
        with open(trace) as f:
            tr = json.load(f)
        os.remove(trace)
        ks = sorted(((float(e["ts"]), float(e["dur"]), (e.get("args") or {}).get("stream", 0), e["name"])
                     for e in tr["traceEvents"] if e.get("cat") in ("kernel", "gpu_memcpy", "gpu_memset")
                     and "ts" in e and "dur" in e), key=lambda k: k[0])
        # split into steps at the rng counter kernel that opens every step
        starts = [i for i, k in enumerate(ks) if "counter_inc" in k[3]]
        step = ks[starts[1]:starts[2]] if len(starts) >= 3 else ks
        t0 = step[0][0]
        t1 = max(s + d for s, d, _, _ in step)
        is_nccl = lambda n: "nccl" in n.lower()
        comp = [(s, s + d) for s, d, _, n in step if not is_nccl(n)]
        comm = [(s, s + d, n) for s, d, _, n in step if is_nccl(n)]

        def union(iv):
            iv = sorted(iv)
            out = []
            for a, b in iv:
                if out and a <= out[-1][1]:
                    out[-1][1] = max(out[-1][1], b)
                else:
                    out.append([a, b])
            return out

        cu = union(comp)
        busy = sum(b - a for a, b in cu)

        def exposed(a, b):          # part of [a, b) not covered by compute kernels
            cov = 0.0
            for x, y in cu:
                lo, hi = max(a, x), min(b, y)
                if hi > lo:
                    cov += hi - lo
            return (b - a) - cov

        comm_total = sum(b - a for a, b, _ in comm)
        comm_exposed = sum(exposed(a, b) for a, b, _ in comm)
        tail = [(a - t0, b - a, exposed(a, b), n[:60]) for a, b, n in comm if exposed(a, b) > 5.0]
        by_stream = {}
        for s, d, st, n in step:
            by_stream.setdefault(str(st), [0, 0.0])
            by_stream[str(st)][0] += 1
            by_stream[str(st)][1] += d
        summary = {"world": world, "step_us": t1 - t0, "kernels": len(step), "compute_busy_us": busy,
                   "idle_us": (t1 - t0) - busy - comm_exposed, "nccl_kernels": len(comm), "nccl_total_us": comm_total,
                   "nccl_exposed_us": comm_exposed,
                   "exposed_nccl_segments": [{"at_us": round(a, 1), "dur_us": round(d, 1), "exposed_us": round(x, 1),
                                              "kernel": n} for a, d, x, n in tail],
                   "streams": {k: {"kernels": v[0], "busy_us": round(v[1], 1)} for k, v in by_stream.items()},
                   "dp_mode": os.environ.get("CRVQA_DP", "sharded")}
        with open(f"{prefix}_N{world}.json", "w") as f:
            json.dump(summary, f, indent=1)
        with open(f"{prefix}_N{world}_kernels.csv", "w") as f:
            f.write("start_us,dur_us,stream,kernel\n")
            for s, d, st, n in step:
                f.write(f"{s - t0:.1f},{d:.1f},{st},{n[:100].replace(',', ';')}\n")
        print(json.dumps(summary)[:1500])
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)


if __name__ == "__main__":
    main()
