"""Where does the end-to-end loop of bench.py lose time against the resident-input loop at N > 1?  Same workload and
engine as bench.py (BASELINE config 2), four loops timed with CUDA events (max over ranks), 20 steps each:
  resident            inputs in HBM, no host sync                      (bench.py `value`)
  resident+readback   inputs in HBM, loss read back every step (pipelined one step)
  h2d                 82 MB host -> device per step through InputPrefetcher, no read-back
  h2d+readback        both                                             (bench.py `e2e`)
    python tests/e2e_probe.py            |  python -m torch.distributed.run --nproc-per-node N ... tests/e2e_probe.py
"""
import json
import os
import sys
import types

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "compress-robust-vqa_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402


def main():
    from hg_transformers._engine import InputPrefetcher
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    args = types.SimpleNamespace(batch=256, ans_num=3129, loss="lpf", config="lxmert")
    wl = bench.build_workload(args, dev, world, local, rank)
    trainer, model, optimizer, scheduler = wl["trainer"], wl["model"], wl["optimizer"], wl["scheduler"]
    host = [t.pin_memory() for t in wl["host"]]
    resident = [t.to(dev) for t in host]
    graphed = trainer._make_graphed_step(model, optimizer, scheduler)
    trainer._zero_grad(optimizer)
    while graphed.graph is None:
        graphed.step(resident)
    for _ in range(5):
        graphed.step(resident)
    pre = InputPrefetcher(dev)
    loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
    loss_ev = [torch.cuda.Event(blocking=True) for _ in range(2)]   # sleep, do not spin: N ranks share the host cores

    def loop(n, h2d, readback):
        handle = pre.stage(host) if h2d else None
        for i in range(n):
            if h2d:
                batch, cur = pre.take(handle), handle
                handle = pre.stage(host) if i + 1 < n else None
            else:
                batch = resident
            loss, _ = graphed.step(batch)
            if h2d:
                pre.release(cur)
            if readback:
                loss_host[i & 1].copy_(loss.detach().reshape(()), non_blocking=True)
                loss_ev[i & 1].record()
                if i > 0:
                    loss_ev[(i - 1) & 1].synchronize()
                    float(loss_host[(i - 1) & 1])
        if readback:
            loss_ev[(n - 1) & 1].synchronize()

    def timed(h2d, readback, n=20):
        loop(3, h2d, readback)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loop(n, h2d, readback)
        e1.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / n], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    out = {"world": world}
    for name, h2d, rb in (("resident", False, False), ("resident+readback", False, True), ("h2d", True, False),
                          ("h2d+readback", True, True), ("resident again", False, False)):
        out[name] = round(timed(h2d, rb), 3)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        torch.cuda.synchronize()
        dist.barrier()
        os._exit(0)


if __name__ == "__main__":
    main()
