"""Host -> device bandwidth of the bench's input batch (82 MB of pinned tensors) when N ranks copy at the same time.
    python tests/h2d_probe.py   |   python -m torch.distributed.run --nproc-per-node N ... tests/h2d_probe.py"""
import json, os, sys
import torch
import torch.distributed as dist
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
host = torch.randn(256, 36, 2048).pin_memory()
dst = torch.empty_like(host, device=dev)
out = {}
for name, src in (("fp32 75.5 MB", host), ("bf16 37.7 MB", host.bfloat16().pin_memory())):
    d = torch.empty_like(src, device=dev)
    for _ in range(3):
        d.copy_(src, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        d.copy_(src, non_blocking=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    out[name] = {"ms": round(ms, 3), "GBps": round(src.numel() * src.element_size() / ms / 1e6, 1)}
if world > 1:
    g = [None] * world
    dist.all_gather_object(g, out)
    out = {f"rank{i}": v for i, v in enumerate(g)}
if rank == 0:
    print(json.dumps({"world": world, "h2d": out}))
