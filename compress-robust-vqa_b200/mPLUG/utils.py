"""Small helpers of the mPLUG driver (reference mPLUG/utils.py): the attribute dict the YAML blocks are wrapped in, the
process-group queries, ``init_distributed_mode`` (one process per GPU, NCCL; rendezvous from the torchrun environment)
and a compact ``MetricLogger`` / ``SmoothedValue`` pair with the interface the loops use (``add_meter``, ``update``,
``log_every``, ``meters[...]``, ``global_avg``, ``synchronize_between_processes``)."""
import os
import time
from collections import defaultdict, deque

import torch
import torch.distributed as dist


class AttrDict(dict):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.__dict__ = self


def is_dist_avail_and_initialized():
    return dist.is_available() and dist.is_initialized()


def get_world_size():
    return dist.get_world_size() if is_dist_avail_and_initialized() else 1


def get_rank():
    return dist.get_rank() if is_dist_avail_and_initialized() else 0


def is_main_process():
    return get_rank() == 0


def save_on_master(*args, **kwargs):
    if is_main_process():
        torch.save(*args, **kwargs)


def compute_n_params(model, return_str=True):
    tot = sum(p.numel() for p in model.parameters())
    if not return_str:
        return tot
    if tot >= 1e6:
        return "{:.1f}M".format(tot / 1e6)
    return "{:.1f}K".format(tot / 1e3)


def init_distributed_mode(args):
    """RANK / WORLD_SIZE / LOCAL_RANK from the launcher -> args.{rank, world_size, gpu, distributed}, set the device and
    join the NCCL group (gloo when no GPU exists, for host-side tests)."""
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        args.rank = int(os.environ["RANK"])
        args.world_size = int(os.environ["WORLD_SIZE"])
        args.gpu = int(os.environ.get("LOCAL_RANK", 0))
    else:
        print("Not using distributed mode")
        args.distributed = False
        return
    args.distributed = True
    if torch.cuda.is_available():
        torch.cuda.set_device(args.gpu)
        args.dist_backend = "nccl"
    else:
        args.dist_backend = "gloo"
    print("| distributed init (rank {}): {}".format(args.rank, getattr(args, "dist_url", "env://")), flush=True)
    dist.init_process_group(backend=args.dist_backend, init_method=getattr(args, "dist_url", "env://"),
                            world_size=args.world_size, rank=args.rank)
    dist.barrier()


class SmoothedValue(object):
    """Windowed and global statistics of a series."""

    def __init__(self, window_size=20, fmt=None):
        self.deque = deque(maxlen=window_size)
        self.total, self.count = 0.0, 0
        self.fmt = fmt or "{median:.4f} ({global_avg:.4f})"

    def update(self, value, n=1):
        self.deque.append(value)
        self.count += n
        self.total += value * n

    def synchronize_between_processes(self):
        if not is_dist_avail_and_initialized():
            return
        t = torch.tensor([self.count, self.total], dtype=torch.float64,
                         device="cuda" if dist.get_backend() == "nccl" else "cpu")
        dist.barrier()
        dist.all_reduce(t)
        self.count, self.total = int(t[0].item()), t[1].item()

    @property
    def median(self):
        return torch.tensor(list(self.deque)).median().item()

    @property
    def avg(self):
        return torch.tensor(list(self.deque), dtype=torch.float32).mean().item()

    @property
    def global_avg(self):
        return self.total / self.count

    @property
    def max(self):
        return max(self.deque)

    @property
    def value(self):
        return self.deque[-1]

    def __str__(self):
        return self.fmt.format(median=self.median, avg=self.avg, global_avg=self.global_avg, max=self.max,
                               value=self.value)


class MetricLogger(object):
    def __init__(self, delimiter="\t"):
        self.meters = defaultdict(SmoothedValue)
        self.delimiter = delimiter

    def update(self, **kwargs):
        for k, v in kwargs.items():
            self.meters[k].update(v.item() if isinstance(v, torch.Tensor) else v)

    def add_meter(self, name, meter):
        self.meters[name] = meter

    def __str__(self):
        return self.delimiter.join("{}: {}".format(n, m) for n, m in self.meters.items())

    def global_avg(self):
        return self.delimiter.join("{}: {:.4f}".format(n, m.global_avg) for n, m in self.meters.items())

    def synchronize_between_processes(self):
        for meter in self.meters.values():
            meter.synchronize_between_processes()

    def log_every(self, iterable, print_freq, header=""):
        start = time.time()
        for i, obj in enumerate(iterable):
            yield obj
            if i % print_freq == 0:
                print(self.delimiter.join([header, "[{}/{}]".format(i, len(iterable)), str(self),
                                           "time: {:.1f}s".format(time.time() - start)]))
