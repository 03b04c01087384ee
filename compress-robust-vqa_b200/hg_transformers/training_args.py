"""TrainingArguments (reference hg_transformers/training_args.py:30-172 plus the stage-2 additions of
prune_debias_VQA.py:339-365).  Device policy is B200-first: ONE process per GPU.  `local_rank == -1`
means a single process on the current CUDA device (the reference would wrap the model in
nn.DataParallel over all visible GPUs -- that mode, with its per-step parameter broadcast, is
deliberately not reproduced); under torchrun every rank drives its own GPU over NCCL."""
import dataclasses
import json
import os
from dataclasses import dataclass, field
from typing import Any, Dict, Optional

import torch


def is_tpu_available():
    return False


@dataclass
class TrainingArguments:
    output_dir: str = field(default="./out")
    overwrite_output_dir: bool = False
    do_train: bool = False
    do_eval: bool = False
    do_predict: bool = False
    evaluate_during_training: bool = False
    per_gpu_train_batch_size: int = 8
    per_gpu_eval_batch_size: int = 8
    gradient_accumulation_steps: int = 1
    learning_rate: float = 5e-5
    gamma: float = 5
    weight_decay: float = 0.0
    adam_epsilon: float = 1e-8
    max_grad_norm: float = 1.0
    num_train_epochs: float = 3.0
    max_steps: int = -1
    warmup_steps: int = 0
    logging_dir: Optional[str] = None
    logging_first_step: bool = False
    logging_steps: int = 500
    save_steps: int = 500
    save_total_limit: Optional[int] = None
    no_cuda: bool = False
    seed: int = 42
    fp16: bool = False
    fp16_opt_level: str = "O1"
    local_rank: int = -1
    tpu_num_cores: Optional[int] = None
    tpu_metrics_debug: bool = False
    # stage-2 additions (prune_debias_VQA.py:339-365)
    use_kd: bool = False
    training_type: Optional[str] = "Masker"
    Masker_type: Optional[str] = None
    FTmodel_type: Optional[str] = None
    FT_type: Optional[str] = None
    label4save: Optional[str] = None
    # B200 engine switches (no counterpart in the reference)
    dist_backend: str = "nccl"
    dataloader_num_workers: int = 1

    @property
    def train_batch_size(self) -> int:
        return self.per_gpu_train_batch_size * max(1, self.n_gpu)

    @property
    def eval_batch_size(self) -> int:
        return self.per_gpu_eval_batch_size * max(1, self.n_gpu)

    def _setup_devices(self):
        cached = self.__dict__.get("_devices")
        if cached is not None:
            return cached
        rank = self.local_rank
        if rank == -1 and int(os.environ.get("WORLD_SIZE", "1")) > 1 and "LOCAL_RANK" in os.environ:
            rank = int(os.environ["LOCAL_RANK"])  # launched by torchrun without --local_rank
            self.local_rank = rank
        if self.no_cuda or not torch.cuda.is_available():
            device, n_gpu = torch.device("cpu"), 0
            if rank != -1 and not torch.distributed.is_initialized():
                torch.distributed.init_process_group(backend="gloo")
        elif rank == -1:
            device, n_gpu = torch.device("cuda", torch.cuda.current_device()), 1
        else:
            torch.cuda.set_device(rank)
            if not torch.distributed.is_initialized():
                torch.distributed.init_process_group(backend=self.dist_backend)
            device, n_gpu = torch.device("cuda", rank), 1
        self.__dict__["_devices"] = (device, n_gpu)
        return device, n_gpu

    @property
    def device(self) -> "torch.device":
        return self._setup_devices()[0]

    @property
    def n_gpu(self):
        return self._setup_devices()[1]

    def to_json_string(self):
        return json.dumps(dataclasses.asdict(self), indent=2)

    def to_sanitized_dict(self) -> Dict[str, Any]:
        d = dataclasses.asdict(self)
        valid = (bool, int, float, str, torch.Tensor)
        return {k: v if isinstance(v, valid) else str(v) for k, v in d.items()}
