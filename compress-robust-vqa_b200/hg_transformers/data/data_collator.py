"""Batch collation (reference hg_transformers/data/data_collator.py:27-97).  The VQA datasets yield
8-tuples (ids, feats, pos, target, qid, img_id, bias, max_label); TrimCollator stacks each field and
pads region features to the longest example."""
from abc import ABC, abstractmethod
from collections.abc import Mapping, Sequence

import numpy as np
import torch
import torch.nn.functional as F
from torch.utils.data.dataloader import default_collate


class DataCollator(ABC):
    @abstractmethod
    def collate_batch(self, batch):
        pass


class DefaultDataCollator(DataCollator):
    def collate_batch(self, batch):
        return default_collate(batch)


class TrimCollator(DataCollator):
    def collate_batch(self, batch):
        first = batch[0]
        if torch.is_tensor(first):
            if first.dim() > 1:  # image features: pad the box dimension
                longest = max(x.size(0) for x in batch)
                return torch.stack([F.pad(x, (0, 0, 0, longest - x.size(0))) for x in batch], 0)
            return torch.stack(batch, 0)
        if isinstance(first, np.ndarray):
            return torch.stack([torch.from_numpy(b) for b in batch], 0)
        if isinstance(first, (int, np.integer)):
            return torch.LongTensor([int(b) for b in batch])
        if isinstance(first, (float, np.floating)):
            return torch.DoubleTensor([float(b) for b in batch])
        if isinstance(first, (str, bytes)):
            return batch
        if isinstance(first, Mapping):
            return {k: default_collate([d[k] for d in batch]) for k in first}
        if isinstance(first, Sequence):
            return [self.collate_batch(list(samples)) for samples in zip(*batch)]
        raise TypeError(f"batch must contain tensors, numbers, dicts or lists; found {type(first)}")
