"""Batch assembly for graph pairs (reference loaders/loaders.py:5-23)."""
import torch
from torch.utils.data import DataLoader

from ..maskedtensors import maskedtensor


def collate_fn_pair(samples_list):
    """Ragged batch: two MaskedTensors (base names 'N' and 'M')."""
    first = [a for a, _ in samples_list]
    second = [b for _, b in samples_list]
    return (maskedtensor.from_list(first, dims=(1, 2), base_name='N'),
            maskedtensor.from_list(second, dims=(1, 2), base_name='M'))


def collate_fn_pair_explore(samples_list):
    """Constant-size batch: two {'input': (B,F,N,N)} dicts."""
    first = torch.stack([a for a, _ in samples_list])
    second = torch.stack([b for _, b in samples_list])
    return {'input': first}, {'input': second}


def siamese_loader(data, batch_size, constant_n_vertices, shuffle=True, num_workers=4):
    assert len(data) > 0
    fn = collate_fn_pair_explore if constant_n_vertices else collate_fn_pair
    return DataLoader(data, batch_size=batch_size, shuffle=shuffle, num_workers=num_workers, collate_fn=fn)
