"""`torch_geometric.loader.DataLoader` (train.py:150-162): torch's DataLoader with PyG's collate rule -- graph objects
are batched with `Batch.from_data_list`, mappings key by key, tensors / numbers by torch's default collate.  The
reference's `SequenceDataset` yields dicts of tensors (dataloader.py:153-160), so only the mapping branch is used."""
from collections.abc import Mapping, Sequence

import torch
from torch.utils.data import DataLoader as _TorchDataLoader
from torch.utils.data.dataloader import default_collate

from cosmology_gnn_simulation_b200.graph import Batch, Data


def _collate(batch):
    elem = batch[0]
    if isinstance(elem, Data):
        return Batch.from_data_list(batch)
    if isinstance(elem, torch.Tensor):
        return default_collate(batch)
    if isinstance(elem, Mapping):
        return {key: _collate([d[key] for d in batch]) for key in elem}
    if isinstance(elem, tuple) and hasattr(elem, "_fields"):
        return type(elem)(*(_collate(list(s)) for s in zip(*batch)))
    if isinstance(elem, Sequence) and not isinstance(elem, (str, bytes)):
        return [_collate(list(s)) for s in zip(*batch)]
    return default_collate(batch)


class DataLoader(_TorchDataLoader):
    def __init__(self, dataset, batch_size: int = 1, shuffle: bool = False, follow_batch=None, exclude_keys=None, **kwargs):
        kwargs.pop("collate_fn", None)
        super().__init__(dataset, batch_size, shuffle, collate_fn=_collate, **kwargs)
