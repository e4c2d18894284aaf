"""Result / batch containers of the boundary (mirrors of pythae's ModelOutput / DatasetOutput and of
multivae.data.datasets.base — /root/reference/src/multivae/data/datasets/base.py:8,97)."""
from collections import OrderedDict

import torch


class _AttrDict(OrderedDict):
    """OrderedDict whose items are also attributes (both directions), like pythae's ModelOutput."""

    def __getitem__(self, k):
        if isinstance(k, str):
            return OrderedDict.__getitem__(self, k)
        return self.to_tuple()[k]

    def __setattr__(self, name, value):
        OrderedDict.__setitem__(self, name, value)
        OrderedDict.__setattr__(self, name, value)

    def __setitem__(self, key, value):
        OrderedDict.__setitem__(self, key, value)
        OrderedDict.__setattr__(self, key, value)

    def to_tuple(self):
        return tuple(OrderedDict.__getitem__(self, k) for k in self.keys())


class ModelOutput(_AttrDict):
    pass


class DatasetOutput(_AttrDict):
    pass


class MultimodalBaseDataset:
    """data: dict modality -> tensor (N, *dims); optional labels."""

    def __init__(self, data, labels=None):
        self.labels = labels
        self.data = data

    def __len__(self):
        length = len(self.data[list(self.data)[0]])
        for m in self.data:
            if len(self.data[m]) != length:
                raise AttributeError("The size of the provided datasets doesn't correspond between modalities!")
        return length

    def __getitem__(self, index):
        X = {m: self.data[m][index] for m in self.data}
        if self.labels is not None:
            return DatasetOutput(data=X, labels=self.labels[index])
        return DatasetOutput(data=X)


class IncompleteDataset(MultimodalBaseDataset):
    """Adds `masks`: dict modality -> bool tensor (N,) telling which samples are available."""

    def __init__(self, data, masks, labels=None):
        super().__init__(data, labels)
        self.masks = masks

    def __getitem__(self, index):
        X = {m: self.data[m][index] for m in self.data}
        Mk = {m: self.masks[m][index] for m in self.masks}
        if self.labels is not None:
            return DatasetOutput(data=X, masks=Mk, labels=self.labels[index])
        return DatasetOutput(data=X, masks=Mk)


def drop_unused_modalities(inputs):
    """Drop modalities unavailable for the entire batch (data/utils.py:53-63)."""
    if not hasattr(inputs, "masks"):
        return inputs
    for m in list(inputs.masks.keys()):
        if not bool(inputs.masks[m].any()):
            inputs.data.pop(m)
            inputs.masks.pop(m)
    return inputs


def set_inputs_to_device(inputs, device):
    """Move every tensor of a batch dict to `device` (non_blocking; keeps masks and labels, unlike
    the reference's CUDA branch, data/utils.py:15-18, which drops them)."""

    def mv(v):
        if torch.is_tensor(v):
            return v.to(device, non_blocking=True)
        if isinstance(v, dict):
            return {k: mv(x) for k, x in v.items()}
        return v

    return DatasetOutput(**{k: mv(v) for k, v in inputs.items()})
