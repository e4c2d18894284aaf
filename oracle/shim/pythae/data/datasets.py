"""DatasetOutput / BaseDataset stand-ins (containers only)."""
from collections import OrderedDict
from typing import Any, Tuple

import torch
from torch.utils.data import Dataset  # noqa


class DatasetOutput(OrderedDict):
    def __getitem__(self, k):
        if isinstance(k, str):
            return dict(self.items())[k]
        return self.to_tuple()[k]

    def __setattr__(self, name, value):
        super().__setitem__(name, value)
        super().__setattr__(name, value)

    def __setitem__(self, key, value):
        super().__setitem__(key, value)
        super().__setattr__(key, value)

    def to_tuple(self) -> Tuple[Any]:
        return tuple(self[k] for k in self.keys())


class BaseDataset(Dataset):
    def __init__(self, data, labels):
        self.data = data
        self.labels = labels

    def __len__(self):
        return len(self.data)

    def __getitem__(self, index):
        return DatasetOutput(data=self.data[index], labels=self.labels[index])
