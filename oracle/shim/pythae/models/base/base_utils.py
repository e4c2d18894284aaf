"""ModelOutput / CPU_Unpickler stand-ins (containers only)."""
import io
import pickle
from collections import OrderedDict
from typing import Any, Tuple

import torch


class ModelOutput(OrderedDict):
    """OrderedDict whose items are also attributes (missing attribute -> AttributeError... the
    real pythae class returns attribute access for set keys only)."""

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


class CPU_Unpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module == "torch.storage" and name == "_load_from_bytes":
            return lambda b: torch.load(io.BytesIO(b), map_location="cpu")
        return super().find_class(module, name)
