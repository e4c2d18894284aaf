"""pythae.config.BaseConfig stand-in: pydantic dataclass with JSON helpers (containers only)."""
import json
import os
from dataclasses import asdict, field
from typing import Any, Dict

from pydantic.dataclasses import dataclass


@dataclass
class BaseConfig:
    name: str = field(init=False)

    def __post_init__(self):
        self.name = self.__class__.__name__

    @classmethod
    def from_dict(cls, config_dict: Dict[str, Any]):
        config_dict = dict(config_dict)
        config_dict.pop("name", None)
        return cls(**config_dict)

    @classmethod
    def from_json_file(cls, json_path):
        with open(json_path) as f:
            d = json.load(f)
        name = d.pop("name", None)
        if name is not None and name != cls.__name__:
            raise ValueError(f"config file is for {name}, not {cls.__name__}")
        return cls.from_dict(d)

    def to_dict(self):
        return asdict(self)

    def to_json_string(self):
        return json.dumps(self.to_dict(), default=str)

    def save_json(self, dir_path, filename):
        with open(os.path.join(dir_path, f"{filename}.json"), "w", encoding="utf-8") as fp:
            fp.write(self.to_json_string())
