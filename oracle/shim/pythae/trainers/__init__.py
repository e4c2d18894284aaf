"""Import-time placeholders (samplers only; not on the hot path)."""


class BaseTrainer:
    def __init__(self, *a, **k):
        raise NotImplementedError("oracle shim placeholder")


class BaseTrainerConfig:
    def __init__(self, *a, **k):
        raise NotImplementedError("oracle shim placeholder")
