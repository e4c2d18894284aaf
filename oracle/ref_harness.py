"""Import the REAL reference (/root/reference/src) through the container-only shim and inject
recorded sampling noise.  BUILD-CONTAINER ONLY test infrastructure: /root/reference does not exist
on the GPU box, so nothing under `-m gpu`, smoke() or bench.py may import this module."""
import contextlib
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src"


def available():
    return os.path.isdir(REF_SRC)


def import_reference():
    if not available():
        raise RuntimeError("reference sources not present (this only works in the build container)")
    for p in (os.path.join(HERE, "shim"), REF_SRC):
        if p not in sys.path:
            sys.path.insert(0, p)
    import multivae  # noqa: F401

    return multivae


class NoiseQueue:
    """FIFO of standard draws `e`; the patched rsample returns loc + scale * e, bit-identical to
    torch.distributions' own formula given the same base draw (see oracle/port/elbo.py header)."""

    def __init__(self, draws=None, record=False, generator=None):
        self.draws = list(draws or [])
        self.record = record
        self.generator = generator
        self.log = []

    def next(self, shape, kind, dtype):
        if self.record:
            if kind == "laplace":
                finfo = torch.finfo(dtype)
                u = torch.empty(shape, dtype=dtype).uniform_(finfo.eps - 1, 1, generator=self.generator)
                e = -u.sign() * torch.log1p(-u.abs().clamp(min=finfo.tiny))
            else:
                e = torch.empty(shape, dtype=dtype).normal_(generator=self.generator)
            self.log.append(e)
            return e
        e = self.draws.pop(0)
        assert tuple(e.shape) == tuple(shape), (e.shape, shape)
        return e.to(dtype)


@contextlib.contextmanager
def injected_noise(queue: NoiseQueue):
    from torch.distributions import Laplace, Normal

    orig_l, orig_n = Laplace.rsample, Normal.rsample

    def l_rsample(self, sample_shape=torch.Size()):
        shape = self._extended_shape(torch.Size(sample_shape))
        return self.loc + self.scale * queue.next(shape, "laplace", self.loc.dtype)

    def n_rsample(self, sample_shape=torch.Size()):
        shape = self._extended_shape(torch.Size(sample_shape))
        return self.loc + queue.next(shape, "normal", self.loc.dtype) * self.scale

    Laplace.rsample, Normal.rsample = l_rsample, n_rsample
    try:
        yield queue
    finally:
        Laplace.rsample, Normal.rsample = orig_l, orig_n
