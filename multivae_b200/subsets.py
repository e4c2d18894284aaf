"""Integer / index logic of the aggregation path (must be bit-exact with the reference)."""
from itertools import chain, combinations

import torch


def all_subsets(mod_names):
    """Insertion-ordered {key: sorted modality list} including the empty subset first
    (reference: MoPoE.all_subsets + set_subsets, models/mopoe/mopoe_model.py:76-106)."""
    xs = list(mod_names)
    out = {}
    for names in chain.from_iterable(combinations(xs, n) for n in range(len(xs) + 1)):
        out["_".join(sorted(names))] = sorted(names)
    return out


def subset_bitmask(mods, order):
    return sum(1 << order.index(m) for m in mods)


def deterministic_selection(num_samples, num_subsets):
    """Sample -> subset index, contiguous balanced slices: i_end = i_start + int(floor(B * float32(1/S))),
    the last subset takes the remainder (reference: deterministic_mixture_component_selection,
    mopoe_model.py:435-465, with w = float32(1/S) built at :337)."""
    w = (1 / float(num_subsets)) * torch.ones(num_subsets)
    out = torch.empty(num_samples, dtype=torch.int32)
    start = 0
    for k in range(num_subsets):
        end = num_samples if k == num_subsets - 1 else start + int(torch.floor(num_samples * w[k]))
        out[start:end] = k
        start = end
    return out


def mvae_random_subsets(mod_names):
    """All subsets of size 2..M-1 in itertools order (reference: MVAE._set_subsets, mvae_model.py:48-51)."""
    xs = list(mod_names)
    out = []
    for i in range(2, len(xs)):
        out += combinations(xs, r=i)
    return out
