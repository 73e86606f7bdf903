"""timm.models.layers.{DropPath, trunc_normal_, to_2tuple} with timm 0.4.12 semantics."""
import collections.abc
from itertools import repeat

import torch
import torch.nn as nn

trunc_normal_ = nn.init.trunc_normal_


def to_2tuple(x):
    return tuple(x) if isinstance(x, collections.abc.Iterable) and not isinstance(x, str) else tuple(repeat(x, 2))


def drop_path(x, drop_prob: float = 0.0, training: bool = False):
    """Stochastic depth per sample (timm/models/layers/drop.py)."""
    if drop_prob == 0.0 or not training:
        return x
    keep_prob = 1 - drop_prob
    shape = (x.shape[0],) + (1,) * (x.ndim - 1)
    random_tensor = keep_prob + torch.rand(shape, dtype=x.dtype, device=x.device)
    random_tensor.floor_()
    return x.div(keep_prob) * random_tensor


class DropPath(nn.Module):
    def __init__(self, drop_prob=None):
        super().__init__()
        self.drop_prob = drop_prob

    def forward(self, x):
        return drop_path(x, self.drop_prob, self.training)
