"""Deterministic, name-keyed parameter fill shared by the golden generator (reference modules),
the oracle and the product modules, so identical weights exist on every side without shipping
100 MB state_dicts.  Test infrastructure."""
from __future__ import annotations

import math
import zlib

import torch


def fill_tensor(name: str, t: torch.Tensor, scale: float = 1.0) -> None:
    """Overwrite `t` in place with values that depend only on (name, shape)."""
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    if not t.is_floating_point():
        return  # num_batches_tracked etc.
    shape = tuple(t.shape)
    leaf = name.rsplit(".", 1)[-1]
    if leaf == "running_var":
        v = torch.rand(shape, generator=g) * 0.5 + 0.75
    elif leaf == "running_mean":
        v = torch.randn(shape, generator=g) * 0.1
    elif t.dim() == 1 and leaf == "weight":          # norm scales
        v = 1.0 + 0.1 * torch.randn(shape, generator=g)
    elif t.dim() == 1:                                 # biases
        v = 0.05 * torch.randn(shape, generator=g)
    elif t.dim() == 0:
        v = torch.randn(shape, generator=g)
    elif leaf in ("sr_seed", "tg_seed", "pos_embed", "queue_source", "queue_target"):
        v = torch.randn(shape, generator=g) * (0.1 if leaf == "pos_embed" else 1.0)
    else:                                              # conv / linear weights: He-style fan-in scaling
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        v = torch.randn(shape, generator=g) * (scale * math.sqrt(2.0 / max(fan_in, 1)))
    with torch.no_grad():
        t.copy_(v.to(t.dtype))


def fill_state(state: dict, scale: float = 1.0, prefix: str = "") -> dict:
    for name, t in state.items():
        fill_tensor(prefix + name, t, scale)
    return state


def fill_module(module: torch.nn.Module, scale: float = 1.0, prefix: str = "") -> torch.nn.Module:
    fill_state(module.state_dict(), scale, prefix)
    return module
