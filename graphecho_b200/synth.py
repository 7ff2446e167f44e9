"""Deterministic synthetic inputs of the shapes BASELINE.json names (no datasets are reachable).

Images: uniform [0,1) like the reference datasets after /255 (datasets/camus.py:105, echo.py:189,
cardiac_uda.py:155).  Clips use the trainers' [b,1,H,W,t] layout.  Masks: one-hot [B,nc,H,W]
float, channel 0 = background, foreground = discs so that every class is present at every
batch index.  All generation happens on the CPU generator (bit-identical on every host) and is
moved to the device by the caller.
"""
from __future__ import annotations

import torch


def images(batch: int, hw: int, seed: int = 0) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, 1, hw, hw, generator=g)


def clips(n_clips: int, hw: int, frames: int, seed: int = 0) -> torch.Tensor:
    """[b,1,H,W,t] as the video datasets return (train_camus_echo.py:246-248)."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(n_clips, 1, hw, hw, frames, generator=g)


def flatten_clips(x: torch.Tensor) -> torch.Tensor:
    """[b,c,h,w,t] -> [b*t,c,h,w], the trainers' permute(0,4,1,2,3).reshape (train_*.py:248/274)."""
    b, c, h, w, t = x.shape
    return x.permute(0, 4, 1, 2, 3).reshape(b * t, c, h, w)


def disc_masks(batch: int, num_classes: int, hw: int, shift: int = 0) -> torch.Tensor:
    s = hw / 256.0
    yy, xx = torch.meshgrid(torch.arange(hw, dtype=torch.float32), torch.arange(hw, dtype=torch.float32), indexing="ij")
    out = torch.zeros(batch, num_classes, hw, hw)
    for b in range(batch):
        k = b % 8
        lv = ((xx - (100 + 5 * k + shift) * s) ** 2 + (yy - 90 * s) ** 2) <= (40 * s) ** 2
        rv = (((xx - 150 * s) ** 2 + (yy - (170 - 3 * k + shift) * s) ** 2) <= (30 * s) ** 2) & ~lv
        la = (((xx - 60 * s) ** 2 + (yy - 180 * s) ** 2) <= (25 * s) ** 2) & ~lv & ~rv
        fg = [lv, rv, la][: max(num_classes - 1, 0)]
        for c, m in enumerate(fg, start=1):
            out[b, c] = m.float()
        if num_classes >= 1:
            union = torch.zeros_like(lv)
            for m in fg:
                union |= m
            out[b, 0] = (~union).float()
    return out


def pyramid(batch: int, hw: int, seed: int = 0, channels: int = 256) -> list[torch.Tensor]:
    """Random stand-in for FPN's [p2,p3,p4,p5] at input size hw (strides 4/8/16/32, ceil)."""
    g = torch.Generator().manual_seed(seed)
    sizes, s = [], hw
    s = -(-s // 2)          # conv stride 2
    s = -(-s // 2)          # max-pool -> stride 4
    for _ in range(4):
        sizes.append(s)
        s = -(-s // 2)
    return [torch.randn(batch, channels, k, k, generator=g) for k in sizes]


def clip_pyramid(n_clips: int, frames: int, hw: int, seed: int = 0, channels: int = 256) -> list[torch.Tensor]:
    """4 x [b,t,C,s,s] as the trainers feed TGCN (train_*.py:300-304)."""
    return [p.reshape(n_clips, frames, channels, p.shape[-1], p.shape[-1])
            for p in pyramid(n_clips * frames, hw, seed, channels)]
