"""Parameter dictionaries for the oracle, shaped after the reference's own state_dicts
(tests/golden/state_contract.json, dumped from the reference by oracle/make_golden.py) and filled
by the name-keyed deterministic fill.  Test infrastructure."""
from __future__ import annotations

import json
from pathlib import Path

import torch

from .detfill import fill_state

CONTRACT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "state_contract.json"


def contract() -> dict:
    return json.loads(CONTRACT.read_text())


def make_params(module: str, fill_prefix: str = "", scale: float = 1.0, key_prefix: str = "",
                requires_grad: bool = False, overrides: dict | None = None) -> dict:
    """Tensors for every state_dict entry of `module`, keyed `key_prefix + name`, filled as
    detfill would fill the reference module with prefix `fill_prefix`."""
    spec = dict(contract()[module])
    spec.update(overrides or {})
    out = {}
    for name, shape in spec.items():
        dtype = torch.long if name.endswith("num_batches_tracked") else torch.float32
        out[name] = torch.zeros(shape, dtype=dtype)
    fill_state(out, scale, fill_prefix)
    res = {}
    for name, t in out.items():
        if requires_grad and t.is_floating_point() and not any(
                name.endswith(s) for s in ("running_mean", "running_var", "sr_seed", "tg_seed")):
            t.requires_grad_(True)
        res[key_prefix + name] = t
    return res
