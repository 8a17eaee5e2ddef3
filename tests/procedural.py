"""Procedurally generated (seeded, per-key) parameter tensors shared by tools/gen_golden.py (which loads them into the
REFERENCE's modules in the build container) and by the tests (which load the same tensors into this repository's modules),
so that multi-gigabyte state dicts never have to be committed: only inputs and reference outputs are stored as fixtures."""
from __future__ import annotations

import zlib
from typing import Dict, Tuple

import torch


def procedural_tensor(key: str, shape: Tuple[int, ...], seed: int = 0) -> torch.Tensor:
    """Deterministic fp32 tensor for a state-dict entry: matrices / conv kernels ~ N(0, 1/fan_in) (activations keep O(1)
    variance through the net), 1-D `weight` (norm gains) ~ 1 + 0.1 N(0,1), 1-D `bias` ~ 0.1 N(0,1)."""
    g = torch.Generator().manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    if len(shape) >= 2:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        return torch.randn(shape, generator=g) * fan_in ** -0.5
    if key.endswith("weight"):
        return 1.0 + 0.1 * torch.randn(shape, generator=g)
    return 0.1 * torch.randn(shape, generator=g)


def procedural_state_dict(shapes: Dict[str, Tuple[int, ...]], seed: int = 0) -> Dict[str, torch.Tensor]:
    return {k: procedural_tensor(k, tuple(s), seed) for k, s in shapes.items()}


def load_procedural(module: torch.nn.Module, seed: int = 0) -> None:
    """Fill `module` (possibly built on the meta device) with the procedural tensors of its own state-dict keys."""
    shapes = {k: tuple(v.shape) for k, v in module.state_dict().items()}
    module.load_state_dict(procedural_state_dict(shapes, seed), strict=True, assign=True)
