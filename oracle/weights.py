"""Deterministic synthetic weights keyed by state-dict key (no pretrained files, no network).

``keyed_state_dict`` fills every tensor from a numpy PCG64 stream seeded with
crc32(key) ^ seed, so the same key always gets the same values regardless of construction
order, torch version or device.  Scales keep activations O(1) through deep stacks:
matrices/filters ~ U(-a, a) with a = sqrt(3 / fan_in) (unit-variance preserving), norm gains
1 + 0.1 U, biases 0.1 U.  TEST INFRASTRUCTURE - see oracle/__init__.py.
"""

from __future__ import annotations

import zlib
from typing import Mapping

import numpy as np
import torch


def keyed_tensor(key: str, shape: tuple[int, ...], seed: int = 0) -> torch.Tensor:
    rng = np.random.Generator(np.random.PCG64(zlib.crc32(key.encode()) ^ seed))
    u = rng.random(size=shape, dtype=np.float32) * 2.0 - 1.0
    is_norm = any(tag in key for tag in ("GroupNorm", "LayerNorm"))
    if key.endswith(".bias"):
        u *= 0.1
    elif is_norm:
        u = 1.0 + 0.1 * u
    elif len(shape) >= 2:
        fan_in = int(np.prod(shape[1:]))
        u *= np.float32(np.sqrt(3.0 / fan_in))
    else:
        u *= 0.1
    return torch.from_numpy(np.ascontiguousarray(u))


def keyed_state_dict(shapes: Mapping[str, tuple[int, ...]], seed: int = 0) -> dict[str, torch.Tensor]:
    from concurrent.futures import ThreadPoolExecutor

    items = list(shapes.items())
    with ThreadPoolExecutor(max_workers=8) as pool:  # numpy releases the GIL while generating
        tensors = list(pool.map(lambda kv: keyed_tensor(kv[0], tuple(kv[1]), seed), items))
    return {k: t for (k, _), t in zip(items, tensors)}
