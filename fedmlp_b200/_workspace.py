"""Per-device scratch buffers handed to the C ABI (the library never allocates)."""
from __future__ import annotations

import torch

_cache: dict = {}


def workspace(name: str, nbytes: int, device) -> torch.Tensor:
    """A uint8 CUDA buffer of at least `nbytes` (256-byte aligned by the caching allocator),
    reused across calls on the same (device, stream)."""
    device = torch.device(device)
    key = (name, device.index, torch.cuda.current_stream(device).cuda_stream)
    buf = _cache.get(key)
    nbytes = max(int(nbytes), 256)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        _cache[key] = buf
    return buf
