"""Device-resident DINO feature table.

The reference keeps the precomputed DINOv2 CLS features as a pinned CPU tensor `[N, Dd]` and does, every step,
`precomputed[indices].to(device, non_blocking=True)` after `.item()` range checks on the indices
(src/open_clip_train/main.py:693-741, src/open_clip_train/train.py:250-280).  A Flickr30k-scale table
(31 k x 768 fp32 = 95 MB; even LAION-scale millions of rows fit the B200's 180 GB) can simply live in HBM:
indexing a CUDA tensor keeps the gather on the device and the later `.to(device)` is a no-op, so train.py needs
no change."""
from __future__ import annotations

import torch


def to_device_table(precomputed: torch.Tensor, device, dtype: torch.dtype | None = None) -> torch.Tensor:
    """Return the table on `device` (optionally down-cast, e.g. to bf16: the loss rounds DINO features to bf16
    anyway).  `table[indices]` then works with CPU or CUDA index tensors."""
    t = precomputed.to(device=device, dtype=dtype or precomputed.dtype, non_blocking=True)
    return t.contiguous()


@torch.no_grad()
def lookup(table: torch.Tensor, indices: torch.Tensor) -> torch.Tensor:
    """`table[indices]` on the table's device, without a host sync; out-of-range indices raise at the next sync
    point (device-side assert) instead of through `.item()` checks."""
    return table.index_select(0, indices.to(table.device, non_blocking=True).long())
