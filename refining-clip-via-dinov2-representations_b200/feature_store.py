"""Device-resident DINO feature store.

The reference keeps the precomputed DINOv2 CLS features as a pinned CPU tensor `[N, Dd]` fp32 and does, every
step (src/open_clip_train/main.py:693-741, src/open_clip_train/train.py:250-280):

    idx_cpu = indices.to("cpu"); mi, ma = idx_cpu.min().item(), idx_cpu.max().item()   # host sync + range check
    dino_features = precomputed[indices].to(device, non_blocking=True)                 # CPU gather + H2D copy

Here the table lives in HBM as bf16 (the loss rounds DINO features to bf16 anyway; Flickr30k-scale: 31 k x 768
x 2 B = 48 MB, LAION-scale millions of rows fit the B200's 180 GB) and the gather is one kernel of libdsoft.so
(`dsoft_gather_rows`) that checks the index range ON THE DEVICE and can write straight into the DINO columns of
the loss's packed operand buffer, so the step upstream of the loss needs no CPU work, no H2D copy and no host
synchronisation.

Drop-in use with the unmodified train loop: `args._precomputed_dino = DinoFeatureStore(table, device)`.
`store[indices]` returns a lazy `DinoRows` handle whose `.to(device, non_blocking=True)` is a no-op, and the loss
module resolves it inside its pack step.  `store.shape` serves train.py's own range check; `store.check()` raises
the reference's ValueError (train.py:261-268) from the device-side record whenever the caller wants it (it is
the only call here that synchronises).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _cabi

_DT = {torch.float32: _cabi.DT_F32, torch.bfloat16: _cabi.DT_BF16, torch.float16: _cabi.DT_F16}
_I64_MAX, _I64_MIN = (1 << 63) - 1, -(1 << 63)


class DinoRows:
    """`store[indices]`: the rows are gathered where they are consumed (the loss's pack step)."""

    def __init__(self, store: "DinoFeatureStore", indices: torch.Tensor):
        self.store = store
        self.indices = indices.to(device=store.device, dtype=torch.int64, non_blocking=True).reshape(-1)

    # what train.py / the loss module ask of `dino_features`
    @property
    def shape(self):
        return torch.Size((self.indices.numel(), self.store.shape[1]))

    def size(self, dim: Optional[int] = None):
        return self.shape if dim is None else self.shape[dim]

    @property
    def device(self):
        return self.store.device

    @property
    def dtype(self):
        return self.store.dtype

    def to(self, *args, **kwargs):  # train.py:280 `.to(device, non_blocking=True)`: already there
        return self

    def materialize(self, dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
        return self.store.lookup(self.indices, dtype=dtype)


class DinoFeatureStore:
    def __init__(self, precomputed: torch.Tensor, device, dtype: torch.dtype = torch.bfloat16):
        device = torch.device(device)
        if device.type != "cuda":
            raise RuntimeError("DinoFeatureStore keeps the table in HBM: it needs a CUDA device (no CPU path)")
        if precomputed.dim() != 2 or precomputed.shape[1] % 8:
            raise ValueError(f"expected a [N, Dd] table with Dd a multiple of 8, got {tuple(precomputed.shape)}")
        if dtype not in _DT:
            raise ValueError(f"unsupported table dtype {dtype}")
        self.table = precomputed.detach().to(device=device, dtype=dtype).contiguous()
        self.table.requires_grad_(False)
        self._lib = _cabi.lib()
        # sticky device-side range record: [bad count, min index, max index, one bad index]
        self.status = torch.tensor([0, _I64_MAX, _I64_MIN, -1], dtype=torch.int64, device=device)

    @property
    def shape(self):
        return self.table.shape

    @property
    def device(self):
        return self.table.device

    @property
    def dtype(self):
        return self.table.dtype

    def __len__(self):
        return self.table.shape[0]

    def __getitem__(self, indices) -> DinoRows:
        return DinoRows(self, torch.as_tensor(indices))

    def _gather(self, indices: torch.Tensor, out_ptr: int, out_dt: int, ld_out: int) -> None:
        if torch.cuda.current_device() != self.device.index:
            torch.cuda.set_device(self.device)
        st = torch.cuda.current_stream(self.device).cuda_stream
        _cabi.check(
            self._lib.dsoft_gather_rows(self.table.data_ptr(), _DT[self.table.dtype], self.table.stride(0),
                                        self.table.shape[0], self.table.shape[1], indices.data_ptr(),
                                        indices.numel(), out_ptr, out_dt, ld_out, self.status.data_ptr(), st),
            "dsoft_gather_rows")

    @torch.no_grad()
    def lookup(self, indices: torch.Tensor, dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
        """`table[indices]` as a new [n, Dd] tensor (bf16 or fp32) without a host sync."""
        idx = indices.to(device=self.device, dtype=torch.int64, non_blocking=True).reshape(-1).contiguous()
        out = torch.empty((idx.numel(), self.table.shape[1]), dtype=dtype, device=self.device)
        self._gather(idx, out.data_ptr(), _DT[dtype], out.stride(0))
        return out

    @torch.no_grad()
    def gather_into_packed(self, rows: DinoRows, gathered: torch.Tensor, row0: int, col0: int) -> None:
        """Write the rows straight into the DINO columns of the loss's packed bf16 buffer (rows row0 .., column
        col0 ..): the gather feeds the tile kernels without an intermediate tensor."""
        idx = rows.indices.contiguous()
        ptr = gathered.data_ptr() + 2 * (row0 * gathered.stride(0) + col0)
        self._gather(idx, ptr, _cabi.DT_BF16, gathered.stride(0))

    def check(self) -> None:
        """Raise the reference's error (train.py:261-268) if any index seen so far was out of range.  Synchronises."""
        bad, mi, ma, eg = (int(v) for v in self.status.tolist())
        if bad:
            n = self.table.shape[0]
            raise ValueError(
                f"[DINO] Out-of-range indices: min={mi}, max={ma}, feats_rows={n}. "
                f"Examples of bad indices: [{eg}] ({bad} in total). "
                "This usually means your dino_index_map does not align with the training CSV order "
                "OR contains placeholder -1 entries.")

    def reset_status(self) -> None:
        self.status.copy_(torch.tensor([0, _I64_MAX, _I64_MIN, -1], dtype=torch.int64), non_blocking=True)


# ---- round-1 helper names, kept for callers
def to_device_table(precomputed: torch.Tensor, device, dtype: Optional[torch.dtype] = None) -> DinoFeatureStore:
    return DinoFeatureStore(precomputed, device, dtype or torch.bfloat16)


@torch.no_grad()
def lookup(table, indices: torch.Tensor) -> torch.Tensor:
    if isinstance(table, DinoFeatureStore):
        return table.lookup(indices)
    return table.index_select(0, indices.to(table.device, non_blocking=True).long())
