"""CyCLIP loss: drop-in for the reference's ``open_clip.loss.CyCLIPLoss`` (loss.py:813-905).

    L = L_CLIP + lambda_inmodal * mean((S_ii - S_tt)^2) + lambda_crossmodal * mean((S_it - S_it^T)^2)

* L_CLIP (loss.py:870-875) runs on the DINO-Soft path's tcgen05 kernels (classic term only: row + column
  log-sum-exp of the never-materialised logits, two-phase backward).
* The two consistency terms (loss.py:877-892) are NOT evaluated from B x B matrices.  With I, T the L2-normalised
  features and the D x D moment matrices  A = I^T I,  C = T^T T,  M = I^T T  :

      sum (S_ii - S_tt)^2   = |A|_F^2 + |C|_F^2 - 2 |M|_F^2
      sum (S_it - S_it^T)^2 = 2 <A, C> - 2 sum_ab M_ab M_ba
      d/dI = 4 (I A - T M^T) / n^2  (in-modal),   4 (I C - T M) / n^2    (cross-modal)
      d/dT = 4 (T C - I M) / n^2    (in-modal),   4 (T A - I M^T) / n^2  (cross-modal)

  i.e. three [D, n] x [n, D] products and four [n, D] x [D, D] products: O(n D^2) work and O(n D) memory instead
  of four n x n matrices (at n = 32768, D = 512: 0.2 TFLOP instead of 4.4, nothing of size n^2).  These are plain
  library GEMMs (cuBLAS, fp32); the moment matrices are accumulated over row blocks in fp64 because each term is a
  difference of nearly equal sums.
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _cabi
from .loss import _DinoSoftFn, _FnConfig, _default_backend

_BLOCK = 4096  # rows per fp32 partial product; partials are summed in fp64


def _moments(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """x^T y for [n, D] fp32 operands, accumulated over row blocks in fp64."""
    acc = torch.zeros((x.shape[1], y.shape[1]), dtype=torch.float64, device=x.device)
    for r in range(0, x.shape[0], _BLOCK):
        acc += (x[r:r + _BLOCK].t() @ y[r:r + _BLOCK]).double()
    return acc


class _CyclicFn(torch.autograd.Function):
    """(image, text) -> (inmodal, crossmodal) of loss.py:877-892 through the D x D moment matrices."""

    @staticmethod
    def forward(ctx, image, text):
        with torch.autocast(image.device.type, enabled=False):
            x, y = image.float(), text.float()
            nx = x.norm(dim=-1, keepdim=True).clamp_min(1e-12)   # F.normalize eps (loss.py:864)
            ny = y.norm(dim=-1, keepdim=True).clamp_min(1e-12)
            I, T = x / nx, y / ny
            n = I.shape[0]
            A, C, M = _moments(I, I), _moments(T, T), _moments(I, T)
            inmod = ((A * A).sum() + (C * C).sum() - 2.0 * (M * M).sum()) / float(n) ** 2
            cross = (2.0 * (A * C).sum() - 2.0 * (M * M.t()).sum()) / float(n) ** 2
        ctx.save_for_backward(I, T, nx, ny, A.float(), C.float(), M.float())
        ctx.dtypes = (image.dtype, text.dtype)
        return inmod.float(), cross.float()

    @staticmethod
    def backward(ctx, g_in, g_cr):
        I, T, nx, ny, A, C, M = ctx.saved_tensors
        with torch.autocast(I.device.type, enabled=False):
            n2 = float(I.shape[0]) ** 2
            a, c = 4.0 * g_in.float() / n2, 4.0 * g_cr.float() / n2
            # gradient w.r.t. the normalised features: [n, D] x [D, D] products
            gI = I @ (a * A + c * C) - T @ (a * M.t() + c * M)
            gT = T @ (a * C + c * A) - I @ (a * M + c * M.t())
            # through F.normalize: (g - x^ <x^, g>) / |x|
            dI = (gI - I * (I * gI).sum(-1, keepdim=True)) / nx
            dT = (gT - T * (T * gT).sum(-1, keepdim=True)) / ny
        return dI.to(ctx.dtypes[0]), dT.to(ctx.dtypes[1])


class CyCLIPLoss(nn.Module):
    """Same constructor, forward signature and return values as the reference class (loss.py:813-905)."""

    def __init__(self, lambda_inmodal: float = 0.25, lambda_crossmodal: float = 0.25, local_loss: bool = False,
                 gather_with_grad: bool = False, cache_labels: bool = False, rank: int = 0, world_size: int = 1,
                 use_horovod: bool = False, process_group=None):
        super().__init__()
        self.lambda_inmodal = lambda_inmodal
        self.lambda_crossmodal = lambda_crossmodal
        self.local_loss = local_loss
        self.gather_with_grad = gather_with_grad
        self.cache_labels = cache_labels
        self.rank = rank
        self.world_size = world_size
        self.use_horovod = use_horovod
        self.process_group = process_group

    def forward(self, image_features, text_features, logit_scale, output_dict: bool = False):
        if self.use_horovod:
            raise NotImplementedError("Horovod is not supported by the B200 path (NCCL only)")
        if self.world_size > 1 and not self.local_loss:
            # the reference evaluates the full [B, B] problem redundantly on every rank (loss.py:843-859);
            # the B200 path shards by row blocks: ask for local_loss=True
            raise NotImplementedError("CyCLIPLoss (B200 build) at world_size > 1 needs local_loss=True")
        device = image_features.device
        cfg = _FnConfig()
        cfg.backend = _default_backend(device)
        cfg.world, cfg.rank, cfg.group = self.world_size, self.rank, self.process_group
        cfg.flags = _cabi.DSOFT_F_ROW_ONLY if (self.world_size > 1 and not self.gather_with_grad) else 0
        cfg.teacher_temp = cfg.text_temp = 0.0
        cfg.lambdas = (1.0, 0.0, 0.0, 0.0)
        cfg.rho, cfg.c_clip, cfg.head_dp = 0.1, 1.0, 0
        terms, _ = _DinoSoftFn.apply(image_features, text_features, logit_scale, None, None, cfg, None, None, None,
                                     None)
        clip_loss = terms[0]
        # world_size > 1 with local_loss: the consistency terms use the LOCAL features (loss.py:843-860)
        L_inmod, L_cross = _CyclicFn.apply(image_features, text_features)
        total = clip_loss + self.lambda_inmodal * L_inmod + self.lambda_crossmodal * L_cross
        if output_dict:
            return {
                "total_loss": total,
                "clip_loss": clip_loss,
                "inmodal_cyclic": L_inmod,
                "crossmodal_cyclic": L_cross,
                "lambda_inmodal": self.lambda_inmodal,
                "lambda_crossmodal": self.lambda_crossmodal,
            }
        return total
