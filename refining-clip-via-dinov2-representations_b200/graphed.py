"""CUDA-graph capture of the loss forward + backward for the launch-bound regime (BASELINE config 2: global batch
4096 on one GPU spends 0.85 ms per step of which 0.35 ms are kernels; SURVEY section 7 "small-B regime").

The C ABI is capture-safe by construction (every call is asynchronous on the caller's stream, no host
synchronisation, tensor maps travel as kernel parameters), so the whole module - projection head, pack, NCCL
all-gather, tile kernels, finalize - records into one forward and one backward graph with
``torch.cuda.make_graphed_callables``.  Shapes, dtypes and the `args` knobs are frozen at capture time.

    step = dinosoft_b200.make_graphed(loss, loss_args, image, text, logit_scale, dino)   # sample tensors
    total, classic, soft = step(image, text, logit_scale, dino)                          # every iteration
    total.backward()
"""
from __future__ import annotations

import torch
import torch.nn as nn


class _TensorOnly(nn.Module):
    """Tensor-in / tensor-out view of the loss module (what graph capture needs)."""

    def __init__(self, loss: nn.Module, args):
        super().__init__()
        self.loss = loss
        self.args = args

    def forward(self, image_features, text_features, logit_scale, dino_features):
        out = self.loss(image_features, text_features, logit_scale, dino_features, self.args, output_dict=True)
        return out["total_loss"], out["classic_loss"], out["soft_loss"]


def make_graphed(loss: nn.Module, args, image_features: torch.Tensor, text_features: torch.Tensor,
                 logit_scale: torch.Tensor, dino_features: torch.Tensor, autocast_dtype=None, num_warmup_iters: int = 3):
    """Returns ``step(image, text, logit_scale, dino) -> (total_loss, classic_loss, soft_loss)`` whose forward and
    backward replay CUDA graphs.  The sample tensors fix shapes / dtypes / requires_grad; `autocast_dtype`
    (e.g. torch.bfloat16) wraps the captured forward in torch.autocast the way train.py:285 calls the loss."""
    if image_features.device.type != "cuda":
        raise RuntimeError("make_graphed needs CUDA tensors")
    if getattr(loss, "image_to_dino_proj", None) is None and getattr(args, "use_projection", True) \
            and dino_features is not None:
        # the head is created lazily from the global RNG at the first forward (loss.py:223-238): do it before capture
        loss.init_proj(image_features.size(-1), dino_features.size(-1), image_features.device,
                       getattr(args, "projection_type", "mlp"), layernorm=getattr(args, "use_layernorm", False))
    mod = _TensorOnly(loss, args)
    if autocast_dtype is not None:
        inner = mod.forward

        def fwd(*a):
            with torch.autocast("cuda", dtype=autocast_dtype, cache_enabled=False):
                return inner(*a)

        mod.forward = fwd
    sample = (image_features, text_features, logit_scale, dino_features)
    return torch.cuda.make_graphed_callables(mod, sample, num_warmup_iters=num_warmup_iters)
