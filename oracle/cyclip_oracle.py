"""CPU restatement of the reference's CyCLIP loss.  TEST INFRASTRUCTURE - never imported by the product package.

Follows `CyCLIPLoss.forward` (src/open_clip/loss.py:867-905) literally - the B x B cosine matrices and the two mean
squared differences - in plain tensor algebra with autograd for the gradients, world size 1.  Pinned by
tests/golden/cyclip_b96.npz, the outputs of the reference class itself (oracle/gen_golden_cyclip.py)."""
from __future__ import annotations

import torch


def loss_and_grads(image, text, scale, lambda_inmodal=0.25, lambda_crossmodal=0.25, dtype=torch.float64):
    im = image.detach().to(dtype).requires_grad_(True)
    tx = text.detach().to(dtype).requires_grad_(True)
    sc = torch.tensor(float(scale), dtype=dtype, requires_grad=True)
    # contrastive term (loss.py:870-875): mean row CE of both directions against the diagonal
    logits = sc * im @ tx.t()
    lse_r = torch.logsumexp(logits, dim=1)
    lse_c = torch.logsumexp(logits, dim=0)
    diag = logits.diagonal()
    clip = 0.5 * ((lse_r - diag).mean() + (lse_c - diag).mean())
    # consistency terms on L2-normalised features (loss.py:877-892)
    I = im / im.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    T = tx / tx.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    S_ii, S_tt, S_it = I @ I.t(), T @ T.t(), I @ T.t()
    cross = ((S_it - S_it.t()) ** 2).mean()
    inmod = ((S_ii - S_tt) ** 2).mean()
    total = clip + lambda_inmodal * inmod + lambda_crossmodal * cross
    total.backward()
    return {"total_loss": float(total), "clip_loss": float(clip), "inmodal_cyclic": float(inmod),
            "crossmodal_cyclic": float(cross), "d_image": im.grad, "d_text": tx.grad, "d_logit_scale": float(sc.grad)}
