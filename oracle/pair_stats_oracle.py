"""CPU restatement of the reference's CLIP-blind pair statistics.  TEST INFRASTRUCTURE - never imported by the
product package.

Follows `_pair_stats` (src/open_clip_train/helpers.py:221-285): cosine matrices of L2-normalised rows
(helpers.py:243-244), the strict upper triangle (247-249), per threshold the number of pairs with CLIP similarity
>= cmin and of those the pairs with DINO similarity <= dmax (255-271), and the `topk` pairs with the largest
CLIP - DINO gap (273-283).  Written with explicit row blocks instead of index tensors so that N = 32768 fits in
memory; pinned by tests/golden/pair_stats_*.npz, which hold the outputs of the reference function itself
(oracle/gen_golden_pairs.py)."""
from __future__ import annotations

import torch


def pair_stats(clip_Z: torch.Tensor, dino_Z: torch.Tensor, thresholds, topk: int = 200, block: int = 2048,
               dtype=torch.float64) -> dict:
    z = clip_Z.to(dtype)
    d = dino_Z.to(dtype)
    n = z.shape[0]
    total = n * (n - 1) // 2
    hi = [0] * len(thresholds)
    blind = [0] * len(thresholds)
    best = None  # (gap, i, j, cs, ds) of the running top-k
    for r0 in range(0, n, block):
        r1 = min(r0 + block, n)
        cs = z[r0:r1] @ z.t()
        ds = d[r0:r1] @ d.t()
        rows = torch.arange(r0, r1).unsqueeze(1)
        cols = torch.arange(n).unsqueeze(0)
        upper = cols > rows
        for k, (cmin, dmax) in enumerate(thresholds):
            m = upper & (cs >= cmin)
            hi[k] += int(m.sum())
            blind[k] += int((m & (ds <= dmax)).sum())
        if topk > 0 and total > 0:
            gap = torch.where(upper, cs - ds, torch.full_like(cs, -1e9)).reshape(-1)
            kk = min(topk, gap.numel())
            val, idx = torch.topk(gap, kk)
            i = idx // n + r0
            j = idx % n
            cand = torch.stack([val, i.to(dtype), j.to(dtype), cs.reshape(-1)[idx], ds.reshape(-1)[idx]], 1)
            cand = cand[val > -1e8]
            best = cand if best is None else torch.cat([best, cand], 0)
            # largest gap first, ties in row-major pair order
            order = torch.argsort(best[:, 1] * n + best[:, 2])
            best = best[order]
            best = best[torch.argsort(best[:, 0], descending=True, stable=True)][:min(topk, total)]
    out = {"total_pairs": total, "results": {}, "thresholds": list(thresholds)}
    for k, (cmin, dmax) in enumerate(thresholds):
        out["results"][f"clip≥{cmin}_dino≤{dmax}"] = {
            "count": blind[k], "percent": 100.0 * blind[k] / (total or 1), "clip_high_count": hi[k],
            "relative_percent": 100.0 * blind[k] / (hi[k] or 1)}
    out["top_pairs"] = [] if best is None else [
        {"i": int(r[1]), "j": int(r[2]), "clip_sim": float(r[3]), "dino_sim": float(r[4]), "gap": float(r[0])}
        for r in best]
    return out
