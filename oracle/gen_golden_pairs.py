"""Golden fixture for the CLIP-blind pair statistics, produced by EXECUTING the reference's `_pair_stats`
(src/open_clip_train/helpers.py:221-285).  helpers.py cannot be imported as a module here (its tail imports the
webdataset pipeline), so the function's source lines are read from the reference file and executed in a namespace
that holds only `torch` and the typing names - the reference code itself runs, nothing of it is stored in the repo.

    python oracle/gen_golden_pairs.py        # build container only (/root/reference)
"""
import os
import re
from typing import List, Tuple

import numpy as np
import torch

REF = "/root/reference/src/open_clip_train/helpers.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "pair_stats_n300.npz")
THRESHOLDS = [(0.9, 0.6), (0.8, 0.5), (0.7, 0.5), (0.6, 0.2)]


def reference_pair_stats():
    src = open(REF).read()
    m = re.search(r"^def _pair_stats\(.*?(?=^def )", src, re.S | re.M)
    ns = {"torch": torch, "List": List, "Tuple": Tuple}
    exec(compile(m.group(0), REF, "exec"), ns)
    return ns["_pair_stats"]


def inputs(seed=5, n=300, d=64, dd=96):
    g = torch.Generator().manual_seed(seed)
    k = n // 10
    cid = torch.randint(0, k, (n,), generator=g)
    clip = torch.randn(k, d, generator=g)[cid] + 0.35 * torch.randn(n, d, generator=g)
    # DINO geometry partly disagrees with CLIP's: a second clustering for a third of the rows
    cid2 = torch.where(torch.rand(n, generator=g) < 0.33, torch.randint(0, k, (n,), generator=g), cid)
    dino = torch.randn(k, dd, generator=g)[cid2] + 0.6 * torch.randn(n, dd, generator=g)
    nz = torch.nn.functional.normalize
    return nz(clip, dim=-1), nz(dino, dim=-1)


def main():
    fn = reference_pair_stats()
    cz, dz = inputs()
    ref = fn(cz, dz, THRESHOLDS)
    keys = list(ref["results"])
    np.savez_compressed(
        OUT, clip=cz.numpy(), dino=dz.numpy(), thresholds=np.array(THRESHOLDS, dtype=np.float64),
        total_pairs=np.int64(ref["total_pairs"]), keys=np.array(keys),
        count=np.array([ref["results"][k]["count"] for k in keys], dtype=np.int64),
        clip_high_count=np.array([ref["results"][k]["clip_high_count"] for k in keys], dtype=np.int64),
        percent=np.array([ref["results"][k]["percent"] for k in keys]),
        relative_percent=np.array([ref["results"][k]["relative_percent"] for k in keys]),
        top_i=np.array([p["i"] for p in ref["top_pairs"]], dtype=np.int64),
        top_j=np.array([p["j"] for p in ref["top_pairs"]], dtype=np.int64),
        top_gap=np.array([p["gap"] for p in ref["top_pairs"]]),
        top_clip=np.array([p["clip_sim"] for p in ref["top_pairs"]]),
        top_dino=np.array([p["dino_sim"] for p in ref["top_pairs"]]))
    print("wrote", OUT, {k: (ref["results"][k]["clip_high_count"], ref["results"][k]["count"]) for k in keys})


if __name__ == "__main__":
    main()
