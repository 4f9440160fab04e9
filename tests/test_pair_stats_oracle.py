"""The pair-statistics oracle against the fixture produced by the reference's own `_pair_stats`
(helpers.py:221-285, executed by oracle/gen_golden_pairs.py)."""
import importlib.util
import os

import numpy as np
import torch

from conftest import GOLDEN_DIR, ROOT


def load_oracle():
    spec = importlib.util.spec_from_file_location("pair_stats_oracle", os.path.join(ROOT, "oracle", "pair_stats_oracle.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def golden():
    z = np.load(os.path.join(GOLDEN_DIR, "pair_stats_n300.npz"))
    thr = [tuple(float(x) for x in row) for row in z["thresholds"]]
    return z, thr


def test_oracle_matches_reference_fixture():
    z, thr = golden()
    got = load_oracle().pair_stats(torch.from_numpy(z["clip"]), torch.from_numpy(z["dino"]), thr, dtype=torch.float32)
    assert got["total_pairs"] == int(z["total_pairs"])
    assert list(got["results"]) == [str(k) for k in z["keys"]]
    for n, k in enumerate(z["keys"]):
        r = got["results"][str(k)]
        assert r["count"] == int(z["count"][n]) and r["clip_high_count"] == int(z["clip_high_count"][n])
        assert abs(r["percent"] - float(z["percent"][n])) < 1e-12
        assert abs(r["relative_percent"] - float(z["relative_percent"][n])) < 1e-12
    assert len(got["top_pairs"]) == len(z["top_i"]) == 200
    np.testing.assert_allclose([p["gap"] for p in got["top_pairs"]], z["top_gap"], atol=2e-6)
    same = sum((p["i"], p["j"]) == (int(i), int(j)) for p, i, j in zip(got["top_pairs"], z["top_i"], z["top_j"]))
    assert same >= 198  # fp32 ties may swap neighbours


def test_oracle_blocking_is_irrelevant():
    z, thr = golden()
    o = load_oracle()
    a = o.pair_stats(torch.from_numpy(z["clip"]), torch.from_numpy(z["dino"]), thr, block=64)
    b = o.pair_stats(torch.from_numpy(z["clip"]), torch.from_numpy(z["dino"]), thr, block=4096)
    assert a["results"] == b["results"]
    assert [(p["i"], p["j"]) for p in a["top_pairs"]] == [(p["i"], p["j"]) for p in b["top_pairs"]]


def test_tiny_inputs():
    o = load_oracle()
    one = o.pair_stats(torch.ones(1, 8), torch.ones(1, 8), [(0.5, 0.5)])
    assert one["total_pairs"] == 0 and one["top_pairs"] == [] and one["results"]["clip≥0.5_dino≤0.5"]["count"] == 0
