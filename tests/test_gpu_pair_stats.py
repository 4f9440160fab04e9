"""dsoft_pair_stats through the package wrapper against the reference fixture and the CPU oracle.
Counts are integers: exact, except for a pair whose cosine sits within the split-bf16 accuracy (~1e-5) of a
threshold - the comparison allows exactly those."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, ROOT

pytestmark = pytest.mark.gpu


def load_oracle():
    spec = importlib.util.spec_from_file_location("pair_stats_oracle", os.path.join(ROOT, "oracle", "pair_stats_oracle.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def near_threshold(clip, dino, thr, eps=3e-5):
    """Per threshold: how many pairs i < j could flip a count under a cosine perturbation of eps."""
    cs = clip.double() @ clip.double().t()
    ds = dino.double() @ dino.double().t()
    up = torch.triu(torch.ones_like(cs, dtype=torch.bool), 1)
    out = []
    for cmin, dmax in thr:
        a = int((up & ((cs - cmin).abs() < eps)).sum())
        b = a + int((up & (cs >= cmin - eps) & ((ds - dmax).abs() < eps)).sum())
        out.append((a, b))
    return out


def check(pkg, clip, dino, thr, topk=200):
    from dinosoft_b200.pair_stats import pair_stats

    want = load_oracle().pair_stats(clip, dino, thr, topk=topk)
    got = pair_stats(clip.cuda(), dino.cuda(), thr, topk=topk)
    slack = near_threshold(clip, dino, thr)
    assert got["total_pairs"] == want["total_pairs"] and list(got["results"]) == list(want["results"])
    for (key, w), (sa, sb) in zip(want["results"].items(), slack):
        g = got["results"][key]
        print(f"[pairs] {key}: clip_high {g['clip_high_count']} (ref {w['clip_high_count']}, slack {sa}), "
              f"blind {g['count']} (ref {w['count']}, slack {sb})")
        assert abs(g["clip_high_count"] - w["clip_high_count"]) <= sa
        assert abs(g["count"] - w["count"]) <= sb
        assert g["percent"] == pytest.approx(100.0 * g["count"] / want["total_pairs"])
        assert g["relative_percent"] == pytest.approx(100.0 * g["count"] / (g["clip_high_count"] or 1))
    assert len(got["top_pairs"]) == len(want["top_pairs"])
    if want["top_pairs"]:
        np.testing.assert_allclose([p["gap"] for p in got["top_pairs"]], [p["gap"] for p in want["top_pairs"]],
                                   atol=5e-5)
        ws = {(p["i"], p["j"]) for p in want["top_pairs"]}
        gs = {(p["i"], p["j"]) for p in got["top_pairs"]}
        assert len(ws & gs) >= len(ws) - 3  # pairs at the cut may swap within the cosine accuracy
        for p in got["top_pairs"][:10]:
            assert p["i"] < p["j"] and p["gap"] == pytest.approx(p["clip_sim"] - p["dino_sim"], abs=1e-6)
    return got


def test_reference_fixture(pkg):
    z = np.load(os.path.join(GOLDEN_DIR, "pair_stats_n300.npz"))
    thr = [tuple(float(x) for x in row) for row in z["thresholds"]]
    got = check(pkg, torch.from_numpy(z["clip"]), torch.from_numpy(z["dino"]), thr)
    for n, k in enumerate(z["keys"]):
        assert abs(got["results"][str(k)]["count"] - int(z["count"][n])) <= 1


@pytest.mark.parametrize("n,d,dd", [(2, 8, 8), (129, 40, 24), (1000, 512, 768), (4096, 512, 768)])
def test_sizes_against_oracle(pkg, n, d, dd):
    g = torch.Generator().manual_seed(n)
    k = max(n // 12, 1)
    cid = torch.randint(0, k, (n,), generator=g)
    clip = torch.nn.functional.normalize(torch.randn(k, d, generator=g)[cid] + 0.4 * torch.randn(n, d, generator=g), dim=-1)
    cid2 = torch.where(torch.rand(n, generator=g) < 0.3, torch.randint(0, k, (n,), generator=g), cid)
    dino = torch.nn.functional.normalize(torch.randn(k, dd, generator=g)[cid2] + 0.7 * torch.randn(n, dd, generator=g), dim=-1)
    check(pkg, clip, dino, [(0.9, 0.6), (0.8, 0.5), (0.5, 0.1)], topk=min(200, n * (n - 1) // 2))


def test_bf16_inputs_and_errors(pkg):
    from dinosoft_b200.pair_stats import pair_stats

    g = torch.Generator().manual_seed(3)
    clip = torch.nn.functional.normalize(torch.randn(600, 64, generator=g), dim=-1).to(torch.bfloat16)
    dino = torch.nn.functional.normalize(torch.randn(600, 96, generator=g), dim=-1).to(torch.bfloat16)
    check(pkg, clip.float(), dino.float(), [(0.2, 0.0)])          # oracle on the same (bf16-representable) values
    got = pair_stats(clip.cuda(), dino.cuda(), [(0.2, 0.0)])      # single-product path
    want = load_oracle().pair_stats(clip.float(), dino.float(), [(0.2, 0.0)])
    assert abs(got["results"]["clip≥0.2_dino≤0.0"]["count"] - want["results"]["clip≥0.2_dino≤0.0"]["count"]) <= 2
    with pytest.raises(RuntimeError):
        pair_stats(clip, dino, [(0.2, 0.0)])                       # CPU tensors: no fallback
    with pytest.raises(ValueError):
        pair_stats(clip.cuda(), dino.cuda()[:10], [(0.2, 0.0)])
