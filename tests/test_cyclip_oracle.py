"""The CyCLIP oracle against the fixture produced by the reference's own class (loss.py:813-905), and the D x D
moment-matrix identities the product uses against the literal B x B evaluation (CPU, fp64)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, ROOT


def load(name):
    spec = importlib.util.spec_from_file_location(name, os.path.join(ROOT, "oracle", name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("tag,dtype,tol", [("f64", torch.float64, 1e-10), ("f32", torch.float32, 2e-4)])
def test_oracle_matches_reference_fixture(tag, dtype, tol):
    z = np.load(os.path.join(GOLDEN_DIR, "cyclip_b96.npz"))
    o = load("cyclip_oracle").loss_and_grads(torch.from_numpy(z["image"]), torch.from_numpy(z["text"]), float(z["scale"]),
                                             float(z["lambda_inmodal"]), float(z["lambda_crossmodal"]), dtype=dtype)
    for k in ("total_loss", "clip_loss", "inmodal_cyclic", "crossmodal_cyclic"):
        assert o[k] == pytest.approx(float(z[f"{tag}_{k}"]), rel=tol), k
    for k in ("d_image", "d_text"):
        ref = torch.from_numpy(z[f"{tag}_{k}"])
        assert (o[k].double() - ref).abs().max() <= tol * 10 * ref.abs().max(), k
    assert o["d_logit_scale"] == pytest.approx(float(z[f"{tag}_d_logit_scale"]), rel=tol * 10, abs=1e-12)


def test_moment_matrix_identities():
    """sum (S_ii - S_tt)^2 and sum (S_it - S_ti)^2 from A = I^T I, C = T^T T, M = I^T T, with gradients."""
    import dinosoft_b200 as pkg
    from dinosoft_b200.cyclip import _CyclicFn

    g = torch.Generator().manual_seed(2)
    x = torch.randn(200, 48, generator=g, dtype=torch.float64).requires_grad_(True)
    y = (x.detach() @ torch.randn(48, 48, generator=g, dtype=torch.float64) * 0.2
         + torch.randn(200, 48, generator=g, dtype=torch.float64)).requires_grad_(True)
    I = x / x.norm(dim=-1, keepdim=True)
    T = y / y.norm(dim=-1, keepdim=True)
    inm = ((I @ I.t() - T @ T.t()) ** 2).mean()
    crs = ((I @ T.t() - T @ I.t()) ** 2).mean()
    (0.3 * inm + 0.7 * crs).backward()
    x2 = x.detach().float().requires_grad_(True)
    y2 = y.detach().float().requires_grad_(True)
    a, b = _CyclicFn.apply(x2, y2)
    (0.3 * a + 0.7 * b).backward()
    assert float(a) == pytest.approx(float(inm), rel=1e-5) and float(b) == pytest.approx(float(crs), rel=1e-5)
    for got, want in ((x2.grad, x.grad), (y2.grad, y.grad)):
        assert (got.double() - want).abs().max() <= 1e-4 * want.abs().max()
