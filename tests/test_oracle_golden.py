"""Pin the oracle against the fixtures produced by executing the reference (oracle/gen_golden.py)."""
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, golden_files


def _cfg_from(oracle, z, soft_scope="local"):
    a = {k[4:]: z[k].item() for k in z.files if k.startswith("arg_")}
    return oracle.OracleConfig(
        lambda_original=float(a.get("lambda_original", 1.0)),
        lambda_soft=float(a.get("lambda_soft", 0.0)),
        soft_mode=str(a.get("soft_mode", "none")),
        teacher_temp=float(a.get("teacher_temp", 0.15)),
        soft_dino_to_text=bool(a.get("soft_dino_to_text", False)),
        text_lambda=float(a.get("text_lambda", 0.2)),
        text_student_temp=float(a.get("text_student_temp", 0.05)),
        world_size=int(z["world"]),
        local_loss=bool(z["local_loss"]),
        gather_with_grad=bool(z["gather_with_grad"]),
        soft_scope=soft_scope,  # the reference computes the soft terms on the local block
        residual_projection=bool(a.get("residual_projection", False)),
        residual_alpha=(None if a.get("residual_alpha") is None else float(a["residual_alpha"])),
        lambda_weighted=float(a.get("lambda_weighted", 0.0)),
        rho=float(a.get("rho", 0.1)),
        c_clip=float(a.get("c_clip", 1.0)),
        weight_text_symmetry=bool(a.get("weight_text_symmetry", False)),
    ), a


def _head(z):
    keys = [k for k in z.files if k.startswith("head_")]
    if not keys:
        return None
    return {k[5:]: torch.from_numpy(z[k]) for k in keys}


@pytest.mark.parametrize("fname", golden_files())
@pytest.mark.parametrize("tag,dtype,rtol", [("f64", torch.float64, 1e-9), ("f32", torch.float32, 2e-4)])
def test_oracle_matches_reference_fixture(oracle, fname, tag, dtype, rtol):
    z = np.load(os.path.join(GOLDEN_DIR, fname))
    cfg, a = _cfg_from(oracle, z)
    img, txt, dino = (torch.from_numpy(z[k]) for k in ("image", "text", "dino"))
    use_proj = bool(a.get("use_projection", True))
    head = _head(z) if use_proj else None
    res = oracle.loss_and_grads(img, txt, float(z["scale"]), dino, cfg, proj_params=head,
                                projection_type=str(a.get("projection_type", "mlp")), dtype=dtype)
    for r, got in enumerate(res["ranks"]):
        for k in ("total_loss", "classic_loss", "soft_loss", "weighted_loss"):
            ref = float(z[f"{tag}_r{r}_{k}"])
            assert got[k] == pytest.approx(ref, rel=rtol, abs=rtol), (k, r)
        for k in ("d_image", "d_text"):
            ref = z[f"{tag}_r{r}_{k}"]
            err = np.abs(got[k].double().numpy() - ref).max() / max(np.abs(ref).max(), 1e-30)
            assert err < rtol * 10, (k, r, err)
        ref = float(z[f"{tag}_r{r}_d_logit_scale"])
        assert got["d_logit_scale"] == pytest.approx(ref, rel=rtol * 10, abs=rtol * 1e-2), r
        if head is not None and cfg.soft_enabled:
            for hk, g in got["d_proj"].items():
                key = f"{tag}_r{r}_dhead_{hk}"
                if key not in z.files or g is None:
                    continue
                ref = z[key]
                err = np.abs(g.double().numpy() - ref).max() / max(np.abs(ref).max(), 1e-30)
                assert err < rtol * 10, (hk, r, err)


def test_local_loss_false_crashes_like_reference(oracle):
    """Reference at W>1 with local_loss=False raises ValueError from cross_entropy (SURVEY.md probe table)."""
    cfg = oracle.OracleConfig(world_size=2, local_loss=False)
    x = torch.nn.functional.normalize(torch.randn(8, 16), dim=-1)
    with pytest.raises(ValueError, match="Expected input batch_size"):
        oracle.rank_loss(x, x, torch.tensor(10.0), None, None, cfg, rank=0)


def test_student_tau_bands(oracle):
    """compute_student_tau (loss.py:166-175): 0.02 on (10,50], 1/s on (50,100], clamp at 0.008/0.02."""
    f = lambda v: float(oracle.compute_student_tau(torch.tensor(v)))
    assert f(14.2857) == pytest.approx(0.02)
    assert f(80.0) == pytest.approx(1 / 80.0)
    assert f(100.0) == pytest.approx(0.01)
    assert f(1000.0) == pytest.approx(0.01)
    assert f(4.0) == pytest.approx(1 / np.exp(4.0), rel=1e-6)  # treated as a raw ln-scale
    assert f(1.0) == pytest.approx(0.02)
