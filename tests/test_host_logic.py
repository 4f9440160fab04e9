"""Host-side logic of the drop-in module on CPU (argument handling, error behaviour, autograd wiring).
The compute backend is replaced by the oracle-backed TEST DOUBLE in tests/oracle_backend.py."""
import os
import types

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR
from oracle_backend import OracleBackend


def args_from(z):
    return types.SimpleNamespace(**{k[4:]: z[k].item() for k in z.files if k.startswith("arg_")})


def test_signature_matches_reference(pkg):
    import inspect

    cls = pkg.ClipLossWithDINOEnhancements
    ctor = list(inspect.signature(cls.__init__).parameters)
    assert ctor[:7] == ["self", "local_loss", "gather_with_grad", "cache_labels", "rank", "world_size", "use_horovod"]
    fwd = list(inspect.signature(cls.forward).parameters)
    assert fwd == ["self", "image_features", "text_features", "logit_scale", "dino_features", "args", "output_dict"]
    for name in ("init_proj", "get_ground_truth", "get_logits"):
        assert hasattr(cls, name)
    m = cls()
    assert m.image_to_dino_proj is None and m.local_loss is False and m.world_size == 1


def test_cpu_tensors_raise_no_fallback(pkg):
    m = pkg.ClipLossWithDINOEnhancements()
    x = torch.nn.functional.normalize(torch.randn(8, 16), dim=-1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(x, x, torch.tensor(10.0), None, None, output_dict=True)


def test_error_behaviour_mirrors_reference(pkg, oracle):
    x = torch.nn.functional.normalize(torch.randn(8, 16), dim=-1)
    m = pkg.ClipLossWithDINOEnhancements(world_size=2, local_loss=False)
    m._backend = OracleBackend(oracle)
    with pytest.raises(ValueError, match=r"Expected input batch_size \(16\) to match target batch_size \(8\)"):
        m(x, x, torch.tensor(10.0), None, None, output_dict=True)
    m = pkg.ClipLossWithDINOEnhancements()
    m._backend = OracleBackend(oracle)
    with pytest.raises(ValueError, match="Unknown projection_type"):
        m(x, x, torch.tensor(10.0), torch.randn(8, 24), types.SimpleNamespace(projection_type="conv"), output_dict=True)
    # lambda_weighted > 0 at world_size > 1: the reference's [b, b] DINO mask cannot broadcast against its
    # [b, B] logits (loss.py:423-446) -> RuntimeError there, RuntimeError here
    m2 = pkg.ClipLossWithDINOEnhancements(world_size=2, local_loss=True, rank=0)
    m2._backend = OracleBackend(oracle)
    with pytest.raises(RuntimeError, match="must match the size of tensor b"):
        m2(x, x, torch.tensor(10.0), torch.randn(8, 24), types.SimpleNamespace(lambda_weighted=0.5, use_projection=False),
           output_dict=True)
    with pytest.raises(NotImplementedError):
        pkg.ClipLossWithDINOEnhancements(use_horovod=True)(x, x, torch.tensor(10.0))


def test_output_dict_false_returns_none_like_reference(pkg, oracle):
    m = pkg.ClipLossWithDINOEnhancements()
    m._backend = OracleBackend(oracle)
    x = torch.nn.functional.normalize(torch.randn(8, 16), dim=-1)
    assert m(x, x, torch.tensor(10.0)) is None


def test_labels_and_tau_helpers(pkg):
    m = pkg.ClipLossWithDINOEnhancements(local_loss=True, rank=3, world_size=4)
    assert m.get_ground_truth(torch.device("cpu"), 5).tolist() == [15, 16, 17, 18, 19]
    assert float(pkg.compute_student_tau(torch.tensor(14.2857))) == pytest.approx(0.02)
    assert float(pkg.compute_student_tau(torch.tensor(80.0))) == pytest.approx(1 / 80)


@pytest.mark.parametrize("fname", ["w1_noproj_text.npz", "w1_classic_only.npz", "w1_scale_below_10.npz",
                                   "w1_mlp_text_scale100.npz", "w1_linear_notext.npz", "w1_mlp_layernorm.npz",
                                   "w1_residual.npz", "w1_residual_alpha.npz", "w1_weighted.npz",
                                   "w1_weighted_sym_all_terms.npz"])
def test_module_wiring_reproduces_reference_fixture(pkg, oracle, fname):
    """Module (host logic) + oracle test double == the reference's own numbers at world_size 1.
    Checks knob decoding, projection-head handling, loss composition and the autograd plumbing."""
    z = np.load(os.path.join(GOLDEN_DIR, fname))
    a = args_from(z)
    img, txt, dino = (torch.from_numpy(z[k]).double() for k in ("image", "text", "dino"))
    m = pkg.ClipLossWithDINOEnhancements()
    m._backend = OracleBackend(oracle)
    use_proj = getattr(a, "use_projection", True)
    if use_proj:
        m.init_proj(img.shape[1], dino.shape[1], "cpu", getattr(a, "projection_type", "mlp"),
                    layernorm=getattr(a, "use_layernorm", False))
        m.image_to_dino_proj = m.image_to_dino_proj.double()
        sd = m.image_to_dino_proj.state_dict()
        mapping = ({"weight": "w0", "bias": "b0"} if getattr(a, "projection_type", "mlp") == "linear" else
                   {"0.weight": "w0", "0.bias": "b0", "2.weight": "w1", "2.bias": "b1", "3.weight": "ln_w", "3.bias": "ln_b"})
        m.image_to_dino_proj.load_state_dict({k: torch.from_numpy(z["head_" + mapping[k]]) for k in sd})
    im = img.clone().requires_grad_(True)
    tx = txt.clone().requires_grad_(True)
    sc = torch.tensor(float(z["scale"]), dtype=torch.float64, requires_grad=True)
    out = m(im, tx, sc, dino, a, output_dict=True)
    assert set(out) == {"total_loss", "classic_loss", "soft_loss", "weighted_loss", "dbg"}
    out["total_loss"].backward()
    # the packed operands are bf16: exact for image/text/dino (bf16-representable fixtures); the head output
    # is rounded, which moves the soft term by O(1e-3) relative at most
    tol = 5e-3 if (use_proj and float(getattr(a, "lambda_soft", 0)) > 0) else 1e-6
    weighted = float(getattr(a, "lambda_weighted", 0.0)) > 0
    if weighted:  # the weighted branch computes in fp32 tensor ops (loss.py:416-471 restated in the module)
        tol = max(tol, 2e-5)
        assert float(out["weighted_loss"]) == pytest.approx(float(z["f64_r0_weighted_loss"]), rel=2e-5)
        # the keys train.py:360-364 formats every 300 steps, as numbers
        msg = (f"{out['dbg']['delta_img_max']:.2f}/{out['dbg']['delta_txt_max']:.2f} "
               f"{out['dbg']['corr_rhat_dprob_img']:.3f}/{out['dbg']['corr_rhat_dprob_txt']:.3f} "
               f"{out['dbg']['pc_err_img']:.2e}/{out['dbg']['pc_err_txt']:.2e}")
        assert "nan" not in msg
        assert float(out["dbg"]["pc_err_img"]) < 1e-2 and float(out["dbg"]["diag_max_img"]) <= float(a.c_clip)
    else:
        assert float(out["weighted_loss"]) == 0.0 and out["dbg"] == {}
    for k in ("total_loss", "classic_loss", "soft_loss"):
        assert float(out[k]) == pytest.approx(float(z["f64_r0_" + k]), rel=tol, abs=1e-9), k
    for got, key in ((im.grad, "d_image"), (tx.grad, "d_text")):
        ref = z["f64_r0_" + key]
        err = np.abs(got.numpy() - ref).max() / np.abs(ref).max()
        assert err < max(tol * 4, 1e-6), (key, err)
    assert float(sc.grad) == pytest.approx(float(z["f64_r0_d_logit_scale"]), rel=max(tol, 1e-6), abs=1e-9)


def test_feature_store_needs_cuda(pkg):
    """No CPU path: the device feature store refuses a CPU device."""
    import torch

    with pytest.raises(RuntimeError, match="needs a CUDA device"):
        pkg.DinoFeatureStore(torch.zeros(4, 8), "cpu")
