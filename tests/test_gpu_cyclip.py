"""CyCLIPLoss drop-in on the GPU against the CPU oracle and the reference fixture: CLIP term on the
tcgen05 kernels, consistency terms through the D x D moment matrices."""
import importlib.util
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN_DIR, ROOT
from gpu_util import rel_err, synth

pytestmark = pytest.mark.gpu


def oracle():
    spec = importlib.util.spec_from_file_location("cyclip_oracle", os.path.join(ROOT, "oracle", "cyclip_oracle.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def run(pkg, img, txt, scale, li, lc, output_dict=True):
    from dinosoft_b200.cyclip import CyCLIPLoss

    loss = CyCLIPLoss(lambda_inmodal=li, lambda_crossmodal=lc)
    im = img.cuda().requires_grad_(True)
    tx = txt.cuda().requires_grad_(True)
    sc = torch.tensor(scale, device="cuda", requires_grad=True)
    out = loss(im, tx, sc, output_dict=output_dict)
    (out["total_loss"] if output_dict else out).backward()
    torch.cuda.synchronize()
    return out, im.grad, tx.grad, sc.grad


def test_reference_fixture(pkg):
    z = np.load(os.path.join(GOLDEN_DIR, "cyclip_b96.npz"))
    out, gi, gt, gs = run(pkg, torch.from_numpy(z["image"]), torch.from_numpy(z["text"]), float(z["scale"]),
                          float(z["lambda_inmodal"]), float(z["lambda_crossmodal"]))
    assert set(out) == {"total_loss", "clip_loss", "inmodal_cyclic", "crossmodal_cyclic", "lambda_inmodal",
                        "lambda_crossmodal"}
    for k in ("total_loss", "clip_loss", "inmodal_cyclic", "crossmodal_cyclic"):
        assert float(out[k].detach()) == pytest.approx(float(z["f64_" + k]), rel=1e-4), k
    for got, k in ((gi, "d_image"), (gt, "d_text")):
        linf, l2 = rel_err(got, torch.from_numpy(z["f64_" + k]))
        assert linf < 1e-3 and l2 < 1e-3, (k, linf, l2)
    assert float(gs) == pytest.approx(float(z["f64_d_logit_scale"]), rel=1e-3, abs=1e-7)


@pytest.mark.parametrize("B,D,scale", [(512, 128, 14.2857), (1000, 512, 50.0), (2048, 512, 30.0)])
def test_against_oracle(pkg, B, D, scale):
    img, txt, _ = synth(31, B, D, 64)
    txt = torch.nn.functional.normalize(0.6 * img + 0.8 * txt, dim=-1).to(torch.bfloat16).float()
    want = oracle().loss_and_grads(img, txt, scale, 0.25, 0.5)
    out, gi, gt, gs = run(pkg, img, txt, scale, 0.25, 0.5)
    for k in ("total_loss", "clip_loss", "inmodal_cyclic", "crossmodal_cyclic"):
        got = float(out[k].detach())
        print(f"[cyclip] B={B} {k}: got {got:.7f} ref {want[k]:.7f}")
        # abs floor as in test_gpu_parity: logits of magnitude ~scale carry an fp32 ulp of up to 8e-6 into lse - L_ii
        assert got == pytest.approx(want[k], rel=1e-4, abs=1e-5), k
    for got, k in ((gi, "d_image"), (gt, "d_text")):
        linf, l2 = rel_err(got, want[k])
        print(f"[cyclip] B={B} {k}: linf={linf:.2e} l2={l2:.2e}")
        assert linf < 1e-3 and l2 < 1e-3, (k, linf, l2)
    assert float(gs) == pytest.approx(want["d_logit_scale"], rel=1e-3, abs=1e-7)


def test_returns_total_without_output_dict_and_runs_at_32768(pkg):
    img, txt, _ = synth(1, 256, 64, 64)
    total, *_ = run(pkg, img, txt, 20.0, 0.25, 0.25, output_dict=False)
    assert total.dim() == 0
    # n = 32768: the reference needs four 4 GiB fp32 matrices plus their autograd copies; here nothing is n x n
    g = torch.Generator(device="cuda").manual_seed(0)
    big_i = torch.nn.functional.normalize(torch.randn(32768, 512, device="cuda", generator=g), dim=-1)
    big_t = torch.nn.functional.normalize(big_i + 0.7 * torch.randn(32768, 512, device="cuda", generator=g), dim=-1)
    torch.cuda.reset_peak_memory_stats()
    out, gi, gt, gs = run(pkg, big_i, big_t, 14.2857, 0.25, 0.25)
    assert torch.isfinite(out["total_loss"]) and torch.isfinite(gi).all() and torch.isfinite(gt).all()
    # sampled exactness: the consistency terms of a 2048-row subset against the literal evaluation in fp64
    sub_i, sub_t = big_i[:2048].double(), big_t[:2048].double()
    lit = float(((sub_i @ sub_i.t() - sub_t @ sub_t.t()) ** 2).mean())
    from dinosoft_b200.cyclip import _CyclicFn

    a, _ = _CyclicFn.apply(big_i[:2048], big_t[:2048])
    assert float(a) == pytest.approx(lit, rel=1e-4)
