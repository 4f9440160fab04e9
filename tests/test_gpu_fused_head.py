"""Fused projection head (dsoft_head_forward): under bf16 autocast the Linear / Linear-ReLU-Linear head
runs on the tcgen05 tile kernel and writes the student columns of the packed buffer.  Checked against the PyTorch
head under the same autocast (the arithmetic the reference gets at train.py:285): bf16 outputs may differ by one
rounding step on isolated elements, nothing more."""
import ctypes as C

import pytest
import torch

from gpu_util import make_args, rel_err, synth

pytestmark = pytest.mark.gpu


def _run(pkg, img, txt, dino, args, fused, seed=3, **ctor):
    from dinosoft_b200 import loss as loss_mod

    old = loss_mod.FUSED_HEAD
    loss_mod.FUSED_HEAD = fused
    try:
        dev = "cuda"
        loss = pkg.ClipLossWithDINOEnhancements(**ctor)
        torch.manual_seed(seed)
        loss.init_proj(img.shape[1], dino.shape[1], dev, args.projection_type)
        im = img.to(dev).requires_grad_(True)
        tx = txt.to(dev).requires_grad_(True)
        sc = torch.tensor(25.0, device=dev, requires_grad=True)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            out = loss(im, tx, sc, dino.to(dev), args, output_dict=True)
        out["total_loss"].backward()
        torch.cuda.synchronize()
        grads = {n: p.grad.detach().clone() for n, p in loss.image_to_dino_proj.named_parameters()}
        return out, im.grad, tx.grad, sc.grad, grads
    finally:
        loss_mod.FUSED_HEAD = old


@pytest.mark.parametrize("ptype", ["mlp", "linear"])
@pytest.mark.parametrize("B,D,Dd", [(512, 128, 192), (1000, 512, 768), (300, 72, 136)])
def test_fused_head_matches_autocast_head(pkg, ptype, B, D, Dd):
    img, txt, dino = synth(17, B, D, Dd)
    args = make_args(use_projection=True, projection_type=ptype)
    o1, gi1, gt1, gs1, gh1 = _run(pkg, img, txt, dino, args, True)
    o0, gi0, gt0, gs0, gh0 = _run(pkg, img, txt, dino, args, False)
    for k in ("total_loss", "classic_loss", "soft_loss"):
        a, b = float(o1[k]), float(o0[k])
        print(f"[fused head {ptype}] {k}: fused {a:.6f} torch {b:.6f}")
        # isolated one-step bf16 differences of the student move the soft term at the 1e-4 level
        assert a == pytest.approx(b, rel=1e-3), k
    assert float(o1["classic_loss"]) == pytest.approx(float(o0["classic_loss"]), rel=1e-6)
    for name, a, b in (("d_image", gi1, gi0), ("d_text", gt1, gt0)):
        linf, l2 = rel_err(a, b)
        print(f"[fused head {ptype}] {name}: linf={linf:.2e} l2={l2:.2e}")
        assert l2 < 1e-2, (name, linf, l2)
    for n in gh0:
        # the PyTorch path rounds every weight gradient to bf16 (the matmul output dtype under autocast); the
        # fused path keeps the fp32 accumulator, so the comparison is at bf16 resolution
        linf, l2 = rel_err(gh1[n], gh0[n])
        print(f"[fused head {ptype}] d {n}: linf={linf:.2e} l2={l2:.2e}")
        assert l2 < 3e-2, (n, linf, l2)


def test_head_forward_values(pkg):
    """dsoft_head_forward alone: student columns and hidden activations against fp32 torch on the same bf16 operands."""
    from dinosoft_b200 import _cabi
    from dinosoft_b200 import loss as loss_mod

    dev = torch.device("cuda", 0)
    be = loss_mod._default_backend(dev)
    b, D, H, Dp, Dd = 700, 512, 640, 768, 768
    g = torch.Generator().manual_seed(1)
    shape = _cabi.Shape(b=b, world=1, rank=0, D=D, Dp=Dp, Dd=Dd, flags=_cabi.DSOFT_F_SOFT, teacher_temp=0.15,
                        text_temp=0.0, rho=0.1, c_clip=1.0)
    plan = be.plan(shape, dev)
    img = torch.nn.functional.normalize(torch.randn(b, D, generator=g), dim=-1).to(dev)
    txt = torch.nn.functional.normalize(torch.randn(b, D, generator=g), dim=-1).to(dev)
    dino = torch.randn(b, Dd, generator=g).to(dev)
    w1 = (torch.randn(H, D, generator=g) / D ** 0.5).to(dev).to(torch.bfloat16)
    w2 = (torch.randn(Dp, H, generator=g) / H ** 0.5).to(dev).to(torch.bfloat16)
    b1 = (0.1 * torch.randn(H, generator=g)).to(dev)
    b2 = (0.1 * torch.randn(Dp, generator=g)).to(dev)
    gathered = torch.zeros((b, plan.row_elems), dtype=torch.bfloat16, device=dev)
    be.pack(plan, img, txt, None, dino, gathered)
    hidden = torch.empty((b, H), dtype=torch.bfloat16, device=dev)
    be.head_forward(plan, gathered, w1, b1, w2, b2, hidden)
    torch.cuda.synchronize()
    x = gathered[:, :D].float()
    h_ref = torch.relu(x @ w1.float().t() + b1)
    assert torch.equal(gathered[:, :D], img.to(torch.bfloat16))
    dh = (hidden.float() - h_ref).abs().max().item()
    assert dh <= 2 ** -8 * max(1.0, h_ref.abs().max().item()), dh            # one bf16 rounding of the output
    z_ref = hidden.float() @ w2.float().t() + b2                               # from the kernel's own hidden
    z = gathered[:, 2 * D:2 * D + Dp].float()
    dz = (z - z_ref).abs().max().item()
    assert dz <= 2 ** -8 * max(1.0, z_ref.abs().max().item()), dz
    assert (z - z_ref).abs().mean().item() < 2e-3 * z_ref.abs().mean().item()
    # linear head
    wl = (torch.randn(Dp, D, generator=g) / D ** 0.5).to(dev).to(torch.bfloat16)
    be.head_forward(plan, gathered, wl, b2, None, None, None)
    torch.cuda.synchronize()
    zl_ref = x @ wl.float().t() + b2
    zl = gathered[:, 2 * D:2 * D + Dp].float()
    assert (zl - zl_ref).abs().max().item() <= 2 ** -8 * max(1.0, zl_ref.abs().max().item())
