"""dinosoft_b200.make_graphed: the loss forward + backward replayed as CUDA graphs give the eager results."""
import pytest
import torch

from gpu_util import make_args, rel_err, synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("autocast", [None, torch.bfloat16])
def test_graphed_equals_eager(pkg, autocast):
    dev = "cuda"
    B, D, Dd = 1024, 256, 384
    args = make_args(use_projection=True)
    loss = pkg.ClipLossWithDINOEnhancements()
    torch.manual_seed(5)
    loss.init_proj(D, Dd, dev, "mlp")
    img, txt, dino = (t.to(dev) for t in synth(40, B, D, Dd))
    scale = torch.tensor(20.0, device=dev, requires_grad=True)
    img.requires_grad_(True)
    txt.requires_grad_(True)
    step = pkg.make_graphed(loss, args, img, txt, scale, dino, autocast_dtype=autocast)
    params = list(loss.image_to_dino_proj.parameters())

    def grads():
        out = [img.grad.clone(), txt.grad.clone(), scale.grad.clone()] + [p.grad.clone() for p in params]
        img.grad = txt.grad = scale.grad = None
        for p in params:
            p.grad = None
        return out

    for trial in range(2):  # fresh inputs on the second replay: the static buffers must be refreshed
        if trial == 1:
            with torch.no_grad():
                a, b, c = (t.to(dev) for t in synth(41, B, D, Dd))
                img.copy_(a); txt.copy_(b); dino.copy_(c)
        t, c_, s_ = step(img, txt, scale, dino)
        t.backward()
        got = [t.detach().clone(), c_.detach().clone(), s_.detach().clone()] + grads()
        if autocast is None:
            out = loss(img, txt, scale, dino, args, output_dict=True)
        else:
            with torch.autocast("cuda", dtype=autocast):
                out = loss(img, txt, scale, dino, args, output_dict=True)
        out["total_loss"].backward()
        want = [out["total_loss"].detach(), out["classic_loss"].detach(), out["soft_loss"].detach()] + grads()
        for k in range(3):
            assert float(got[k]) == pytest.approx(float(want[k]), rel=1e-6), (trial, k)
        for g, w in zip(got[3:], want[3:]):
            assert rel_err(g, w)[0] < 1e-5, trial
