"""Behaviour a training loop relies on beyond raw parity: dtypes, eval mode, CUDA-graph capture, repeat calls."""
import pytest
import torch

from gpu_util import make_args, oracle_cfg, rel_err, synth

pytestmark = pytest.mark.gpu


def _inputs(B=256, D=128, Dd=192, seed=1):
    img, txt, dino = synth(seed, B, D, Dd)
    return img.cuda(), txt.cuda(), dino.cuda()


def test_bf16_and_fp16_inputs_match_fp32_inputs(pkg):
    """Inputs are bf16-representable, so feeding them as bf16 / fp32 must give identical results."""
    img, txt, dino = _inputs()
    args = make_args()
    outs = []
    for dt in (torch.float32, torch.bfloat16):
        loss = pkg.ClipLossWithDINOEnhancements()
        im = img.detach().clone().to(dt).requires_grad_(True)
        tx = txt.detach().clone().to(dt).requires_grad_(True)
        sc = torch.tensor(20.0, device="cuda", requires_grad=True)
        o = loss(im, tx, sc, dino, args, output_dict=True)
        o["total_loss"].backward()
        assert im.grad.dtype == dt and tx.grad.dtype == dt
        outs.append((float(o["classic_loss"]), im.grad.float(), float(sc.grad)))
    # tau_t is rounded through the student dtype like the reference (loss.py:369): only the CLIP part is identical
    assert outs[0][0] == outs[1][0]
    linf, _ = rel_err(outs[1][1], outs[0][1])
    assert linf < 2e-2  # bf16 gradient storage + bf16(0.15) teacher temperature


def test_no_grad_eval_and_missing_dino(pkg):
    img, txt, dino = _inputs()
    loss = pkg.ClipLossWithDINOEnhancements()
    with torch.no_grad():
        o = loss(img, txt, torch.tensor(14.0, device="cuda"), dino, make_args(), output_dict=True)
    assert not o["total_loss"].requires_grad and torch.isfinite(o["total_loss"])
    # no DINO features for this batch (train.py zeroes lambda_soft, make_effective_args): classic only
    o2 = loss(img, txt, torch.tensor(14.0, device="cuda"), None, make_args(lambda_soft=0.0), output_dict=True)
    assert float(o2["soft_loss"]) == 0.0 and float(o2["total_loss"]) == float(o2["classic_loss"])
    assert float(o2["classic_loss"]) == pytest.approx(float(o["classic_loss"]), rel=1e-6)


def test_non_contiguous_and_repeated_calls(pkg):
    img, txt, dino = _inputs()
    wide = torch.zeros(img.shape[0], img.shape[1] * 2, device="cuda")
    wide[:, ::2] = img
    loss = pkg.ClipLossWithDINOEnhancements()
    args = make_args()
    a = loss(img, txt, torch.tensor(20.0, device="cuda"), dino, args, output_dict=True)["total_loss"]
    b = loss(wide[:, ::2], txt, torch.tensor(20.0, device="cuda"), dino, args, output_dict=True)["total_loss"]
    assert float(a) == float(b)
    # two forwards before two backwards (e.g. gradient accumulation of two micro-batches): per-call buffers
    i1 = img.clone().requires_grad_(True)
    i2 = img.flip(0).clone().requires_grad_(True)
    l1 = loss(i1, txt, torch.tensor(20.0, device="cuda"), dino, args, output_dict=True)["total_loss"]
    l2 = loss(i2, txt.flip(0), torch.tensor(20.0, device="cuda"), dino.flip(0), args, output_dict=True)["total_loss"]
    l1.backward()
    l2.backward()
    assert float(l1) == pytest.approx(float(l2), rel=1e-6)
    linf, _ = rel_err(i2.grad.flip(0), i1.grad)
    assert linf < 1e-4


def test_cuda_graph_capture_and_replay(pkg):
    """The whole loss fwd+bwd is stream-ordered with no host synchronisation, so it can be captured."""
    img, txt, dino = _inputs(512, 128, 192, seed=4)
    loss = pkg.ClipLossWithDINOEnhancements()
    args = make_args(use_projection=True)
    torch.manual_seed(0)
    loss.init_proj(128, 192, "cuda", "mlp")
    im = img.clone().requires_grad_(True)
    tx = txt.clone().requires_grad_(True)
    sc = torch.tensor(25.0, device="cuda", requires_grad=True)

    def step():
        out = loss(im, tx, sc, dino, args, output_dict=True)
        gi, gt, gs = torch.autograd.grad(out["total_loss"], [im, tx, sc])
        return out["total_loss"].detach(), gi, gt, gs

    eager = [t.clone() for t in step()]
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(2):
            step()
    torch.cuda.current_stream().wait_stream(s)
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        static = step()
    with torch.no_grad():
        im.copy_(img.flip(0))  # new data in the captured input buffers
        tx.copy_(txt.flip(0))
    g.replay()
    torch.cuda.synchronize()
    replay_flipped = [t.clone() for t in static]
    with torch.no_grad():
        im.copy_(img)
        tx.copy_(txt)
    g.replay()
    torch.cuda.synchronize()
    for a, b in zip(eager, static):
        assert torch.equal(a, b)
    assert not torch.equal(replay_flipped[1], eager[1])


def test_forked_and_serial_launches_agree_bitwise(pkg):
    """The independent tile kernels run on forked streams by default (include/dsoft.h, dsoft_set_concurrency);
    serial launches on the caller's stream must give bit-identical losses and gradients, also when the caller's
    stream is not the default stream."""
    from dinosoft_b200 import _cabi

    lib = _cabi.lib()
    img, txt, dino = _inputs(B=640, D=128, Dd=192, seed=5)
    args = make_args(text_lambda=0.5, use_projection=True)
    res = []
    side = torch.cuda.Stream()
    try:
        for on, stream in ((1, None), (0, None), (1, side)):
            lib.dsoft_set_concurrency(on)
            loss = pkg.ClipLossWithDINOEnhancements()
            torch.manual_seed(0)  # same lazily created MLP head in every variant
            loss.init_proj(img.shape[1], dino.shape[1], img.device, projection_type="mlp")
            im = img.clone().requires_grad_(True)
            tx = txt.clone().requires_grad_(True)
            sc = torch.tensor(25.0, device="cuda", requires_grad=True)
            ctx = torch.cuda.stream(stream) if stream is not None else torch.cuda.stream(torch.cuda.current_stream())
            if stream is not None:
                stream.wait_stream(torch.cuda.current_stream())
            with ctx:
                o = loss(im, tx, sc, dino, args, output_dict=True)
                o["total_loss"].backward()
            torch.cuda.synchronize()
            res.append((o["total_loss"].detach().clone(), im.grad.clone(), tx.grad.clone(), sc.grad.clone(),
                        [prm.grad.clone() for prm in loss.parameters()]))
    finally:
        lib.dsoft_set_concurrency(-1)  # back to the automatic choice
    for other in res[1:]:
        assert torch.equal(res[0][0], other[0])
        assert torch.equal(res[0][1], other[1]) and torch.equal(res[0][2], other[2])
        assert torch.equal(res[0][3], other[3])
        for a, b in zip(res[0][4], other[4]):
            assert torch.equal(a, b)


def test_backward_falls_back_when_the_gradient_matrices_do_not_fit(pkg):
    """Two-phase backward chosen in the forward, but its scratch cannot be allocated at backward time: the fused
    backward must take over on the same saved state and give the same gradients (to rounding of the sums)."""
    from dinosoft_b200 import loss as loss_mod

    img, txt, dino = _inputs(B=384, D=128, Dd=192, seed=9)
    args = make_args(use_projection=True)
    old = loss_mod.GMAT
    grads = []
    try:
        for sabotage in (False, True):
            loss_mod.GMAT = "always"
            m = pkg.ClipLossWithDINOEnhancements()
            torch.manual_seed(0)
            m.init_proj(img.shape[1], dino.shape[1], img.device, "mlp")
            im = img.clone().requires_grad_(True)
            tx = txt.clone().requires_grad_(True)
            sc = torch.tensor(20.0, device="cuda", requires_grad=True)
            out = m(im, tx, sc, dino, args, output_dict=True)["total_loss"]
            plan = next(p for p in loss_mod._cuda_backend._plans.values()
                        if p.shape.b == 384 and p.shape.flags & 16)
            keep = plan.scratch_numel
            if sabotage:
                plan.scratch_numel = 1 << 42  # 16 TiB of fp32: torch raises OutOfMemoryError
            try:
                out.backward()
            finally:
                plan.scratch_numel = keep
            torch.cuda.synchronize()
            grads.append((im.grad.clone(), tx.grad.clone(), sc.grad.clone()))
    finally:
        loss_mod.GMAT = old
        loss_mod._gmat_decisions.clear()
    for a, b in zip(grads[0], grads[1]):
        linf, l2 = rel_err(b, a)
        assert linf < 2e-4 and l2 < 2e-4
