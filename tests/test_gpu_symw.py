"""Symmetric soft tiles across ranks (DSOFT_SYM_W, include/dsoft.h): every pair of row blocks of the teacher /
student / text Gram matrices is computed by ONE of its two ranks; column sums (forward) and transposed gradient
products (backward) for the other rank's rows are exchanged between the phases of the C calls (4 operand statistics, 1 soft
part, 3 CLIP part - independent of 1 and of the exchange -, 2 finalize).

All ranks of a W-rank job are played on ONE GPU through the C ABI, the two exchanges being done by hand exactly as
`_SymW.exchange_forward / exchange_backward` do over NCCL (same layout object).  Every rank is compared with the
fp64 oracle of the reference semantics (global soft scope, gather_with_grad): loss 1e-4, gradients 1e-3."""
import pytest
import torch

from gpu_util import rel_err, synth
from test_gpu_multirank import GRAD_RTOL, LAMBDAS, LOSS_RTOL, TEACHER_TEMP, TEXT_TEMP, oracle_ranks

pytestmark = pytest.mark.gpu


def cuda_ranks_symw(pkg, img, txt, dino, student, scale, W):
    from dinosoft_b200 import _cabi
    from dinosoft_b200.loss import CudaBackend

    dev = torch.device("cuda", 0)
    be = CudaBackend()
    B, D = img.shape
    b = B // W
    flags = _cabi.DSOFT_F_SOFT | _cabi.DSOFT_F_TEXT | _cabi.DSOFT_F_GMAT
    Dp = 0 if student is None else student.shape[1]
    plans = [be.plan(_cabi.Shape(b=b, world=W, rank=r, D=D, Dp=Dp, Dd=dino.shape[1], flags=flags,
                                 teacher_temp=TEACHER_TEMP, text_temp=TEXT_TEMP), dev) for r in range(W)]
    assert all(pl.symw is not None for pl in plans), "the plans should share the symmetric tiles"
    gathered = torch.empty((B, plans[0].row_elems), dtype=torch.bfloat16, device=dev)
    cu = lambda t: None if t is None else t.to(dev)
    for r, pl in enumerate(plans):
        rows = slice(r * b, (r + 1) * b)
        be.pack(pl, cu(img[rows]), cu(txt[rows]), cu(None if student is None else student[rows]), cu(dino[rows]),
                gathered)
    ls = torch.tensor([scale], dtype=torch.float32, device=dev)
    lse_all = torch.empty((W, 5, b), dtype=torch.float32, device=dev)
    states = [torch.empty(pl.state_numel, dtype=torch.float32, device=dev) for pl in plans]
    fscr = [torch.empty(pl.fwd_scratch_numel, dtype=torch.float32, device=dev) for pl in plans]
    losses = [torch.empty(6, dtype=torch.float32, device=dev) for _ in plans]
    for r, pl in enumerate(plans):
        for phase in (4, 3, 1):  # the CLIP part needs the operand statistics only
            be.forward(pl, gathered, ls, LAMBDAS, states[r], fscr[r], lse_all[r], losses[r], phase=phase)
    # ---- forward exchange: column sums of primed block k -> rank (r + k) % W
    inbox = [torch.zeros((6, b), dtype=torch.float32, device=dev) for _ in plans]
    for r, pl in enumerate(plans):
        cs = pl.symw.colsum(fscr[r])
        for peer, k, rows in pl.symw.sends:
            inbox[peer][:, :rows] += cs[:, k * b:k * b + rows]
    for r, pl in enumerate(plans):
        pl.symw.colsum(fscr[r])[:, :b] += inbox[r]
        be.forward(pl, gathered, ls, LAMBDAS, states[r], fscr[r], lse_all[r], losses[r], phase=2)
    # ---- backward
    gout = torch.tensor([0.0, 0.0, 0.0, 0.0, 1.0, 0.0], dtype=torch.float32, device=dev)
    scr = [torch.empty(pl.scratch_numel, dtype=torch.float32, device=dev) for pl in plans]
    outs = []
    for r, pl in enumerate(plans):
        d_image = torch.empty((b, D), dtype=torch.float32, device=dev)
        d_text = torch.empty((b, D), dtype=torch.float32, device=dev)
        d_student = torch.empty((b, Dp), dtype=torch.float32, device=dev) if Dp else None
        d_scale = torch.empty(1, dtype=torch.float32, device=dev)
        outs.append((d_image, d_text, d_student, d_scale))
        for phase in (4, 3, 1):
            be.backward(pl, gathered, states[r], scr[r], lse_all, gout, LAMBDAS, d_image, d_text, d_student, d_scale,
                        phase=phase)
    # the receive lists must mirror the send lists
    for r, pl in enumerate(plans):
        expect = sorted((s, k, rows) for s, q in enumerate(plans) for peer, k, rows in q.symw.sends if peer == r)
        assert expect == sorted(pl.symw.recvs), (r, expect, pl.symw.recvs)
    for which in (0, 1):
        for r, pl in enumerate(plans):
            rem = pl.symw.remote(scr[r], which)
            for peer, k, rows in pl.symw.sends:
                plans[peer].symw.own(scr[peer], which)[:rows] += rem[(k - 1) * b:(k - 1) * b + rows]
    res = []
    for r, pl in enumerate(plans):
        d_image, d_text, d_student, d_scale = outs[r]
        be.backward(pl, gathered, states[r], scr[r], lse_all, gout, LAMBDAS, d_image, d_text, d_student, d_scale,
                    phase=2)
        torch.cuda.synchronize()
        lo = losses[r].cpu()
        res.append(dict(classic=float(lo[0]), soft=float(lo[3]), total=float(lo[4]), d_image=d_image.cpu(),
                        d_text=d_text.cpu(), d_student=None if d_student is None else d_student.cpu(),
                        d_scale=float(d_scale)))
    return res


@pytest.mark.parametrize("W,b,proj", [(2, 512, True), (4, 512, False), (8, 512, True), (3, 512, False), (2, 1024, False)])
def test_every_rank_against_oracle(pkg, oracle, W, b, proj, monkeypatch):
    monkeypatch.setenv("DSOFT_SYM_W", "1")  # small blocks: the plan would not share the tiles on its own
    D, Dd = 128, 192
    B, scale = W * b, 30.0
    img, txt, dino = synth(57 + W, B, D, Dd)
    student = None
    if proj:
        g = torch.Generator().manual_seed(W)
        mix = torch.randn(D, Dd, generator=g) / D ** 0.5
        student = ((img @ mix) * 2.5 + 0.1 * torch.randn(B, Dd, generator=g)).to(torch.bfloat16).float()
    want = oracle_ranks(oracle, img, txt, dino, student, scale, W, "global", True)
    got = cuda_ranks_symw(pkg, img, txt, dino, student, scale, W)
    for r in range(W):
        o, ref = got[r], want[r]
        t = ref["terms"]
        for k, key in (("classic", "classic_loss"), ("soft", "soft_loss"), ("total", "total_loss")):
            assert o[k] == pytest.approx(float(t[key]), rel=LOSS_RTOL, abs=1e-5), (r, k, o[k], float(t[key]))
        worst = 0.0
        for k in ("d_image", "d_text", "d_student"):
            if ref[k] is None:
                continue
            linf, l2 = rel_err(o[k], ref[k])
            worst = max(worst, linf, l2)
            assert linf < GRAD_RTOL and l2 < GRAD_RTOL, (f"rank {r} of {W}", k, linf, l2)
        assert o["d_scale"] == pytest.approx(ref["d_scale"], rel=GRAD_RTOL, abs=1e-7), (r, "d_scale")
        print(f"[symw] W={W} b={b} rank {r}: worst grad err {worst:.2e}")
