"""Parity of the CUDA path (through the module -> C ABI) with the CPU oracle on identical inputs.

Tolerances are north_star's: loss <= 1e-4 relative, gradients <= 1e-3 relative (max-abs over max-abs and
L2 over L2), fp32 accumulation, bf16 operands."""
import numpy as np
import pytest
import torch

from gpu_util import head_params_of, make_args, oracle_cfg, rel_err, synth, synth_aligned

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["two_phase", "fused"])
def backward_path(request, pkg):
    """Both backward implementations must meet the same bar: the two-phase one (fp16 logit-gradient matrices +
    gradient GEMMs, DSOFT_F_GMAT) and the fused one that keeps the logit gradients in shared memory."""
    from dinosoft_b200 import loss as loss_mod

    old = loss_mod.GMAT
    loss_mod.GMAT = "always" if request.param == "two_phase" else "never"
    yield request.param
    loss_mod.GMAT = old

LOSS_RTOL = 1e-4
GRAD_RTOL = 1e-3


def run_cuda(pkg, img, txt, dino, scale, args, head_seed=7, **ctor):
    dev = "cuda"
    loss = pkg.ClipLossWithDINOEnhancements(**ctor)
    im = img.to(dev).requires_grad_(True)
    tx = txt.to(dev).requires_grad_(True)
    sc = torch.tensor(scale, device=dev, requires_grad=True)
    dn = None if dino is None else dino.to(dev)
    if dn is not None and args.use_projection:
        torch.manual_seed(head_seed)
        loss.init_proj(img.shape[1], dino.shape[1], dev, args.projection_type,
                       layernorm=getattr(args, "use_layernorm", False))
    grabbed = {}
    hook = None
    if loss.image_to_dino_proj is not None:  # gradient w.r.t. the raw head output = the kernels' d_student
        hook = loss.image_to_dino_proj.register_forward_hook(
            lambda mod, inp, o: (grabbed.__setitem__("student", o.detach().to(torch.bfloat16).float().cpu()),
                                 o.register_hook(lambda g: grabbed.__setitem__("d_student", g.detach().clone())))
            and None)
    out = loss(im, tx, sc, dn, args, output_dict=True)
    out["total_loss"].backward()
    torch.cuda.synchronize()
    if hook is not None:
        hook.remove()
    loss._test_d_student = grabbed.get("d_student")
    loss._test_student = grabbed.get("student")
    return loss, out, im.grad, tx.grad, sc.grad


def check_against_oracle(pkg, oracle, B, D, Dd, scale, args, seed=0, clustered=True, inputs=None):
    img, txt, dino = synth(seed, B, D, Dd, clustered) if inputs is None else inputs
    loss, out, gi, gt, gs = run_cuda(pkg, img, txt, dino, scale, args)
    head = None
    if args.use_projection and loss.image_to_dino_proj is not None:
        head = {k: v.detach().cpu() for k, v in
                head_params_of(loss.image_to_dino_proj, args.projection_type,
                               getattr(args, "use_layernorm", False)).items()}
    cfg = oracle_cfg(oracle, args, round_student_bf16=True)
    # the oracle evaluates the head in fp64; it is handed the VALUES of the student operand the CUDA path used
    # (its fp32 head output rounded to bf16), so that a value on a bf16 rounding boundary cannot differ
    sv = loss._test_student if (head is not None and not getattr(args, "residual_projection", False)) else None
    ref = oracle.loss_and_grads(img, txt, scale, dino, cfg, proj_params=head,
                                projection_type=args.projection_type, dtype=torch.float64,
                                student_values=sv)["ranks"][0]
    for k in ("total_loss", "classic_loss", "soft_loss", "weighted_loss"):
        got = float(out[k].detach())
        print(f"[parity] B={B} D={D} Dd={Dd} s={scale} {k}: got={got:.7f} ref={ref[k]:.7f} rel={abs(got - ref[k]) / max(abs(ref[k]), 1e-30):.2e}")
        # abs floor: logits of magnitude ~scale carry an fp32 ulp of up to 8e-6 into (lse - L_ii)
        assert got == pytest.approx(ref[k], rel=LOSS_RTOL, abs=1e-5), (k, got, ref[k])
    for name, got, want in (("d_image", gi, ref["d_image"]), ("d_text", gt, ref["d_text"])):
        linf, l2 = rel_err(got, want)
        print(f"[parity] B={B} D={D} Dd={Dd} s={scale} {name}: linf={linf:.2e} l2={l2:.2e}")
        assert linf < GRAD_RTOL and l2 < GRAD_RTOL, (name, linf, l2)
    print(f"[parity] d_logit_scale got={float(gs):.6e} ref={ref['d_logit_scale']:.6e}")
    assert float(gs) == pytest.approx(ref["d_logit_scale"], rel=GRAD_RTOL, abs=1e-7)
    if "d_student" in ref and loss._test_d_student is not None and not getattr(args, "residual_projection", False):
        linf, l2 = rel_err(loss._test_d_student, ref["d_student"])
        print(f"[parity] B={B} D={D} Dd={Dd} s={scale} d_student: linf={linf:.2e} l2={l2:.2e}")
        assert linf < GRAD_RTOL and l2 < GRAD_RTOL, ("d_student", linf, l2)
    if head is not None and "d_proj" in ref:
        got_head = head_params_of(loss.image_to_dino_proj, args.projection_type, getattr(args, "use_layernorm", False))
        for k, want in ref["d_proj"].items():
            if want is None:
                continue
            linf, l2 = rel_err(got_head[k].grad, want)
            print(f"[parity] head {k}: linf={linf:.2e} l2={l2:.2e}")
            assert linf < GRAD_RTOL and l2 < GRAD_RTOL, ("head " + k, linf, l2)
    return out, ref


def test_classic_only_small(pkg, oracle):
    check_against_oracle(pkg, oracle, 256, 512, 768, 14.2857, make_args(lambda_soft=0.0, soft_mode="none"))


def test_config1_noproj_text(pkg, oracle):
    """BASELINE config 1 sizes (B=256, D=512, Dd=768), student = image features."""
    check_against_oracle(pkg, oracle, 256, 512, 768, 14.2857, make_args())


def test_config1_mlp_text(pkg, oracle):
    check_against_oracle(pkg, oracle, 256, 512, 768, 14.2857, make_args(use_projection=True))


def test_scale_100_linear_head(pkg, oracle):
    check_against_oracle(pkg, oracle, 384, 512, 768, 100.0,
                         make_args(use_projection=True, projection_type="linear", soft_dino_to_text=False))


DBG_TOL = {  # relative tolerance per diagnostic (fp32 kernels vs the fp64 restatement in tests/weighted_ref.py)
    "pos_frac_img": 2e-3, "neg_frac_img": 2e-3, "pos_frac_txt": 2e-3, "neg_frac_txt": 2e-3,
    "corr_rhat_dprob_img": 2e-3, "corr_rhat_dprob_txt": 2e-3,
}


@pytest.mark.parametrize("sym", [False, True])
@pytest.mark.parametrize("B,D,Dd,scale,proj,soft", [
    (256, 128, 192, 30.0, True, True),      # everything on: classic + soft + weighted
    (2048, 128, 192, 60.0, False, True),    # several column splits / row pairs
    (200, 72, 136, 20.0, False, False),     # ragged, weighted without the soft term
])
def test_weighted_ce_branch(pkg, oracle, sym, B, D, Dd, scale, proj, soft):
    """lambda_weighted > 0 (loss.py:416-471 + diagnostics 479-595) through the fused kernels (DSOFT_F_WEIGHTED)."""
    from weighted_ref import weighted_ce_branch

    args = make_args(use_projection=proj, lambda_weighted=0.6, rho=0.2, c_clip=0.5, weight_text_symmetry=sym,
                     lambda_soft=0.5 if soft else 0.0, soft_mode="kl_teacher" if soft else "none")
    out, ref = check_against_oracle(pkg, oracle, B, D, Dd, scale, args, seed=11)
    assert float(out["weighted_loss"]) > 0.0
    img, txt, dino = synth(11, B, D, Dd)
    _, want = weighted_ce_branch(img.double(), txt.double(), torch.tensor(scale, dtype=torch.float64), dino.double(),
                                 0.2, 0.5, sym)
    assert set(out["dbg"]) == set(want)
    for k, w in want.items():
        got, w = float(out["dbg"][k]), float(w)
        tol = DBG_TOL.get(k, 5e-4)
        print(f"[weighted dbg] {k}: got={got:.6e} ref={w:.6e}")
        # pc_err is a cancellation residue (exactly 0 without the clamp): absolute floor of an fp32 row sum
        assert got == pytest.approx(w, rel=tol, abs=2e-5 if k.startswith("pc_err") else 1e-6), k


def test_iid_gaussian_flat_teacher(pkg, oracle):
    check_against_oracle(pkg, oracle, 512, 256, 384, 30.0, make_args(), seed=3, clustered=False)


@pytest.mark.parametrize("B,D,Dd", [(200, 72, 136), (130, 64, 64), (1000, 320, 200)])
def test_ragged_sizes(pkg, oracle, B, D, Dd):
    """Rows/columns that are not multiples of the 128 x 128 tile, K not a multiple of 64."""
    check_against_oracle(pkg, oracle, B, D, Dd, 20.0, make_args(use_projection=True), seed=5)


def test_column_split_many_tiles(pkg, oracle):
    """B large enough that every kernel splits the column range over several CTAs."""
    check_against_oracle(pkg, oracle, 2048, 128, 192, 50.0, make_args(use_projection=True), seed=9)


@pytest.mark.parametrize("scale,noise", [(14.2857, 0.3), (100.0, 0.3), (100.0, 1.0)])
def test_near_converged_pairs(pkg, oracle, scale, noise):
    """p_ii ~ 1: the CE gradient is the small residual (p_ii - 1) t_i + ...; the diagonal is handled in fp32."""
    inputs = synth_aligned(11, 256, 128, 192, noise)
    check_against_oracle(pkg, oracle, 256, 128, 192, scale, make_args(), inputs=inputs)
    check_against_oracle(pkg, oracle, 256, 128, 192, scale, make_args(use_projection=True), inputs=inputs)


def test_output_dict_false_returns_none(pkg):
    img, txt, dino = synth(0, 128, 64, 64)
    loss = pkg.ClipLossWithDINOEnhancements()
    assert loss(img.cuda(), txt.cuda(), torch.tensor(10.0).cuda(), dino.cuda(), make_args()) is None


def test_matches_golden_fixture_directly(pkg):
    """CUDA path vs the reference-generated fixture (no projection, so no operand-rounding caveat)."""
    import os
    from conftest import GOLDEN_DIR

    z = np.load(os.path.join(GOLDEN_DIR, "w1_noproj_text.npz"))
    img, txt, dino = (torch.from_numpy(z[k]) for k in ("image", "text", "dino"))
    args = make_args(use_projection=False)
    _, out, gi, gt, gs = run_cuda(pkg, img, txt, dino, float(z["scale"]), args)
    assert float(out["total_loss"]) == pytest.approx(float(z["f64_r0_total_loss"]), rel=LOSS_RTOL)
    assert float(out["classic_loss"]) == pytest.approx(float(z["f64_r0_classic_loss"]), rel=LOSS_RTOL)
    assert float(out["soft_loss"]) == pytest.approx(float(z["f64_r0_soft_loss"]), rel=LOSS_RTOL)
    for got, key in ((gi, "d_image"), (gt, "d_text")):
        linf, l2 = rel_err(got, torch.from_numpy(z["f64_r0_" + key]))
        print(f"[parity] golden {key}: linf={linf:.2e} l2={l2:.2e}")
        assert linf < GRAD_RTOL and l2 < GRAD_RTOL, (key, linf, l2)
    assert float(gs) == pytest.approx(float(z["f64_r0_d_logit_scale"]), rel=GRAD_RTOL, abs=1e-7)


def test_config2_b4096_mlp_head(pkg, oracle):
    """BASELINE config 2: global batch 4096, D=512, DINOv2-B dim 768, MLP head (fp64 oracle, ~2 GB host RAM)."""
    check_against_oracle(pkg, oracle, 4096, 512, 768, 14.2857, make_args(use_projection=True), seed=2)


def test_config4_dims_vitl(pkg, oracle):
    """BASELINE config 4 feature dims (ViT-L/14: D=768, DINOv2-L: 1024, four 256-feature chunks with the head)."""
    check_against_oracle(pkg, oracle, 640, 768, 1024, 50.0, make_args(use_projection=False), seed=4)
    check_against_oracle(pkg, oracle, 384, 768, 1024, 50.0, make_args(use_projection=True), seed=4)


@pytest.mark.parametrize("scale,teacher_temp,text_temp", [
    (1.0, 0.15, 0.05),     # scale <= 10 is treated as a raw ln-scale by compute_student_tau (loss.py:172)
    (200.0, 0.15, 0.02),   # beyond the model's clamp: CLIP logits use it as is, tau_s saturates at 0.01
    (60.0, 0.03, 0.01),    # very sharp teacher / text student: exponent ranges of +-70 / +-290 in log2 units
    (14.2857, 1.0, 0.5),   # nearly flat teacher
])
def test_temperature_and_scale_extremes(pkg, oracle, scale, teacher_temp, text_temp):
    args = make_args(use_projection=True, teacher_temp=teacher_temp, text_student_temp=text_temp)
    check_against_oracle(pkg, oracle, 384, 128, 192, scale, args, seed=13)
    args = make_args(use_projection=False, teacher_temp=teacher_temp, text_student_temp=text_temp)
    check_against_oracle(pkg, oracle, 384, 128, 192, scale, args, seed=14, clustered=False)


def test_clip_one_pass_risky_chunks(pkg, oracle):
    """One-pass CLIP forward (MODE_CLIP_SYM): at scale 100 a 32 x 32 chunk that holds a near-duplicate pair (logit
    ~ +144 in log2 units) next to rows / columns whose best logit is far lower cannot share one exponent reference;
    those chunks must take the exact path (per-row and per-column maxima).  Anti-aligned and duplicate pairs are
    planted at warp, tile and column-split boundaries."""
    B, D, Dd = 1024, 128, 192
    img, txt, dino = synth(21, B, D, Dd)
    g = torch.Generator().manual_seed(5)
    r = lambda x: x.to(torch.bfloat16).to(torch.float32)
    for i in (0, 31, 32, 255, 256, 300, 511, 777, 1023):
        txt[i] = r(-img[i])                      # matched pair with cosine -1: LSE lower bound far below the rest
    for i in (1, 33, 257, 301, 640, 1022):
        txt[i] = img[i]                          # duplicates: logit = scale
        txt[(i + 7) % B] = r(torch.nn.functional.normalize(img[i] + 0.05 * torch.randn(D, generator=g), dim=-1))
    check_against_oracle(pkg, oracle, B, D, Dd, 100.0, make_args(), inputs=(img, txt, dino))
    check_against_oracle(pkg, oracle, B, D, Dd, 100.0, make_args(lambda_soft=0.0, soft_mode="none"),
                         inputs=(img, txt, dino))


def test_clip_one_pass_equals_two_pass(pkg, monkeypatch):
    """Same losses and gradients whether the text -> image statistics come from the column partials of the one-pass
    kernel or from the second (transposed) launch (DSOFT_CLIP_SYM=0)."""
    from dinosoft_b200 import loss as loss_mod

    img, txt, dino = synth(8, 1536, 512, 768)
    res = []
    for flag in ("1", "0"):
        monkeypatch.setenv("DSOFT_CLIP_SYM", flag)
        loss_mod._cuda_backend._plans.clear()     # the switch is read when a plan is created
        _, out, gi, gt, gs = run_cuda(pkg, img, txt, dino, 40.0, make_args())
        res.append((float(out["total_loss"]), float(out["classic_loss"]), gi.clone(), gt.clone(), float(gs)))
    loss_mod._cuda_backend._plans.clear()
    (t1, c1, gi1, gt1, s1), (t0, c0, gi0, gt0, s0) = res
    assert c1 == pytest.approx(c0, rel=2e-6) and t1 == pytest.approx(t0, rel=2e-6)
    assert rel_err(gi1, gi0)[0] < 1e-4 and rel_err(gt1, gt0)[0] < 1e-4
    assert s1 == pytest.approx(s0, rel=1e-5, abs=1e-8)
