"""Helpers shared by the GPU parity tests (synthetic inputs per SURVEY.md 8(d), comparisons)."""
import types

import torch


def synth(seed, B, D, Dd, clustered=True, device="cpu"):
    """Clustered embeddings; every value is bf16-representable so kernel and oracle see identical inputs."""
    g = torch.Generator().manual_seed(seed)
    K = max(B // 16, 2)
    cid = torch.randint(0, K, (B,), generator=g)

    def make(d, scale=1.0):
        cent = torch.randn(K, d, generator=g)
        x = (cent[cid] if clustered else 0.0) + (0.5 if clustered else 1.0) * torch.randn(B, d, generator=g)
        return x * scale

    img = torch.nn.functional.normalize(make(D), dim=-1)
    txt = torch.nn.functional.normalize(make(D), dim=-1)
    dino = make(Dd, 3.0)
    r = lambda x: x.to(torch.bfloat16).to(torch.float32).to(device)
    return r(img), r(txt), r(dino)


def synth_aligned(seed, B, D, Dd, noise=0.3, device="cpu"):
    """Near-converged regime: text_i ~ image_i, so p_ii -> 1 and the CE gradient is a small residual."""
    g = torch.Generator().manual_seed(seed)
    r = lambda x: x.to(torch.bfloat16).to(torch.float32).to(device)
    img = torch.nn.functional.normalize(torch.randn(B, D, generator=g), dim=-1)
    txt = torch.nn.functional.normalize(img + noise * torch.randn(B, D, generator=g) / D ** 0.5 * 4, dim=-1)
    dino = 3.0 * torch.randn(B, Dd, generator=g)
    return r(img), r(txt), r(dino)


def make_args(**kw):
    base = dict(use_projection=False, projection_type="mlp", lambda_soft=0.5, soft_mode="kl_teacher",
                soft_dino_to_text=True, text_lambda=0.5, text_student_temp=0.02, teacher_temp=0.15,
                lambda_original=1.0, lambda_weighted=0.0, rho=0.1, c_clip=1.0, weight_text_symmetry=False)
    base.update(kw)
    return types.SimpleNamespace(**base)


def oracle_cfg(oracle, args, **kw):
    return oracle.OracleConfig(
        lambda_original=args.lambda_original, lambda_soft=args.lambda_soft, soft_mode=args.soft_mode,
        teacher_temp=args.teacher_temp, soft_dino_to_text=args.soft_dino_to_text, text_lambda=args.text_lambda,
        text_student_temp=args.text_student_temp, lambda_weighted=getattr(args, "lambda_weighted", 0.0),
        rho=getattr(args, "rho", 0.1), c_clip=getattr(args, "c_clip", 1.0),
        weight_text_symmetry=getattr(args, "weight_text_symmetry", False), **kw)


def head_params_of(module, projection_type, layernorm=False):
    p = {}
    if projection_type == "linear":
        p["w0"], p["b0"] = module.weight, module.bias
    else:
        p["w0"], p["b0"] = module[0].weight, module[0].bias
        p["w1"], p["b1"] = module[2].weight, module[2].bias
        if layernorm:
            p["ln_w"], p["ln_b"] = module[3].weight, module[3].bias
    return p


def rel_err(got, ref):
    got = got.detach().double().cpu()
    ref = ref.detach().double().cpu()
    linf = (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-30)
    l2 = (got - ref).norm().item() / max(ref.norm().item(), 1e-30)
    return linf, l2
