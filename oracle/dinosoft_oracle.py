"""CPU oracle for the DINO-Soft loss hot path.  TEST INFRASTRUCTURE - NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module, and only as the checker / the timed CPU baseline.  The product path
(``refining-clip-via-dinov2-representations_b200``) never imports it and has no CPU fallback.

What it is: a plain restatement, in torch CPU tensor algebra (fp32 or fp64), of the algorithm in the
reference's ``src/open_clip/loss.py`` (paths below are relative to /root/reference):

    gather_features ........................... src/open_clip/loss.py:23-81
    compute_student_tau ....................... src/open_clip/loss.py:166-175
    ClipLossWithDINOEnhancements.get_logits ... src/open_clip/loss.py:254-274
    classic CE ................................ src/open_clip/loss.py:313-319
    projection + normalise .................... src/open_clip/loss.py:322-347
    KL-teacher soft term ...................... src/open_clip/loss.py:356-384
    text-text KL term ......................... src/open_clip/loss.py:387-397
    total / return dict ....................... src/open_clip/loss.py:473-477, 598-607

Pinning: the reference has no tests or golden vectors for this path, so the oracle
is pinned against the reference itself, executed in the build container by ``oracle/gen_golden.py``
(outputs committed under ``tests/golden/``; ``tests/test_oracle_golden.py`` checks every fixture).
The arithmetic is all third-party (PyTorch ATen: mm, log_softmax, softmax, kl_div, normalize,
cross_entropy); it is written out explicitly here (log-sum-exp, soft-max, KL sums) instead of calling
the fused ``torch.nn.functional`` ops so that the oracle is an independent statement of the math.

World-size > 1 semantics are modelled without any process group: the caller hands over the *global*
tensors (all ranks concatenated rank-major, which is what ``gather_features`` builds) and a rank; the
oracle evaluates that rank's loss.  Gradients follow the reference's convention: with
``gather_with_grad=True`` the gradient a rank ends up with for its local features is the derivative of
the SUM of all ranks' losses (all-gather backward = reduce-scatter, loss.py:59-64).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, Optional

import torch

__all__ = [
    "OracleConfig",
    "compute_student_tau",
    "round_bf16",
    "mlp_head",
    "rank_loss",
    "weighted_ce",
    "loss_and_grads",
]


@dataclass
class OracleConfig:
    """The knobs ``forward`` reads from ``args`` (loss.py:303-310, 353-354, 368, 387-391, 474)."""

    lambda_original: float = 1.0
    lambda_soft: float = 0.0
    soft_mode: str = "none"
    teacher_temp: float = 0.15
    soft_dino_to_text: bool = False
    text_lambda: float = 0.2
    text_student_temp: float = 0.05
    # sharding
    world_size: int = 1
    local_loss: bool = False
    gather_with_grad: bool = False
    soft_scope: str = "global"  # "global": soft terms over all B columns; "local": reference at W>1
    # residual projection (loss.py:331-343): only active when the head keeps the width (Dd == D)
    residual_projection: bool = False
    residual_alpha: Optional[float] = None
    # denominator-modulated ("weighted") CE branch (loss.py:416-471); single-rank only in the reference
    lambda_weighted: float = 0.0
    rho: float = 0.1
    c_clip: float = 1.0
    weight_text_symmetry: bool = False
    # operand rounding of the B200 path (the reference has none): student operand rounded to bf16
    round_student_bf16: bool = False
    extra: Dict[str, object] = field(default_factory=dict)

    @property
    def soft_enabled(self) -> bool:
        return self.lambda_soft > 0.0 and self.soft_mode == "kl_teacher"

    @property
    def text_enabled(self) -> bool:
        return self.soft_enabled and bool(self.soft_dino_to_text) and float(self.text_lambda) > 0.0


def compute_student_tau(logit_scale: torch.Tensor) -> torch.Tensor:
    """loss.py:166-175 - student temperature from the (detached) logit scale."""
    val = logit_scale.detach()
    mult = torch.where(val > 10, val, val.exp())
    mult = torch.clamp(mult, max=100)
    return (1.0 / mult).clamp(min=0.008, max=0.02)


class _RoundBF16STE(torch.autograd.Function):
    """Round to bf16 with a straight-through gradient (models the operand rounding of the B200 path)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g


def round_bf16(x: torch.Tensor) -> torch.Tensor:
    return _RoundBF16STE.apply(x)


def _l2_normalize(x: torch.Tensor, eps: float = 1e-12) -> torch.Tensor:
    """F.normalize(x, dim=-1): x / max(||x||_2, eps)  (loss.py:345-347, 358-359, 392)."""
    n = x.pow(2).sum(dim=-1, keepdim=True).sqrt().clamp_min(eps)
    return x / n


def _row_lse(x: torch.Tensor) -> torch.Tensor:
    m = x.max(dim=1, keepdim=True).values
    m = torch.where(torch.isfinite(m), m, torch.zeros_like(m))
    return (m + (x - m).exp().sum(dim=1, keepdim=True).log()).squeeze(1)


def _cross_entropy_mean(logits: torch.Tensor, labels: torch.Tensor) -> torch.Tensor:
    """F.cross_entropy(logits, labels) with mean reduction (loss.py:317-319)."""
    picked = logits.gather(1, labels.view(-1, 1)).squeeze(1)
    return (_row_lse(logits) - picked).mean()


def _kl_batchmean(student_logits: torch.Tensor, q: torch.Tensor) -> torch.Tensor:
    """F.kl_div(log_softmax(student), q, reduction='batchmean') (loss.py:382-383, 395-396).

    kl_div's pointwise term is xlogy(q, q) - q * log_p, i.e. 0 where q == 0; 'batchmean' divides the
    total by the number of rows."""
    log_p = student_logits - _row_lse(student_logits).unsqueeze(1)
    qlogq = torch.where(q > 0, q * q.clamp_min(torch.finfo(q.dtype).tiny).log(), torch.zeros_like(q))
    return (qlogq - q * log_p).sum() / student_logits.shape[0]


def _modulated_logits(logits: torch.Tensor, r: torch.Tensor, rho: float, c_clip: float) -> torch.Tensor:
    """One direction of the DINO-guided logit modulation (loss.py:432-446 / 450-463).

    ``r`` is the DINO dissimilarity in the orientation of ``logits`` with a zero diagonal.  The offsets are
    p-centred with the UNMODIFIED row soft-max (gradient flows through it), clamped to +-c_clip and scaled by
    beta = rho * median_row(std_row(logits)) / c_clip, a host float (``.item()``: no gradient)."""
    p_base = torch.softmax(logits, dim=1)
    centred = (r - (p_base * r).sum(dim=1, keepdim=True)).clamp(min=-c_clip, max=c_clip)
    with torch.no_grad():
        # torch.std: unbiased; torch.median: the lower of the two middle values for an even row count
        sigma = torch.median(logits.float().std(dim=1)).clamp(min=1e-6)
    beta = float(rho * sigma / c_clip)
    n = logits.shape[0]
    idx = torch.arange(n)
    delta = beta * centred
    delta = delta.clone()
    delta[idx, idx] = 0.0  # masked_fill(eye, 0): the positive pair's logit is never shifted
    return logits + delta, beta


def weighted_ce(logits_it: torch.Tensor, logits_ti: torch.Tensor, dino: torch.Tensor, labels: torch.Tensor,
                cfg: "OracleConfig") -> Dict[str, object]:
    """Denominator-modulated CLIP CE (loss.py:416-471): 0.5 * (CE(L_it + Delta_it) + CE(L_ti [+ Delta_ti])).

    Needs square [b, b] logits: the reference builds ``r`` and the diagonal mask from the LOCAL batch
    (loss.py:423-429), so the branch only runs with one rank."""
    b = dino.shape[0]
    if logits_it.shape != (b, b):
        raise RuntimeError(
            f"The size of tensor a ({logits_it.shape[1]}) must match the size of tensor b ({b}) at non-singleton "
            "dimension 1")
    with torch.no_grad():
        dn = _l2_normalize(dino)
        r = 1.0 - (dn @ dn.T).clamp(-1, 1)
        idx = torch.arange(b)
        r[idx, idx] = 0.0
    it_tilde, beta_img = _modulated_logits(logits_it, r, float(cfg.rho), float(cfg.c_clip))
    beta_txt = 0.0
    if cfg.weight_text_symmetry:
        ti_tilde, beta_txt = _modulated_logits(logits_ti, r.T, float(cfg.rho), float(cfg.c_clip))
    else:
        ti_tilde = logits_ti
    loss = 0.5 * (_cross_entropy_mean(it_tilde, labels) + _cross_entropy_mean(ti_tilde, labels))
    return {"loss": loss, "beta_img": beta_img, "beta_txt": beta_txt}


def mlp_head(x: torch.Tensor, params: Dict[str, torch.Tensor], projection_type: str = "mlp") -> torch.Tensor:
    """The lightweight projection head (loss.py:214-238): Linear or Linear-ReLU-Linear[-LayerNorm]."""
    if projection_type == "linear":
        return x @ params["w0"].T + params["b0"]
    if projection_type != "mlp":
        raise ValueError(f"Unknown projection_type: {projection_type}")
    h = torch.relu(x @ params["w0"].T + params["b0"])
    y = h @ params["w1"].T + params["b1"]
    if "ln_w" in params:
        mu = y.mean(dim=-1, keepdim=True)
        var = (y - mu).pow(2).mean(dim=-1, keepdim=True)
        y = (y - mu) / (var + 1e-5).sqrt() * params["ln_w"] + params["ln_b"]
    return y


def rank_loss(
    image_all: torch.Tensor,
    text_all: torch.Tensor,
    logit_scale: torch.Tensor,
    dino_all: Optional[torch.Tensor],
    student_all: Optional[torch.Tensor],
    cfg: OracleConfig,
    rank: int = 0,
) -> Dict[str, torch.Tensor]:
    """Loss of one rank, restating ``ClipLossWithDINOEnhancements.forward`` (loss.py:292-477).

    ``*_all`` are the global ``[B, .]`` tensors (rank-major concatenation).  ``student_all`` is the raw
    output of the projection head for every row (``None``: the student is the image feature itself,
    loss.py:347)."""
    W = cfg.world_size
    B = image_all.shape[0]
    assert B % W == 0
    b = B // W
    rows = slice(rank * b, (rank + 1) * b)
    img, txt = image_all[rows], text_all[rows]

    # ---- get_logits (loss.py:254-274) + get_ground_truth (loss.py:241-252)
    if W > 1:
        if not cfg.local_loss:
            # loss.py:269 builds [B, B] logits but labels come from the local b (loss.py:302, 314)
            raise ValueError(f"Expected input batch_size ({B}) to match target batch_size ({b}).")
        cols_i, cols_t = image_all, text_all
        if not cfg.gather_with_grad:  # loss.py:65-79: gathered copies carry no gradient
            cols_i, cols_t = image_all.detach(), text_all.detach()
        logits_it = logit_scale * (img @ cols_t.T)
        logits_ti = logit_scale * (txt @ cols_i.T)
        labels = torch.arange(b) + b * rank
    else:
        logits_it = logit_scale * (img @ txt.T)
        logits_ti = logit_scale * (txt @ img.T)
        labels = torch.arange(b)
    classic = 0.5 * (_cross_entropy_mean(logits_it, labels) + _cross_entropy_mean(logits_ti, labels))

    out = {"classic_loss": classic}
    soft_img = torch.zeros((), dtype=image_all.dtype)
    soft_txt = torch.zeros((), dtype=image_all.dtype)
    if cfg.soft_enabled and dino_all is not None:
        # column scope of the B x B soft matrices
        if W > 1 and cfg.soft_scope == "local":
            cols = rows
            diag0 = 0
        else:
            cols = slice(0, B)
            diag0 = rank * b
        stu = image_all if student_all is None else student_all
        if cfg.round_student_bf16 and student_all is not None:
            stu = round_bf16(stu)
        z_all = _l2_normalize(_l2_normalize(stu))  # normalised twice: loss.py:345/347 then 358
        d_all = _l2_normalize(dino_all)
        z_cols, d_cols = z_all[cols], d_all[cols]
        if W > 1 and not cfg.gather_with_grad and cfg.soft_scope != "local":
            z_cols = z_cols.detach()
        tau_s = compute_student_tau(logit_scale)
        tau_t = float(cfg.teacher_temp)
        s_student = (z_all[rows] @ z_cols.T) / tau_s  # loss.py:372
        s_teacher = (d_all[rows] @ d_cols.T) / tau_t  # loss.py:373
        eye = torch.zeros_like(s_teacher, dtype=torch.bool)
        idx = torch.arange(b)
        eye[idx, idx + diag0] = True
        s_teacher = s_teacher.masked_fill(eye, float("-inf"))  # loss.py:376-377
        with torch.no_grad():  # loss.py:379-380
            q = (s_teacher - _row_lse(s_teacher).unsqueeze(1)).exp()
        soft_img = _kl_batchmean(s_student, q)  # loss.py:382-383
        if cfg.text_enabled:  # loss.py:387-397
            t_all = _l2_normalize(text_all)
            t_cols = t_all[cols]
            if W > 1 and not cfg.gather_with_grad and cfg.soft_scope != "local":
                t_cols = t_cols.detach()
            s_tt = (t_all[rows] @ t_cols.T) / float(cfg.text_student_temp)
            soft_txt = _kl_batchmean(s_tt, q)
    soft = soft_img + (float(cfg.text_lambda) * soft_txt if cfg.text_enabled else 0.0)
    weighted = torch.zeros((), dtype=image_all.dtype)
    if float(cfg.lambda_weighted) > 0.0 and dino_all is not None and b > 1:  # loss.py:422
        wres = weighted_ce(logits_it, logits_ti, dino_all[rows], labels, cfg)
        weighted = wres["loss"]
        out.update(beta_img=wres["beta_img"], beta_txt=wres["beta_txt"])
    total = (float(cfg.lambda_original) * classic + float(cfg.lambda_soft) * soft
             + float(cfg.lambda_weighted) * weighted)  # loss.py:473-477
    out.update(soft_img=soft_img, soft_txt=soft_txt, soft_loss=soft, weighted_loss=weighted, total_loss=total)
    return out


def loss_and_grads(
    image_all: torch.Tensor,
    text_all: torch.Tensor,
    logit_scale: float,
    dino_all: Optional[torch.Tensor],
    cfg: OracleConfig,
    proj_params: Optional[Dict[str, torch.Tensor]] = None,
    projection_type: str = "mlp",
    dtype: torch.dtype = torch.float64,
    student_values: Optional[torch.Tensor] = None,
) -> Dict[str, object]:
    """Evaluate every rank's loss and the gradients the reference's backward would deliver.

    Returns, per rank r: the loss terms, ``d_image[r]``/``d_text[r]`` ([b, D], gradient of this rank's
    local features), ``d_logit_scale[r]`` and (with a projection head) ``d_student[r]`` (gradient w.r.t.
    the raw head output) and ``d_proj[r]`` (head parameters).  With ``gather_with_grad=True`` feature
    gradients are d(sum_r loss_r)/d(local features); otherwise only the rank's own loss contributes.

    ``student_values`` ([B, Dp], optional): the student operand the implementation under test really used (its
    own head output after bf16 rounding).  The head stays in the graph, but its VALUE is replaced by these
    numbers, so that a head output that lands on the other side of a bf16 rounding boundary in fp32 than in
    fp64 does not show up as a gradient difference."""
    W = cfg.world_size
    B = image_all.shape[0]
    b = B // W
    res = {"ranks": []}

    def leafs():
        im = image_all.detach().to(dtype).clone().requires_grad_(True)
        tx = text_all.detach().to(dtype).clone().requires_grad_(True)
        sc = torch.tensor(float(logit_scale), dtype=dtype, requires_grad=True)
        dn = None if dino_all is None else dino_all.detach().to(dtype)
        pp = None
        if proj_params is not None:
            pp = {k: v.detach().to(dtype).clone().requires_grad_(True) for k, v in proj_params.items()}
        return im, tx, sc, dn, pp

    for r in range(W):
        im, tx, sc, dn, pp = leafs()
        student = None
        if pp is not None and dn is not None:
            student = mlp_head(im, pp, projection_type)
            if cfg.residual_projection and student.shape == im.shape:  # loss.py:331-343
                if cfg.residual_alpha is None:
                    student = im + student
                else:
                    student = cfg.residual_alpha * im + (1 - cfg.residual_alpha) * student
            if student_values is not None:
                student = student + (student_values.to(dtype) - student).detach()
            student.retain_grad()
        # this rank's own loss terms
        own = rank_loss(im, tx, sc, dn, student, cfg, rank=r)
        if cfg.gather_with_grad and W > 1:
            # features: derivative of the sum of all ranks' losses (all-gather backward sums over ranks);
            # logit_scale: own loss only (DDP averages parameter gradients later)
            others = [rank_loss(im, tx, sc, dn, student, cfg, rank=k)["total_loss"] for k in range(W) if k != r]
            g_sc, = torch.autograd.grad(own["total_loss"], sc, retain_graph=True)
            total = own["total_loss"] + sum(others)
            total.backward()
        else:
            own["total_loss"].backward()
            g_sc = sc.grad
        rows = slice(r * b, (r + 1) * b)
        entry = {k: float(v.detach()) if torch.is_tensor(v) else float(v) for k, v in own.items()}
        entry["d_image"] = im.grad[rows].detach().clone()
        entry["d_text"] = tx.grad[rows].detach().clone()
        entry["d_logit_scale"] = float(g_sc)
        if student is not None and student.grad is not None:
            entry["d_student"] = student.grad[rows].detach().clone()
            # head parameters: this rank back-propagates d_student through its own head on its local rows
            # (the residual mix scales the head's share by 1 or 1 - alpha)
            local_out = mlp_head(im.detach()[rows], pp, projection_type)
            if cfg.residual_projection and local_out.shape[1] == im.shape[1] and cfg.residual_alpha is not None:
                local_out = (1 - cfg.residual_alpha) * local_out
            g_pp = torch.autograd.grad(local_out, list(pp.values()), grad_outputs=entry["d_student"],
                                       allow_unused=True)
            entry["d_proj"] = {k: (None if g is None else g.detach().clone()) for k, g in zip(pp.keys(), g_pp)}
        res["ranks"].append(entry)
    return res


def algorithmic_flops(B: int, D: int, Dp: int, Dd: int, text: bool, soft: bool = True) -> float:
    """F_alg = 2 B^2 (3D + 2Dp + Dd [+ 2D])  (Dp = student width)."""
    f = 2.0 * B * B * 3 * D
    if soft:
        f += 2.0 * B * B * (2 * Dp + Dd)
    if text:
        f += 2.0 * B * B * 2 * D
    return f
