"""TEST INFRASTRUCTURE: plain-torch restatement of the reference's denominator-modulated CE branch
(src/open_clip/loss.py:416-471) and of its diagnostics (loss.py:479-595), [B, B] intermediates and all.

Round 1 shipped this as the module's implementation of the branch; the product now runs the branch inside
libdsoft.so (DSOFT_F_WEIGHTED) and this file remains as (a) the CPU stand-in used by tests/oracle_backend.py for
the host-logic tests and (b) the fp64 checker of the diagnostic scalars in tests/test_gpu_parity.py.  The loss
value itself is pinned by the reference-generated fixtures (tests/golden/w1_weighted*.npz)."""
import torch
import torch.nn.functional as F


def shift_logits(logits: torch.Tensor, dissim: torch.Tensor, rho: float, c_clip: float):
    """logits + beta * clamp(r - E_p[r]) with a zero diagonal; p = soft-max of the UNMODIFIED rows (carries
    gradient), beta = rho * median(row std) / c_clip (no gradient)."""
    p_rows = torch.softmax(logits, dim=1)
    r_hat = (dissim - (p_rows * dissim).sum(dim=1, keepdim=True)).clamp(min=-c_clip, max=c_clip)
    with torch.no_grad():
        beta = rho * torch.median(logits.std(dim=1)).clamp(min=1e-6) / c_clip
    delta = (beta * r_hat).clone()
    delta.diagonal().zero_()
    return logits + delta, delta, r_hat, p_rows, beta


def row_corr_mean(a: torch.Tensor, b: torch.Tensor, eps: float = 1e-9) -> torch.Tensor:
    a = a - a.mean(dim=1, keepdim=True)
    b = b - b.mean(dim=1, keepdim=True)
    den = a.pow(2).sum(dim=1).sqrt() * b.pow(2).sum(dim=1).sqrt() + eps
    return ((a * b).sum(dim=1) / den).mean()


def weighted_ce_branch(image_features, text_features, logit_scale, dino_features, rho, c_clip, text_sym):
    """Returns (weighted_loss, dbg).  dbg holds the reference's diagnostic keys as 0-dim tensors (formatting
    one of them, as train.py:360-364 does every 300 steps, is what synchronises)."""
    B = image_features.shape[0]
    img, txt = image_features, text_features
    logits_i = logit_scale * (img @ txt.T)
    logits_t = logits_i.T  # one rank: logits_per_text is the exact transpose (loss.py:272-273)
    labels = torch.arange(B, device=img.device)
    with torch.no_grad():
        dn = F.normalize(dino_features, dim=-1)
        dissim = 1.0 - (dn @ dn.T).clamp(-1, 1)
        dissim.diagonal().zero_()
    tilde_i, delta_i, rhat_i, p_i, beta_i = shift_logits(logits_i, dissim, rho, c_clip)
    if text_sym:
        tilde_t, delta_t, rhat_t, p_t, beta_t = shift_logits(logits_t, dissim.T, rho, c_clip)
    else:
        tilde_t, delta_t, rhat_t, p_t, beta_t = logits_t, None, None, None, None
    ce_i = F.cross_entropy(tilde_i, labels)
    ce_t = F.cross_entropy(tilde_t, labels)
    loss = 0.5 * (ce_i + ce_t)

    with torch.no_grad():
        zero = torch.zeros((), device=img.device)
        off = float(B * B - B)

        def side(delta, r_hat, p_base, tilde):
            if delta is None:
                return dict(pc=zero, dmax=zero, dmean=zero, dstd=zero, diag=zero, corr=zero, pos=zero)
            p_mod = torch.softmax(tilde, dim=1)
            d_abs = delta.abs()
            pos = ((r_hat > 0).float().sum() - (r_hat.diagonal() > 0).float().sum()) / off
            return dict(pc=(p_base * r_hat).sum(dim=1).abs().mean(), dmax=d_abs.max(), dmean=d_abs.mean(),
                        dstd=d_abs.std(), diag=r_hat.diagonal().abs().max(),
                        corr=row_corr_mean(r_hat, p_mod - p_base), pos=pos)

        si = side(delta_i, rhat_i, p_i.detach(), tilde_i.detach())
        st = side(delta_t, None if rhat_t is None else rhat_t, None if p_t is None else p_t.detach(),
                  tilde_t.detach())
        p_t_base = torch.softmax(logits_t.detach(), dim=1)
        dbg = {
            "pc_err_img": si["pc"], "pc_err_txt": st["pc"],
            "diag_max_img": si["diag"], "diag_max_txt": st["diag"],
            "delta_img_max": si["dmax"], "delta_img_mean": si["dmean"], "delta_img_std": si["dstd"],
            "delta_txt_max": st["dmax"], "delta_txt_mean": st["dmean"], "delta_txt_std": st["dstd"],
            "l1_prob_shift_img": (torch.softmax(tilde_i.detach(), dim=1) - p_i.detach()).abs().sum(dim=1).mean(),
            "l1_prob_shift_txt": (torch.softmax(tilde_t.detach(), dim=1) - p_t_base).abs().sum(dim=1).mean(),
            "corr_rhat_dprob_img": si["corr"], "corr_rhat_dprob_txt": st["corr"],
            "ce_img_base": F.cross_entropy(logits_i.detach(), labels),
            "ce_txt_base": F.cross_entropy(logits_t.detach(), labels),
            "ce_img_mod": ce_i.detach(), "ce_txt_mod": ce_t.detach(),
            "pos_frac_img": si["pos"], "neg_frac_img": 1.0 - si["pos"],
            "pos_frac_txt": st["pos"], "neg_frac_txt": (1.0 - st["pos"]) if text_sym else zero,
            "beta_img": beta_i, "beta_txt": beta_t if text_sym else zero,
            "rho": rho, "clip_c": c_clip,
        }
    return loss, dbg
