"""TEST DOUBLE (lives under tests/ on purpose): a stand-in for the CUDA backend built on the CPU oracle.

It lets the host-side logic of the loss module (packing layout, the all-gather plumbing, rank / label
handling, flag decoding, autograd wiring, loss composition) run under `gloo` on a CPU-only box.  The
product never selects it: `ClipLossWithDINOEnhancements` resolves its backend to the C-ABI CUDA backend
and raises on non-CUDA tensors; tests inject this object through the private `_backend` attribute."""
import torch

from weighted_ref import weighted_ce_branch

F_SOFT, F_TEXT, F_SOFT_LOCAL, F_ROW_ONLY = 1, 2, 4, 8
F_WEIGHTED, F_WSYM = 32, 64
DBG_ORDER = ("pc_err_img", "pc_err_txt", "diag_max_img", "diag_max_txt", "delta_img_max", "delta_img_mean",
             "delta_img_std", "delta_txt_max", "delta_txt_mean", "delta_txt_std", "l1_prob_shift_img",
             "l1_prob_shift_txt", "corr_rhat_dprob_img", "corr_rhat_dprob_txt", "ce_img_base", "ce_txt_base",
             "ce_img_mod", "ce_txt_mod", "pos_frac_img", "neg_frac_img", "pos_frac_txt", "neg_frac_txt", "beta_img",
             "beta_txt")


class _Plan:
    pass


class OracleBackend:
    name = "oracle-test-double"

    def __init__(self, oracle):
        self.oracle = oracle
        self.calls = []

    def plan(self, shape, device=None):
        p = _Plan()
        p.shape = shape
        p.soft = bool(shape.flags & F_SOFT)
        p.weighted = bool(shape.flags & F_WEIGHTED)
        p.proj = p.soft and shape.Dp > 0
        p.offI, p.offT = 0, shape.D
        p.offZ = 2 * shape.D if p.proj else 0
        p.offD = 2 * shape.D + (shape.Dp if p.proj else 0)
        p.row_elems = p.offD + (shape.Dd if (p.soft or p.weighted) else 0)
        p.dino_col = p.offD
        p.state_numel = 8
        p.scratch_numel = 8
        p.fwd_scratch_numel = 8
        p.flops, p.launches_fwd, p.launches_bwd = 0.0, 0, 0
        return p

    def pack(self, plan, image, text, student, dino, gathered):
        s = plan.shape
        rows = slice(s.rank * s.b, (s.rank + 1) * s.b)
        gathered[rows, plan.offI:plan.offI + s.D] = image.to(torch.bfloat16)
        gathered[rows, plan.offT:plan.offT + s.D] = text.to(torch.bfloat16)
        if plan.proj:
            gathered[rows, plan.offZ:plan.offZ + s.Dp] = student.to(torch.bfloat16)
        if plan.soft or plan.weighted:
            gathered[rows, plan.offD:plan.offD + s.Dd] = dino.to(torch.bfloat16)
        self.calls.append("pack")

    def _decode(self, plan, gathered):
        s = plan.shape
        g = gathered.to(torch.float64)
        img = g[:, plan.offI:plan.offI + s.D]
        txt = g[:, plan.offT:plan.offT + s.D]
        stu = g[:, plan.offZ:plan.offZ + s.Dp] if plan.proj else None
        dino = g[:, plan.offD:plan.offD + s.Dd] if (plan.soft or plan.weighted) else None
        return img, txt, stu, dino

    def _cfg(self, plan):
        s = plan.shape
        return self.oracle.OracleConfig(
            lambda_original=1.0, lambda_soft=1.0 if plan.soft else 0.0,
            soft_mode="kl_teacher" if plan.soft else "none", teacher_temp=s.teacher_temp,
            soft_dino_to_text=bool(s.flags & F_TEXT), text_lambda=1.0, text_student_temp=s.text_temp or 0.05,
            world_size=s.world, local_loss=True, gather_with_grad=not (s.flags & F_ROW_ONLY),
            soft_scope="local" if (s.flags & F_SOFT_LOCAL) else "global")

    def _weighted(self, plan, img, txt, scale, dino):
        s = plan.shape
        return weighted_ce_branch(img, txt, scale, dino, float(s.rho), float(s.c_clip), bool(s.flags & F_WSYM))

    def forward(self, plan, gathered, logit_scale, lambdas, state, scratch, lse_local, losses, dbg=None):
        img, txt, stu, dino = self._decode(plan, gathered)
        out = self.oracle.rank_loss(img, txt, logit_scale.double()[0], dino if plan.soft else None, stu,
                                    self._cfg(plan), rank=plan.shape.rank)
        lo, ls, tl, lw = lambdas
        c, si, st = float(out["classic_loss"]), float(out["soft_img"]), float(out["soft_txt"])
        losses[0], losses[1], losses[2] = c, si, st
        losses[3] = si + tl * st
        losses[4] = lo * c + ls * (si + tl * st)
        losses[5] = 0.0
        if plan.weighted:
            w, d = self._weighted(plan, img, txt, logit_scale.double()[0], dino)
            losses[5] = float(w)
            losses[4] = float(losses[4]) + lw * float(w)
            if dbg is not None:
                for i, k in enumerate(DBG_ORDER):
                    dbg[i] = float(d[k])
        lse_local.zero_()
        state[0] = float(logit_scale[0])
        self.calls.append("forward")

    @torch.enable_grad()  # called from inside autograd.Function.backward, where grad mode is off
    def backward(self, plan, gathered, state, scratch, lse_all, gout, lambdas, d_image, d_text, d_student, d_scale):
        s = plan.shape
        cfg = self._cfg(plan)
        img, txt, stu, dino = (None if t is None else t.clone() for t in self._decode(plan, gathered))
        img.requires_grad_(True)
        txt.requires_grad_(True)
        if stu is not None:
            stu.requires_grad_(True)
        sc = torch.tensor(float(state[0]), dtype=torch.float64, requires_grad=True)
        lo, ls, tl, lw = lambdas
        g5 = gout.double()
        gsoft = g5[3] + ls * g5[4]
        g = torch.stack([g5[0] + lo * g5[4], g5[1] + gsoft, g5[2] + tl * gsoft, g5[5] + lw * g5[4]])

        def weighted(rank):
            o = self.oracle.rank_loss(img, txt, sc, dino if plan.soft else None, stu, cfg, rank=rank)
            t = g[0] * o["classic_loss"] + g[1] * o["soft_img"] + g[2] * o["soft_txt"]
            if plan.weighted:
                t = t + g[3] * self._weighted(plan, img, txt, sc, dino)[0]
            return t

        own = weighted(s.rank)
        gs, = torch.autograd.grad(own, sc, retain_graph=True)
        total = own
        if cfg.gather_with_grad and s.world > 1:
            for k in range(s.world):
                if k != s.rank:
                    total = total + weighted(k)
        total.backward()
        rows = slice(s.rank * s.b, (s.rank + 1) * s.b)
        d_image.copy_(img.grad[rows])
        d_text.copy_(txt.grad[rows])
        if d_student is not None:
            d_student.copy_(stu.grad[rows])
        d_scale[0] = float(gs)
        self.calls.append("backward")
