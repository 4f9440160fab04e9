"""Golden fixture for the CyCLIP loss, produced by EXECUTING the reference's `CyCLIPLoss` (src/open_clip/loss.py:
813-905; loaded by file path, the file only needs torch) on seeded inputs in fp32 and fp64.

    python oracle/gen_golden_cyclip.py       # build container only (/root/reference)
"""
import importlib.util
import os

import numpy as np
import torch

REF = "/root/reference/src/open_clip/loss.py"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "cyclip_b96.npz")


def main():
    spec = importlib.util.spec_from_file_location("ref_open_clip_loss", REF)
    ref = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(ref)
    g = torch.Generator().manual_seed(11)
    B, D = 96, 64
    r = lambda x: x.to(torch.bfloat16).float()
    cid = torch.randint(0, 12, (B,), generator=g)
    cent = torch.randn(12, D, generator=g)
    img = r(torch.nn.functional.normalize(cent[cid] + 0.5 * torch.randn(B, D, generator=g), dim=-1))
    txt = r(torch.nn.functional.normalize(cent[cid] @ torch.randn(D, D, generator=g) / D ** 0.5
                                          + 0.5 * torch.randn(B, D, generator=g), dim=-1))
    out = {"image": img.numpy(), "text": txt.numpy(), "scale": np.float64(25.0), "lambda_inmodal": np.float64(0.25),
           "lambda_crossmodal": np.float64(0.5)}

    class Float64CyCLIP(ref.CyCLIPLoss):
        @staticmethod
        def _cosine_normalize(x):  # the reference forces fp32 here (loss.py:862-864); the fp64 pin keeps fp64
            return torch.nn.functional.normalize(x, dim=-1)

    for tag, dt, cls in (("f32", torch.float32, ref.CyCLIPLoss), ("f64", torch.float64, Float64CyCLIP)):
        loss = cls(lambda_inmodal=0.25, lambda_crossmodal=0.5)
        im = img.detach().clone().to(dt).requires_grad_(True)
        tx = txt.detach().clone().to(dt).requires_grad_(True)
        sc = torch.tensor(25.0, dtype=dt, requires_grad=True)
        res = loss(im, tx, sc, output_dict=True)
        res["total_loss"].backward()
        for k in ("total_loss", "clip_loss", "inmodal_cyclic", "crossmodal_cyclic"):
            out[f"{tag}_{k}"] = np.float64(float(res[k].detach()))
        out[f"{tag}_d_image"] = im.grad.double().numpy()
        out[f"{tag}_d_text"] = tx.grad.double().numpy()
        out[f"{tag}_d_logit_scale"] = np.float64(float(sc.grad))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, {k: float(v) for k, v in out.items() if k.endswith("_loss") or k.endswith("cyclic")})


if __name__ == "__main__":
    main()
