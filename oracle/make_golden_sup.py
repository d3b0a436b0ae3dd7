#!/usr/bin/env python
"""Generate tests/golden/reference_golden_sup.npz from the UNMODIFIED reference (TEST INFRASTRUCTURE).

Supervised-branch fixtures: ``CrossEntropyLoss2d`` (generalframework/loss/loss.py:12-25)
exactly as ``CoTrainer._train_loop`` calls it (``criterions['sup'](pred, gt.squeeze(1))``,
trainer/cotraining_totalloss.py:211) plus the ``DiceMeter.add(pred, gt)`` of the next line, with the
autograd gradients of the reference, in fp32 and fp64.  Run in the build container only:

    python oracle/make_golden_sup.py
"""
import os
import sys
import warnings

import numpy as np
import torch

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
from generalframework.loss import CrossEntropyLoss2d, get_loss_fn  # noqa: E402
from generalframework.metrics import DiceMeter  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
SEED = 1234


def main():
    G = {}
    shapes = {2: (3, 16, 24), 4: (4, 16, 16), 19: (2, 12, 12), 5: (2, 9, 7)}
    for C, (B, H, W) in shapes.items():
        g = torch.Generator(); g.manual_seed(SEED + 500 + C)
        x = 3 * torch.randn(B, C, H, W, generator=g)
        gt = torch.randint(0, C, (B, 1, H, W), generator=g)
        gt_ign = gt.clone(); gt_ign[torch.rand(B, 1, H, W, generator=g) < 0.1] = 255
        cw = (0.25 + torch.rand(C, generator=g)).tolist()
        gout = torch.randn(B, H, W, generator=g)
        cases = {
            "plain": dict(gt=gt, weight=None, kw={}),
            "ignore": dict(gt=gt_ign, weight=None, kw={}),
            "weighted": dict(gt=gt, weight=cw, kw={}),
            "weighted_ignore": dict(gt=gt_ign, weight=cw, kw={}),
            "sum": dict(gt=gt_ign, weight=cw, kw=dict(size_average=False)),
            "none": dict(gt=gt_ign, weight=cw, kw=dict(reduce=False)),
            "confident": dict(gt=gt, weight=None, kw={}, scale=10.0),
        }
        for name, cs in cases.items():
            key = f"ce_C{C}_{name}"
            xx = x * (cs.get("scale", 3.0) / 3.0)   # x is 3*randn; "confident" is 10*randn
            G[key + "/x"] = xx.numpy(); G[key + "/gt"] = cs["gt"].numpy()
            G[key + "/weight"] = np.asarray(cs["weight"] if cs["weight"] is not None else [], dtype=np.float32)
            G[key + "/gout"] = gout.numpy()
            for suf, dt in (("32", torch.float32), ("64", torch.float64)):
                z = xx.to(dt).clone().requires_grad_()
                if name == "plain":
                    crit = get_loss_fn("cross_entropy")       # the registry path (loss/__init__.py:6-16)
                else:
                    crit = CrossEntropyLoss2d(weight=cs["weight"], **cs["kw"])
                if dt == torch.float64 and cs["weight"] is not None:
                    crit.loss.weight = crit.loss.weight.double()
                out = crit(z, cs["gt"].squeeze(1))
                if out.dim() == 0:
                    (0.37 * out).backward()
                else:
                    out.backward(gout.to(dt))
                G[key + "/ref_loss" + suf] = out.detach().numpy()
                G[key + "/ref_gz" + suf] = z.grad.numpy()
            if name in ("plain", "confident"):  # the meter of the next line of the loop, same tensors
                m = DiceMeter(method="2d", C=C); m.add(xx, cs["gt"])
                G[key + "/ref_dice2d"] = m.log.numpy()
        # all pixels ignored: mean = 0/0 = NaN, gradients 0
        z = x.clone().requires_grad_()
        out = CrossEntropyLoss2d()(z, torch.full((B, H, W), 255, dtype=torch.long))
        out.backward()
        G[f"ce_C{C}_all_ignored/ref_loss32"] = out.detach().numpy()
        G[f"ce_C{C}_all_ignored/ref_gz32"] = z.grad.numpy()
        G[f"ce_C{C}_all_ignored/x"] = x.numpy()
    path = os.path.join(OUT, "reference_golden_sup.npz")
    np.savez_compressed(path, **G)
    print(f"wrote {path}: {len(G)} arrays, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
