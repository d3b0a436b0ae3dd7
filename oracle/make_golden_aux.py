#!/usr/bin/env python
"""Generate tests/golden/reference_golden_aux.npz from the UNMODIFIED reference (TEST INFRASTRUCTURE).

Class-map / one-hot helpers (generalframework/utils/utils.py:73-80,154-235), the functional Dice of the supervised
baseline (dice_coef / dice_batch), the evaluation script's ensemble voting (``Ensembleway``, Summary.py:88-120 --
Summary.py is an argparse script, so the class definition is lifted out of its source with ``ast`` at run time
and executed against the reference's own helpers; nothing is copied into this repository) and the kappa meters
(generalframework/metrics/kappa.py, sklearn).  Run in the build container only:

    python oracle/make_golden_aux.py
"""
import ast
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
from generalframework.utils import utils as U  # noqa: E402
from generalframework.metrics.kappa import Kappa2Annotator, KappaMetrics  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
SEED = 1234


def load_ensembleway(num_classes_holder):
    """exec the reference's ``class Ensembleway`` (Summary.py) in a namespace holding its free names."""
    path = os.path.join(ref_shim.REFERENCE_ROOT, "Summary.py")
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == "Ensembleway")
    from typing import List
    ns = {"torch": torch, "np": np, "List": List, "Tensor": torch.Tensor, "class2one_hot": U.class2one_hot,
          "config": num_classes_holder}
    exec(compile(ast.Module(body=[node], type_ignores=[]), path, "exec"), ns)
    return ns["Ensembleway"]


def main():
    G = {}
    holder = {"Arch": {"num_classes": 4}}
    Ens = load_ensembleway(holder)
    shapes = {2: (3, 16, 24), 4: (4, 16, 16), 19: (2, 12, 12), 5: (2, 9, 7)}
    for C, (B, H, W) in shapes.items():
        g = torch.Generator(); g.manual_seed(SEED + 900 + C)
        x = 3 * torch.randn(B, C, H, W, generator=g)
        # exact ties (first index must win) and a repeated maximum
        x[0, :, 0, 0] = 1.0
        x[0, :, 0, 1] = torch.tensor([0.5] + [2.0] * (C - 1))
        p = torch.softmax(x, 1)
        gt = torch.randint(0, C, (B, H, W), generator=g)
        key = f"aux_C{C}"
        G[key + "/x"] = x.numpy(); G[key + "/p"] = p.numpy(); G[key + "/gt"] = gt.numpy()
        G[key + "/pred2class_x"] = U.pred2class(x).numpy()
        G[key + "/probs2class_p"] = U.probs2class(p).numpy()
        G[key + "/class2one_hot_gt"] = U.class2one_hot(gt, C).numpy()
        G[key + "/probs2one_hot_p"] = U.probs2one_hot(p).numpy()
        G[key + "/predlogit2one_hot_x"] = U.predlogit2one_hot(x).numpy()
        lab_oh, pred_oh = U.class2one_hot(gt, C), U.probs2one_hot(p)
        G[key + "/dice_coef"] = U.dice_coef(lab_oh, pred_oh).numpy()
        G[key + "/dice_batch"] = U.dice_batch(lab_oh, pred_oh).numpy()
        G[key + "/intersection"] = U.intersection(lab_oh, pred_oh).numpy()
        G[key + "/one_hot_true"] = np.asarray(bool(U.one_hot(lab_oh)))
        broken = lab_oh.clone(); broken[0, 0, 1, 1] = 1 - broken[0, 0, 1, 1]
        G[key + "/one_hot_broken"] = np.asarray(bool(U.one_hot(broken)))
        two = lab_oh.clone(); two[0, :, 2, 2] = 0; two[0, 0, 2, 2] = 2
        G[key + "/one_hot_value2"] = np.asarray(bool(U.one_hot(two)))
        # ---- ensemble voting, K = 2, 3, 4 views
        for K in (2, 3, 4):
            views = [torch.softmax(x + 1.5 * torch.randn(B, C, H, W, generator=g), 1) for _ in range(K)]
            vk = f"{key}_K{K}"
            for k, v in enumerate(views):
                G[f"{vk}/view{k}"] = v.numpy()
            soft = Ens("soft")(views)
            G[vk + "/soft"] = soft.numpy()
            G[vk + "/soft_class"] = U.pred2class(soft).numpy()
            holder["Arch"]["num_classes"] = C
            hard = torch.cat([Ens("hard")([v[b:b + 1] for v in views]) for b in range(B)], 0)  # the reference votes at B = 1
            G[vk + "/hard"] = hard.numpy()
            # ---- kappa: every model against the voted target, considered classes = all but background
            preds = [v.max(1)[1] for v in views]
            target = soft.max(1)[1]
            considered = list(range(1, C)) if C > 2 else [0, 1]
            km = KappaMetrics(); km.add(predicts=preds, target=target, considered_classes=considered)
            G[vk + "/kappa_considered"] = np.asarray(considered)
            G[vk + "/kappa_vs_vote"] = np.asarray(km.kappa[0], dtype=np.float64)
            k2 = Kappa2Annotator(); k2.add(preds[0], preds[1], gt=gt, considered_classes=considered)
            G[vk + "/kappa2"] = np.asarray(k2.kappa[0], dtype=np.float64)
            k2n = Kappa2Annotator(); k2n.add(preds[0], preds[1], gt=gt, considered_classes=None)
            G[vk + "/kappa2_all"] = np.asarray(k2n.kappa[0], dtype=np.float64)
    path = os.path.join(OUT, "reference_golden_aux.npz")
    np.savez_compressed(path, **G)
    print(f"wrote {path}: {len(G)} arrays, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
