#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference (TEST INFRASTRUCTURE).

Run in the build container only (needs /root/reference):

    python oracle/make_golden.py

Every array under the ``ref_`` prefix is an output of the reference's own code
(``generalframework.loss`` / ``.metrics`` / ``.utils.AEGenerator``, imported via
oracle/ref_shim.py, torch CPU) on the seeded inputs stored beside it.  The
fixtures pin the oracle (tests/test_oracle_vs_golden.py, CPU) and the CUDA path
(tests/test_gpu_parity.py, GPU box where /root/reference does not exist).
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
from generalframework.loss import (JSD, JSD_2D, Entropy_2D, KL_div, KL_Divergence_2D,  # noqa: E402
                                   KL_Divergence_2D_Logit, get_loss_fn)
from generalframework.metrics import DiceMeter, IoU  # noqa: E402
from generalframework.utils.AEGenerator import FSGMGenerator, VATGenerator  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
SEED = 1234  # the reference's Seed (config/ACDC_config_cotraing.yaml:80)


def gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def jsd_case(zs, w, gout):
    """zs: list of fp32 logits.  Returns dict of reference outputs in fp32 and fp64."""
    out = {}
    for name, dt in (("32", torch.float32), ("64", torch.float64)):
        zz = [z.detach().to(dt).clone().requires_grad_() for z in zs]
        ps = [torch.softmax(z, 1) for z in zz]
        for p in ps:
            p.retain_grad()
        crit = get_loss_fn("jsd")  # the registry path CoTrainer uses (loss/__init__.py:6-16)
        m = crit(ps)
        (w * m.mean()).backward(retain_graph=True)
        out["ref_map" + name] = m.detach().numpy()
        out["ref_mean" + name] = np.asarray(m.mean().item())
        out["ref_gz" + name] = np.stack([z.grad.numpy() for z in zz])
        # parity-mode boundary: d(map . gout)/d probs
        for p in ps:
            p.grad = None
        for z in zz:
            z.grad = None
        (m * gout.to(dt)).sum().backward()
        out["ref_gp" + name] = np.stack([p.grad.numpy() for p in ps])
        out["ref_probs" + name] = np.stack([p.detach().numpy() for p in ps])
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    G = {}

    # ---------------- JSD (a1-a4) ----------------
    shapes = {(2, 4): (2, 16, 16), (3, 4): (2, 16, 16), (2, 2): (2, 16, 24), (2, 19): (1, 8, 12),
              (4, 19): (1, 8, 8), (4, 3): (2, 5, 7)}
    for (K, C), (B, H, W) in shapes.items():
        g = gen(SEED + 17 * K + C)
        base = [3 * torch.randn(B, C, H, W, generator=g) for _ in range(K)]
        variants = {
            "spread": base,
            "confident": [10 * torch.randn(B, C, H, W, generator=g) for _ in range(K)],
            "agree": [base[0] + 0.1 * torch.randn(B, C, H, W, generator=g) for _ in range(K)],
            "identical": [base[0].clone() for _ in range(K)],
            "saturated": [80 * torch.sign(torch.randn(B, C, H, W, generator=g)) for _ in range(K)],
        }
        gout = torch.randn(B, H, W, generator=g)
        for vname, zs in variants.items():
            key = f"jsd_K{K}_C{C}_{vname}"
            w = 0.37
            G[key + "/z"] = np.stack([z.numpy() for z in zs])
            G[key + "/w"] = np.asarray(w, dtype=np.float64)
            G[key + "/gout"] = gout.numpy()
            for k, v in jsd_case(zs, w, gout).items():
                G[key + "/" + k] = v
        # JSD (N-d variant, loss.py:165-180) and Entropy_2D on the spread case
        ps = [torch.softmax(z, 1) for z in base]
        G[f"jsd_K{K}_C{C}_spread/ref_JSD_reduce"] = np.asarray(JSD()(ps, reduce=True).item(), dtype=np.float32)
        G[f"jsd_K{K}_C{C}_spread/ref_JSD_map"] = JSD()(ps, reduce=False).numpy()
        G[f"jsd_K{K}_C{C}_spread/ref_entropy0"] = Entropy_2D()(ps[0]).numpy()

    # ---------------- KL family (a5-a7) ----------------
    for C, (B, H, W) in {2: (2, 16, 16), 4: (2, 16, 16), 19: (1, 8, 12)}.items():
        g = gen(SEED + 100 + C)
        zp = 3 * torch.randn(B, C, H, W, generator=g)
        zy = 3 * torch.randn(B, C, H, W, generator=g)
        gout = torch.randn(B, H, W, generator=g)
        key = f"kl_C{C}"
        G[key + "/zp"] = zp.numpy(); G[key + "/zy"] = zy.numpy(); G[key + "/gout"] = gout.numpy()
        for name, dt in (("32", torch.float32), ("64", torch.float64)):
            p = torch.softmax(zp.to(dt), 1).requires_grad_()
            y = torch.softmax(zy.to(dt), 1).requires_grad_()
            m = KL_Divergence_2D(reduce=False)(p, y)
            m.backward(gout.to(dt))
            G[key + "/p" + name] = p.detach().numpy(); G[key + "/y" + name] = y.detach().numpy()
            G[key + "/ref_map" + name] = m.detach().numpy()
            G[key + "/ref_gp" + name] = p.grad.numpy(); G[key + "/ref_gy" + name] = y.grad.numpy()
            G[key + "/ref_mean" + name] = np.asarray(KL_Divergence_2D(reduce=True)(p.detach(), y.detach()).item())
            # the trainer's composite: softmax -> KL(reduce=True)(adv, real.detach()) -> backward
            zl = zp.to(dt).clone().requires_grad_()
            loss = KL_Divergence_2D(reduce=True)(torch.softmax(zl, 1), y.detach())
            loss.backward()
            G[key + "/ref_gzp_mean" + name] = zl.grad.numpy()
            # logits variants
            ql = zy.to(dt).clone().requires_grad_(); pl = zp.to(dt).clone().requires_grad_()
            ml = VATGenerator.kl_div_with_logit(ql, pl)
            ml.backward(gout.to(dt))
            G[key + "/ref_logit_map" + name] = ml.detach().numpy()
            G[key + "/ref_logit_gpl" + name] = pl.grad.numpy(); G[key + "/ref_logit_gql" + name] = ql.grad.numpy()
            m2 = KL_Divergence_2D_Logit(reduce=False)(zp.to(dt), zy.to(dt))
            G[key + "/ref_logit2d_map" + name] = m2.numpy()
            G[key + "/ref_kldiv_map" + name] = KL_div(reduce=False)(p.detach(), y.detach()).numpy()

    # ---------------- VAT / FGSM elementwise (a8, a9) ----------------
    g = gen(SEED + 200)
    for name, shape in {"med": (4, 1, 32, 32), "city": (2, 3, 16, 24), "odd": (3, 1, 7, 9)}.items():
        d = torch.randn(*shape, generator=g)
        G[f"vat_{name}/d"] = d.numpy()
        G[f"vat_{name}/ref_l2"] = VATGenerator._l2_normalize(d.clone()).numpy()
        img = torch.rand(*shape, generator=g); grad = torch.randn(*shape, generator=g)
        grad.view(-1)[::7] = 0.0
        adv, noise = FSGMGenerator.adversarial_fgsm(img, grad, epsilon=0.05)
        G[f"vat_{name}/img"] = img.numpy(); G[f"vat_{name}/grad"] = grad.numpy()
        G[f"vat_{name}/ref_fgsm_adv"] = adv.numpy(); G[f"vat_{name}/ref_fgsm_noise"] = noise.numpy()

    # ---------------- Dice (a10, a11) ----------------
    for C, (B, H, W) in {2: (3, 16, 24), 4: (4, 16, 16), 19: (2, 12, 12), 5: (2, 9, 7)}.items():
        g = gen(SEED + 300 + C)
        x = 3 * torch.randn(B, C, H, W, generator=g)
        gt = torch.randint(0, C, (B, 1, H, W), generator=g)
        cases = {"logits": (x, gt), "probs": (torch.softmax(x, 1), gt)}
        # exact ties: duplicate class 0's score into the arg-max class on a third of the pixels
        xt = x.clone()
        mx = xt.max(1, keepdim=True)[0]
        tie = torch.rand(B, 1, H, W, generator=g) < 0.33
        xt[:, 0:1] = torch.where(tie, mx, xt[:, 0:1])
        xt[:, C - 1:C] = torch.where(tie, mx, xt[:, C - 1:C])
        cases["ties"] = (xt, gt)
        cases["all_background"] = (torch.cat([torch.full((B, 1, H, W), 5.0), -5.0 * torch.ones(B, C - 1, H, W)], 1),
                                   torch.zeros(B, 1, H, W, dtype=torch.long))
        gts = torch.zeros(B, 1, H, W, dtype=torch.long)
        gts[:, 0, 1, 1] = C - 1
        cases["single_pixel"] = (x, gts)
        cases["gt3d"] = (x, gt.squeeze(1))  # test/test_listAggregatedMeter.py:12-13 passes [B,H,W] labels
        for cname, (xx, gg) in cases.items():
            key = f"dice_C{C}_{cname}"
            G[key + "/x"] = xx.numpy(); G[key + "/gt"] = gg.numpy()
            m2 = DiceMeter(method="2d", C=C); m2.add(xx, gg)
            m3 = DiceMeter(method="3d", C=C); m3.add(xx, gg)
            G[key + "/ref_2d"] = m2.log.numpy(); G[key + "/ref_3d"] = m3.log.numpy()
        # meter statistics after several adds (value() structure, dice_meter.py:57-64)
        m = DiceMeter(method="2d", C=C, report_axises=[1] if C > 1 else "all")
        for j in range(3):
            xx = 3 * torch.randn(B, C, H, W, generator=g); gg = torch.randint(0, C, (B, 1, H, W), generator=g)
            G[f"dice_C{C}_meter/x{j}"] = xx.numpy(); G[f"dice_C{C}_meter/gt{j}"] = gg.numpy()
            m.add(xx, gg)
        (rm, rs), (ms, ss) = m.value()
        G[f"dice_C{C}_meter/ref_report"] = np.asarray([rm.item(), rs.item()], dtype=np.float32)
        G[f"dice_C{C}_meter/ref_means"] = ms.numpy(); G[f"dice_C{C}_meter/ref_stds"] = ss.numpy()
        # contract: out-of-range labels raise AssertionError (utils/utils.py:190)
        bad = gt.clone(); bad[0, 0, 0, 0] = C
        try:
            DiceMeter(method="2d", C=C).add(x, bad)
            raised = False
        except AssertionError:
            raised = True
        G[f"dice_C{C}_logits/ref_bad_label_raises"] = np.asarray(raised)

    # ---------------- IoU / confusion (a12, a13) ----------------
    for C, (B, H, W) in {4: (2, 16, 16), 19: (2, 16, 24), 2: (2, 8, 8)}.items():
        g = gen(SEED + 400 + C)
        key = f"iou_C{C}"
        iou = IoU(C, ignore_index=255)
        for j in range(2):
            x = 3 * torch.randn(B, C, H, W, generator=g)
            gt = torch.randint(0, C, (B, 1, H, W), generator=g)
            gt[torch.rand(B, 1, H, W, generator=g) < 0.05] = 255
            if j == 1:  # exact ties -> first index
                x[:, 1] = x[:, 0]
            iou.add(predicted=x, target=gt)
            G[f"{key}/x{j}"] = x.numpy(); G[f"{key}/gt{j}"] = gt.numpy()
            G[f"{key}/ref_conf_after{j}"] = iou.conf_metric.conf.astype(np.int64).copy()
        v = iou.value()
        for k2 in ("Overall_Acc", "Mean_Acc", "FreqW_Acc", "Validated_Mean_IoU", "Mean_IoU"):
            G[f"{key}/ref_{k2}"] = np.asarray(v[k2], dtype=np.float64)
        G[f"{key}/ref_Class_IoU"] = v["Class_IoU"].numpy()
        # integer-map input ([N,H,W] ints, iou.py:49-50)
        iou2 = IoU(C, ignore_index=255)
        pm = torch.randint(0, C, (B, H, W), generator=g)
        iou2.add(pm, gt.squeeze(1))
        G[f"{key}/pred_map"] = pm.numpy(); G[f"{key}/ref_conf_from_map"] = iou2.conf_metric.conf.astype(np.int64).copy()

    path = os.path.join(OUT, "reference_golden.npz")
    np.savez_compressed(path, **G)
    print(f"wrote {path}: {len(G)} arrays, {os.path.getsize(path) / 1e6:.2f} MB")


if __name__ == "__main__":
    main()
