#!/usr/bin/env python
"""tools/cotrain_bench.py -- co-training ITERATIONS/s (BASELINE.json metric, second half) on 1..N GPUs.

One iteration = the body of ``CoTrainer._train_loop`` (generalframework/trainer/cotraining_totalloss.py:203-264):
K labeled forward passes + CE + Dice meters, the unlabeled K-view JSD + Dice meters, the FGSM adversarial KL,
backward, K optimizer steps, and the per-iteration reporting.  Two arms on the SAME networks and data:

  ours   dct_b200.cotrain.CoTrainStep: fused kernels for every loss / meter line, device-resident reporting,
         one process per GPU with DDP (torchrun) when --gpus > 1
  aten   the reference's op-by-op composition restated in stock PyTorch on the same GPU (its losses, its
         DiceMeter helper chain with the host-side ``torch.unique`` checks, its per-iteration ``value()``
         / ``.cpu()`` progress-bar reads) -- bench-only code, the meaningful denominator;
         single GPU (the reference's multi-GPU path is nn.DataParallel, not reproduced).
  nets   the networks alone (same forward/backward passes with a trivial loss): the floor neither arm can beat.

The networks are BASELINE's own (ENet / UNet / Cityscapes ENet built by the staged reference's ``get_arch``, random init,
stock cuDNN) when baseline/_ref is staged (tools/stage_reference.sh), else a small stand-in UNet: they are out of scope
(SURVEY.md 8), and what is measured is how much of an iteration the loss / metric path costs.
Prints one JSON line per arm (rank 0) and writes them to --out.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from torch import nn  # noqa: E402

CONFIGS = {  # name: (K, C, Cin, H, W, B_lab, B_unlab)
    "c1": (2, 4, 1, 256, 256, 4, 4),      # ACDC, config/ACDC_config_cotraing.yaml batch sizes
    "c2": (3, 4, 1, 256, 256, 32, 32),    # ACDC, 3 views, batch 32
    "c3": (2, 2, 1, 512, 512, 4, 4),      # Spleen
    "c4": (2, 19, 3, 512, 1024, 4, 4),    # Cityscapes-shaped (batch 4/GPU: activations of the stand-in UNet)
}


# --------------------------------------------------------------------------------------------- stand-in network
def block(i, o):
    return nn.Sequential(nn.Conv2d(i, o, 3, padding=1, bias=False), nn.BatchNorm2d(o), nn.ReLU(inplace=True),
                         nn.Conv2d(o, o, 3, padding=1, bias=False), nn.BatchNorm2d(o), nn.ReLU(inplace=True))


class SmallUNet(nn.Module):
    def __init__(self, cin, num_classes, base=16):
        super().__init__()
        b = base
        self.e1, self.e2, self.e3, self.mid = block(cin, b), block(b, 2 * b), block(2 * b, 4 * b), block(4 * b, 8 * b)
        self.d3, self.d2, self.d1 = block(12 * b, 4 * b), block(6 * b, 2 * b), block(3 * b, b)
        self.head = nn.Conv2d(b, num_classes, 1)

    def forward(self, x):
        e1 = self.e1(x)
        e2 = self.e2(F.max_pool2d(e1, 2))
        e3 = self.e3(F.max_pool2d(e2, 2))
        m = self.mid(F.max_pool2d(e3, 2))
        up = lambda t, ref: F.interpolate(t, size=ref.shape[2:], mode="nearest")  # noqa: E731
        d3 = self.d3(torch.cat((up(m, e3), e3), 1))
        d2 = self.d2(torch.cat((up(d3, e2), e2), 1))
        d1 = self.d1(torch.cat((up(d2, e1), e1), 1))
        return self.head(d1)


REF_ARCH = {"c1": "enet", "c2": "unet", "c3": "enet", "c4": "deeplabenet"}   # config/*_cotraing.yaml "Arch: name"


def make_net(config, cin, C, base=16, prefer_reference=True):
    """The network of BASELINE's config when the unmodified reference is staged (baseline/_ref or /root/reference: ENet
    generalframework/arch/enet.py:234-244, UNet arch/network.py:196-290, Cityscapes ENet arch/deeplab/enet.py:485-648,
    built through the reference's own ``get_arch``), else the stand-in UNet above.  Returns (module, description)."""
    if prefer_reference:
        try:
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import ref_shim
            if ref_shim.reference_available():
                ref_shim.install()
                from generalframework.arch import get_arch
                name = REF_ARCH[config]
                kw = {"num_classes": C}
                net = get_arch(name, kw)
                x = torch.zeros(1, cin, 256, 256)
                with torch.no_grad():
                    net.eval()
                    assert net(x).shape[1] == C
                return net, f"reference get_arch('{name}') ({sum(p.numel() for p in net.parameters()) / 1e6:.2f} M parameters), fp32, random init"
        except Exception as e:  # input channels / optional packages the arch wants: fall back, say why
            return SmallUNet(cin, C, base), f"stand-in SmallUNet(base={base}) (reference arch unavailable: {type(e).__name__}: {e})"
    return SmallUNet(cin, C, base), f"stand-in SmallUNet(base={base}), fp32, random init"


# --------------------------------------------------------------------------------------------- the ATen arm
class Aten:
    """The reference's composition, restated op by op (utils/utils.py:130-231, loss/loss.py:12-25,70-84,110-134,
    183-196, metrics/dice_meter.py:12-83, utils/AEGenerator.py:16-51)."""

    @staticmethod
    def simplex(t, axis=1):
        s = t.sum(axis).type(torch.float32)
        return torch.allclose(s, torch.ones_like(s, dtype=torch.float32))

    @staticmethod
    def sset(a, sub):
        return set(torch.unique(a.cpu()).numpy()).issubset(sub)

    @classmethod
    def one_hot(cls, t):
        return cls.simplex(t, 1) and cls.sset(t, [0, 1])

    @classmethod
    def class2one_hot(cls, seg, C):
        assert cls.sset(seg, list(range(C)))
        res = torch.stack([seg == c for c in range(C)], dim=1).type(torch.int32)
        assert cls.one_hot(res)
        return res

    @classmethod
    def probs2one_hot(cls, probs):
        C = probs.shape[1]
        assert cls.simplex(probs)
        assert cls.simplex(probs)                      # probs2class asserts it again
        res = cls.class2one_hot(probs.argmax(dim=1), C)
        assert cls.one_hot(res)
        return res

    @classmethod
    def dice_coef(cls, label, pred):
        assert cls.one_hot(label) and cls.one_hot(pred)
        assert cls.sset(label, [0, 1]) and cls.sset(pred, [0, 1])      # intersection()
        inter = torch.einsum("bcwh->bc", label & pred).type(torch.float32)
        sizes = (torch.einsum("bcwh->bc", label) + torch.einsum("bcwh->bc", pred)).type(torch.float32)
        return (2 * inter + 1e-8) / (sizes + 1e-8)

    class DiceMeter:
        def __init__(self, C, axises):
            self.C, self.axises, self.log = C, axises, []

        def add(self, pred_logit, gt):
            oh_pred = Aten.probs2one_hot(F.softmax(pred_logit, 1))
            oh_mask = Aten.class2one_hot(gt.squeeze(1), pred_logit.shape[1])
            self.log.append(Aten.dice_coef(oh_pred, oh_mask))

        def value(self):
            log = torch.cat(self.log)
            means, stds = log.mean(0), log.std(0)
            rm = log[:, self.axises].mean(1)
            return (rm.mean(), rm.std()), (means, stds)

    @classmethod
    def entropy(cls, p):
        assert cls.simplex(p)
        return -1.0 * (p * (p + 1e-16).log()).sum(1)

    @classmethod
    def jsd(cls, probs):
        for p in probs:
            assert cls.simplex(p)
        mean = sum(probs[1:], probs[0]) / len(probs)
        return cls.entropy(mean) - sum(cls.entropy(p) for p in probs) / len(probs)

    @classmethod
    def kl(cls, p, y, eps=1e-10):
        assert cls.simplex(p) and cls.simplex(y)
        return ((y * (y + eps).log()).sum(1) - (y * (p + eps).log()).sum(1)).mean()

    @staticmethod
    def ce(logits, gt):
        return F.nll_loss(F.log_softmax(logits, 1), gt, ignore_index=255)


def aten_iteration(nets, opts, lab, unlab, meters, umeters, axises, cfg):
    K = len(nets)
    total, sup_vals = 0, []
    for k, (img, gt) in enumerate(lab):
        pred = nets[k](img)
        sup = Aten.ce(pred, gt.squeeze(1))
        meters[k].add(pred, gt)
        sup_vals.append(sup.detach().data.cpu())
        total = total + sup
    uimg, ugt = unlab
    probs = [F.softmax(n(uimg), 1) for n in nets]
    for k in range(K):
        umeters[k].add(probs[k], ugt)
    jsd = Aten.jsd(probs).mean()
    jsd.item()
    total = total + cfg["cot"] * jsd
    if cfg["adv"]:
        img = torch.cat((lab[1][0], uimg), 0).clone().requires_grad_(True)
        nets[1].zero_grad()
        pred = nets[1](img)
        gt = torch.cat((lab[1][1], pred.max(1)[1][lab[1][1].shape[0]:].unsqueeze(1)), 0)
        Aten.ce(pred, gt.squeeze(1)).backward()
        noise = cfg["eps"] * img.grad.sign()
        img_adv = (img + noise).detach()
        nets[1].zero_grad()
        adv = Aten.kl(F.softmax(nets[0](img_adv), 1), F.softmax(pred, 1).detach())
        adv.item()
        total = total + cfg["advw"] * adv
    for o in opts:
        o.zero_grad()
    total.backward()
    for o in opts:
        o.step()
    # progress bar (cotraining_totalloss.py:251-264): value() per reported class and once for the mean, per meter
    for m in list(meters) + list(umeters):
        for n in axises:
            m.value()[1][0][n].cpu()
        m.value()[0][0].cpu()
    return total


def nets_only_iteration(nets, opts, lab, unlab, cfg):
    """Same forward/backward passes and optimizer steps with a trivial loss (no loss/metric path)."""
    total = 0
    for k, (img, _) in enumerate(lab):
        total = total + nets[k](img).mean()
    uimg = unlab[0]
    for n in nets:
        total = total + n(uimg).mean()
    if cfg["adv"]:
        img = torch.cat((lab[1][0], uimg), 0).clone().requires_grad_(True)
        (g,) = torch.autograd.grad(nets[1](img).mean(), img)
        total = total + nets[0]((img + g).detach()).mean()
    for o in opts:
        o.zero_grad(set_to_none=True)
    total.backward()
    for o in opts:
        o.step()
    return total


def measure(config, arms, iters, warmup, dev, rank, world, local, base=16, adv=True, batch=None, prefer_reference=True,
            on_line=None):
    """Iterations/s of the arms on ``config`` (one process per GPU; DDP inside the ``ours`` / ``nets`` arms at world > 1).
    Returns the JSON-able lines (every rank computes them; rank 0's are the ones to print)."""
    import dct_b200
    from dct_b200.cotrain import CoTrainConfig, CoTrainStep
    K, C, cin, H, W, BL, BU = CONFIGS[config]
    if batch is not None:
        BL = BU = int(batch)
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    lab = [(torch.rand(BL, cin, H, W, device=dev, generator=g),
            torch.randint(0, C, (BL, 1, H, W), device=dev, generator=g)) for _ in range(K)]
    unlab = (torch.rand(BU, cin, H, W, device=dev, generator=g), torch.randint(0, C, (BU, 1, H, W), device=dev, generator=g))
    axises = list(range(1, C)) if C > 2 else [0, 1]
    cfg = {"cot": 0.5, "advw": 0.05, "eps": 0.03, "adv": bool(adv) and K >= 2}
    old_mode = dct_b200.set_check_mode("deferred")
    lines = []
    net_desc = [None]

    def fresh():
        torch.manual_seed(1234)
        nets = []
        for _ in range(K):
            n, net_desc[0] = make_net(config, cin, C, base, prefer_reference)
            nets.append(n.to(dev).train())
        opts = [torch.optim.Adam(n.parameters(), lr=1e-3, weight_decay=1e-4) for n in nets]
        return nets, opts

    def timed(fn):
        for _ in range(warmup):
            fn()
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ms = torch.tensor([e0.elapsed_time(e1), wall * 1e3], dtype=torch.float64, device=dev)
        if world > 1:
            torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        return float(ms[0]) / iters, float(ms[1]) / iters

    for arm in arms:
        if arm == "aten" and world > 1:
            continue
        nets, opts = fresh()
        extra = {}
        if arm == "ours":
            step = CoTrainStep(nets, opts, CoTrainConfig(num_classes=C, train_jsd=True, train_adv=cfg["adv"], cot_weight=cfg["cot"],
                                                         adv_weight=cfg["advw"], fgsm_eps=cfg["eps"],
                                                         meter="iou" if config == "c4" else "dice"), dev)
            fn = lambda: step.step(lab, unlab)  # noqa: E731
            ms, wall = timed(fn)
            rep = step.report.reduce()
            dct_b200.raise_if_flagged()
            extra = {"losses": [round(float(v), 5) for v in rep["losses"]],
                     "unlab_dice_view0": [round(float(v), 4) for v in rep["unlab_dice"][0]]}
            if "confusion" in rep:   # Cityscapes flavour: IoU meter on the labeled branch, counted by the loss kernel
                conf0 = rep["confusion"][0].double()
                iu = conf0.diag() / (conf0.sum(0) + conf0.sum(1) - conf0.diag()).clamp(min=1)
                extra["lab_mean_iou_view0"] = round(float(iu.mean()), 4)
                extra["lab_pixels_counted_view0"] = int(conf0.sum())
        elif arm == "aten":
            meters = [Aten.DiceMeter(C, axises) for _ in range(K)]
            umeters = [Aten.DiceMeter(C, axises) for _ in range(K)]
            fn = lambda: aten_iteration(nets, opts, lab, unlab, meters, umeters, axises, cfg)  # noqa: E731
            ms, wall = timed(fn)
        else:
            if world > 1:
                from torch.nn.parallel import DistributedDataParallel as DDP
                nets = [DDP(n, device_ids=[local], gradient_as_bucket_view=True, broadcast_buffers=False) for n in nets]
            fn = lambda: nets_only_iteration(nets, opts, lab, unlab, cfg)  # noqa: E731
            ms, wall = timed(fn)
        line = {"metric": "co-train iterations/sec", "arm": arm, "value": 1e3 / ms, "unit": "iterations/s",
                "ms_per_iter": ms, "wall_ms_per_iter": wall, "n_gpus": world, "images_per_iter_per_gpu": K * BL + BU,
                "images_per_sec": (K * BL + BU) * world * 1e3 / ms, "iters": iters, "warmup": warmup,
                "config": {"workload": config, "K": K, "C": C, "H": H, "W": W, "B_lab": BL, "B_unlab": BU, "adv": cfg["adv"],
                           "net": net_desc[0], "scaling": "weak"}, **extra}
        lines.append(line)
        if on_line is not None:
            on_line(line, lines)
        del nets, opts
        torch.cuda.empty_cache()
    dct_b200.set_check_mode(old_mode)
    return lines


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="c1", choices=sorted(CONFIGS))
    ap.add_argument("--arms", default="ours,aten,nets")
    ap.add_argument("--iters", type=int, default=60)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--base", type=int, default=16, help="width of the stand-in UNet")
    ap.add_argument("--adv", type=int, default=1)
    ap.add_argument("--batch", type=int, default=None, help="override the labeled / unlabeled batch per GPU")
    ap.add_argument("--stand-in", action="store_true", help="the stand-in UNet even if the reference's architectures are staged")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "cotrain"))
    args = ap.parse_args()

    from dct_b200.cotrain import init_distributed
    rank, world, local = init_distributed()
    dev = torch.device("cuda", local)
    os.makedirs(args.out, exist_ok=True)

    def on_line(line, lines):
        if rank == 0:
            print(json.dumps(line), flush=True)
            with open(os.path.join(args.out, f"cotrain_{args.config}_n{world}.jsonl"), "w") as f:   # after every arm
                for ln in lines:
                    f.write(json.dumps(ln) + "\n")

    measure(args.config, args.arms.split(","), args.iters, args.warmup, dev, rank, world, local, base=args.base, adv=bool(args.adv),
            batch=args.batch, prefer_reference=not args.stand_in, on_line=on_line)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
