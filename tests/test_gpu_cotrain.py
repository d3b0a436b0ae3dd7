"""GPU: one co-training iteration through CoTrainStep against a plain PyTorch fp32 composition of the
same iteration (generalframework/trainer/cotraining_totalloss.py:203-250; Cityscapes flavour cotraining_city.py:228-262)
on tiny stand-in networks: total loss within 1e-5, parameter gradients within 1e-4 of their scale, meters bit-exact.
"""
import copy

import numpy as np
import pytest
import torch
import torch.nn.functional as F
from torch import nn

pytestmark = pytest.mark.gpu


def tiny_net(cin, C, seed):
    torch.manual_seed(seed)
    return nn.Sequential(nn.Conv2d(cin, 8, 3, padding=1), nn.ReLU(), nn.Conv2d(8, C, 1))


def torch_iteration(nets, lab, unlab, C, cot_w, adv_w, eps, ignore_index=255):
    """The reference iteration in stock ATen ops (fp32): sup CE + mean JSD + FGSM adversarial KL."""
    total, sups = 0, []
    for k, (img, gt) in enumerate(lab):
        s = F.cross_entropy(nets[k](img), gt.squeeze(1), ignore_index=ignore_index)
        sups.append(s)
        total = total + s
    uimg = unlab[0]
    probs = [F.softmax(n(uimg), 1) for n in nets]
    ent = lambda p: -(p * (p + 1e-16).log()).sum(1)  # noqa: E731
    mean = sum(probs[1:], probs[0]) / len(probs)
    jsd = (ent(mean) - sum(ent(p) for p in probs) / len(probs)).mean()
    total = total + cot_w * jsd
    # _FSGM_adv_training (cotraining_totalloss.py:366-393)
    img = torch.cat((lab[1][0], uimg), 0).detach().clone().requires_grad_(True)
    pred = nets[1](img)
    gt = torch.cat((lab[1][1], pred.detach().max(1)[1][lab[1][1].shape[0]:].unsqueeze(1)), 0)
    (g_img,) = torch.autograd.grad(F.cross_entropy(pred, gt.squeeze(1), ignore_index=ignore_index), img)
    img_adv = (img.detach() + eps * g_img.sign()).detach()
    real = F.softmax(pred.detach(), 1)
    p = F.softmax(nets[0](img_adv), 1)
    adv = ((real * (real + 1e-10).log()).sum(1) - (real * (p + 1e-10).log()).sum(1)).mean()
    total = total + adv_w * adv
    return total, sups, jsd, adv


@pytest.mark.parametrize("meter,C,cin,H,W", [("dice", 4, 1, 32, 32), ("iou", 19, 3, 16, 32), ("dice", 2, 1, 24, 40)])
def test_cotrain_step_matches_torch_composition(meter, C, cin, H, W):
    import dct_b200
    from dct_b200.cotrain import CoTrainConfig, CoTrainStep
    dev = torch.device("cuda:0")
    K, BL, BU = 2, 2, 3
    g = torch.Generator().manual_seed(2024 + C)
    lab = [(torch.rand(BL, cin, H, W, generator=g).to(dev), torch.randint(0, C, (BL, 1, H, W), generator=g).to(dev)) for _ in range(K)]
    if meter == "iou":
        for _, gt in lab:
            gt[torch.rand(gt.shape, device=dev) < 0.1] = 255
    unlab = (torch.rand(BU, cin, H, W, generator=g).to(dev), torch.randint(0, C, (BU, 1, H, W), generator=g).to(dev))
    nets = [tiny_net(cin, C, 10 + k).to(dev) for k in range(K)]
    ref_nets = copy.deepcopy(nets)
    opts = [torch.optim.SGD(n.parameters(), lr=0.0) for n in nets]       # lr 0: gradients stay inspectable
    cot_w, adv_w, eps = 0.5, 0.05, 0.03
    old = dct_b200.set_check_mode("deferred")
    try:
        step = CoTrainStep(nets, opts, CoTrainConfig(num_classes=C, train_jsd=True, train_adv=True, cot_weight=cot_w,
                                                     adv_weight=adv_w, fgsm_eps=eps, meter=meter), dev, ddp=False)
        total = step.step(lab, unlab)
        rep = step.report.reduce()
        dct_b200.raise_if_flagged()
    finally:
        dct_b200.set_check_mode(old)
    ref_total, ref_sups, ref_jsd, ref_adv = torch_iteration(ref_nets, lab, unlab, C, cot_w, adv_w, eps)
    ref_total.backward()
    assert abs(total.item() - ref_total.item()) <= 1e-5 * max(1.0, abs(ref_total.item()))
    want = torch.stack([s.detach() for s in ref_sups] + [ref_jsd.detach(), ref_adv.detach()]).double().cpu()
    assert torch.allclose(rep["losses"], want, rtol=1e-5, atol=1e-6), (rep["losses"], want)
    for n, rn in zip(nets, ref_nets):
        for p, rp in zip(n.parameters(), rn.parameters()):
            scale = float(rp.grad.abs().max()) + 1e-12
            assert float((p.grad - rp.grad).abs().max()) <= 1e-4 * scale
    with torch.no_grad():
        if meter == "iou":
            for k in range(K):
                pred = ref_nets[k](lab[k][0]).max(1)[1].view(-1)
                gt = lab[k][1].view(-1)
                keep = (gt >= 0) & (gt < C)
                conf = torch.bincount(gt[keep] * C + pred[keep], minlength=C * C).view(C, C).cpu()
                assert torch.equal(rep["confusion"][k], conf)
            assert int(rep["unlab_dice"].numel()) == K * C      # no meters on the unlabeled branch: counts stay empty
        else:
            for k in range(K):
                pred = ref_nets[k](lab[k][0]).max(1)[1]
                gt = lab[k][1].squeeze(1)
                i = torch.stack([((pred == c) & (gt == c)).sum() for c in range(C)]).float()
                s = torch.stack([(pred == c).sum() + (gt == c).sum() for c in range(C)]).float()
                assert torch.equal(rep["lab_dice"][k], ((2 * i + 1e-8) / (s + 1e-8)).cpu())
