#!/usr/bin/env python
"""Second batch of fixtures from the UNMODIFIED reference (TEST INFRASTRUCTURE): the backward passes that
tests/golden/reference_golden.npz does not hold, and the VAT power iteration.

Run in the build container only (needs /root/reference):

    python oracle/make_golden_r2.py        ->  tests/golden/reference_golden_r2.npz

Every ``ref_`` array is an output of the reference's own modules (imported through oracle/ref_shim.py, torch CPU) or of
torch autograd through them, on the seeded inputs stored beside it:

  ent_C*   Entropy_2D / Entropy (loss/loss.py:53-84): map, and d(sum map*gout)/dp          -> dct_entropy_bwd_f32
  sm_C*    F.softmax(z, 1) (models/segmentators.py:50): p, and dz under an upstream gp      -> dct_softmax_bwd_f32
  kldiv_C* KL_div (loss/loss.py:87-107): map + both gradients under gout; mean + gradients  -> dct_kl_div_{fwd,bwd}_f32
  vat_*    VATGenerator.__call__ (utils/AEGenerator.py:93-119) AS INTENDED: the method itself raises at :107 (it passes
           an attribute __init__ never sets), so the fixture composes the reference's own static
           helpers ``_l2_normalize`` / ``kl_div_with_logit`` in exactly the order of :93-119, with the start direction
           ``d`` stored (the method draws it from the global RNG) and a small seeded conv net whose weights are stored.
"""
import os
import sys
import warnings

import numpy as np
import torch
import torch.nn as nn

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
from generalframework.loss import Entropy, Entropy_2D, KL_div  # noqa: E402
from generalframework.utils.AEGenerator import VATGenerator  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
SEED = 1234


def gen(seed):
    g = torch.Generator()
    g.manual_seed(seed)
    return g


def vat_net(cin, C, g):
    """conv3x3 -> tanh -> conv3x3: small, smooth, every logit depends on a 5x5 patch of the image."""
    net = nn.Sequential(nn.Conv2d(cin, 8, 3, padding=1), nn.Tanh(), nn.Conv2d(8, C, 3, padding=1))
    with torch.no_grad():
        for p in net.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * 0.5)
    return net


def vat_intended(net, img, d0, xi, eps, ip):
    """AEGenerator.py:93-119 with the reference's own helpers; ``d0`` replaces the N(0,1) draw of :97."""
    with torch.no_grad():
        pred = net(img)
    d = VATGenerator._l2_normalize(d0.clone())
    net.zero_grad()
    for _ in range(ip):
        d = xi * VATGenerator._l2_normalize(d)
        d.requires_grad = True
        y_hat = net(img + d)
        delta_kl = VATGenerator.kl_div_with_logit(pred.detach(), y_hat)
        delta_kl.mean().backward()
        d = d.grad.data.clone()
        net.zero_grad()
    d = VATGenerator._l2_normalize(d)
    r_adv = eps * d
    img_adv = torch.clamp(img + r_adv.detach(), 0, 1)
    return img_adv.detach(), r_adv.detach()


def main():
    os.makedirs(OUT, exist_ok=True)
    G = {}
    for C, (B, H, W) in {2: (2, 16, 24), 4: (2, 16, 16), 19: (1, 8, 12), 5: (2, 5, 7)}.items():
        g = gen(SEED + 500 + C)
        z = 3 * torch.randn(B, C, H, W, generator=g)
        gout = torch.randn(B, H, W, generator=g)
        # ---- Entropy_2D / Entropy backward
        key = f"ent_C{C}"
        G[key + "/gout"] = gout.numpy()
        for name, dt in (("32", torch.float32), ("64", torch.float64)):
            p = torch.softmax(z.to(dt), 1).detach().requires_grad_()
            m = Entropy_2D()(p)
            m.backward(gout.to(dt))
            G[key + "/p" + name] = p.detach().numpy()
            G[key + "/ref_map" + name] = m.detach().numpy()
            G[key + "/ref_gp" + name] = p.grad.numpy()
            p2 = p.detach().clone().requires_grad_()
            Entropy()(p2).mean().backward()
            G[key + "/ref_gp_mean" + name] = p2.grad.numpy()
        # ---- softmax backward
        key = f"sm_C{C}"
        gp = torch.randn(B, C, H, W, generator=g)
        G[key + "/z"] = z.numpy(); G[key + "/gp"] = gp.numpy()
        for name, dt in (("32", torch.float32), ("64", torch.float64)):
            zz = z.to(dt).clone().requires_grad_()
            p = torch.softmax(zz, 1)
            p.backward(gp.to(dt))
            G[key + "/ref_p" + name] = p.detach().numpy()
            G[key + "/ref_gz" + name] = zz.grad.numpy()
        # ---- KL_div forward + backward (both arguments)
        key = f"kldiv_C{C}"
        zq = 3 * torch.randn(B, C, H, W, generator=g)
        G[key + "/gout"] = gout.numpy()
        for name, dt in (("32", torch.float32), ("64", torch.float64)):
            p = torch.softmax(z.to(dt), 1).detach().requires_grad_()
            q = torch.softmax(zq.to(dt), 1).detach().requires_grad_()
            m = KL_div(reduce=False)(p, q)
            m.backward(gout.to(dt))
            G[key + "/p" + name] = p.detach().numpy(); G[key + "/q" + name] = q.detach().numpy()
            G[key + "/ref_map" + name] = m.detach().numpy()
            G[key + "/ref_gp" + name] = p.grad.numpy(); G[key + "/ref_gq" + name] = q.grad.numpy()
            p2 = p.detach().clone().requires_grad_(); q2 = q.detach().clone().requires_grad_()
            mean = KL_div(reduce=True)(p2, q2)
            (0.37 * mean).backward()
            G[key + "/ref_mean" + name] = np.asarray(mean.item())
            G[key + "/ref_gp_mean" + name] = p2.grad.numpy(); G[key + "/ref_gq_mean" + name] = q2.grad.numpy()

    # ---- VAT power iteration as intended (a8 + a6 composed)
    for name, (B, cin, C, H, W, xi, eps, ip) in {"med_ip1": (3, 1, 4, 16, 16, 4.0, 2.5, 1),
                                                  "med_ip2": (2, 1, 4, 16, 24, 4.0, 0.7, 2),
                                                  "city_ip1": (2, 3, 19, 8, 16, 6.0, 1.5, 1)}.items():
        g = gen(SEED + 600 + len(name) + ip)
        net = vat_net(cin, C, g)
        img = torch.rand(B, cin, H, W, generator=g)
        d0 = torch.randn(B, cin, H, W, generator=g)
        key = f"vat_{name}"
        for i, p in enumerate(net.parameters()):
            G[f"{key}/w{i}"] = p.detach().numpy()
        G[key + "/img"] = img.numpy(); G[key + "/d0"] = d0.numpy()
        G[key + "/hyper"] = np.asarray([xi, eps, ip, C], dtype=np.float64)
        # float32 only: the reference's _l2_normalize asserts against a float32 torch.ones (AEGenerator.py:75)
        adv, r = vat_intended(net, img, d0, xi, eps, ip)
        G[key + "/ref_img_adv32"] = adv.numpy(); G[key + "/ref_r_adv32"] = r.numpy()

    path = os.path.join(OUT, "reference_golden_r2.npz")
    np.savez_compressed(path, **G)
    print(f"wrote {path}: {len(G)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
