"""GPU parity of the backward kernels and the VAT power iteration against tests/golden/reference_golden_r2.npz
(oracle/make_golden_r2.py: outputs of the unmodified reference and of torch autograd through it), plus the same-device
ATen comparison of the pinned Dice arg-max on near-tie inputs.

  dct_entropy_bwd_f32   Entropy_2D / Entropy       generalframework/loss/loss.py:53-84 (back-propagated in the
                                                   reference's own test/test_loss.py:52)
  dct_softmax_bwd_f32   F.softmax(z, 1)            generalframework/models/segmentators.py:50
  dct_kl_div_{fwd,bwd}  KL_div                     generalframework/loss/loss.py:87-107
  VATGenerator.__call__                            generalframework/utils/AEGenerator.py:93-119

Tolerance: 1e-5 of the quantity's largest magnitude (tests/util.py), the fp32 bar of BASELINE.json's north_star.
"""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

from util import assert_close, cases

pytestmark = pytest.mark.gpu

G2 = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden_r2.npz"))


@pytest.fixture(scope="module")
def dct():
    import dct_b200
    return dct_b200


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def T(a, dev, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return t.requires_grad_() if grad else t


def N(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("case", cases(G2, "ent_"))
def test_entropy_backward_vs_reference(case, dct, dev):
    p = T(G2[case + "/p32"], dev, grad=True)
    gout = T(G2[case + "/gout"], dev)
    m = dct.Entropy_2D()(p)
    assert_close(N(m), G2[case + "/ref_map32"], floor=1.0, what="entropy map")
    m.backward(gout)
    scale = float(np.abs(G2[case + "/ref_gp64"]).max())
    assert_close(N(p.grad), G2[case + "/ref_gp32"], floor=scale, what="d Entropy_2D / dp")
    assert_close(N(p.grad), G2[case + "/ref_gp64"], floor=scale, what="d Entropy_2D / dp (fp64 reference)")
    # Entropy (N-d) under .mean(): the upstream arrives as a broadcast scalar, not a contiguous map
    p2 = T(G2[case + "/p32"], dev, grad=True)
    dct.Entropy()(p2).mean().backward()
    assert_close(N(p2.grad), G2[case + "/ref_gp_mean32"], floor=float(np.abs(G2[case + "/ref_gp_mean64"]).max()),
                 what="d mean Entropy / dp")


@pytest.mark.parametrize("case", cases(G2, "sm_"))
def test_softmax_dim1_forward_backward_vs_reference(case, dct, dev):
    z = T(G2[case + "/z"], dev, grad=True)
    gp = T(G2[case + "/gp"], dev)
    p = dct.softmax_dim1(z)
    assert_close(N(p), G2[case + "/ref_p32"], rtol=2e-7, floor=1.0, what="softmax")
    p.backward(gp)
    scale = float(np.abs(G2[case + "/ref_gz64"]).max())
    assert_close(N(z.grad), G2[case + "/ref_gz32"], floor=scale, what="softmax backward")
    assert_close(N(z.grad), G2[case + "/ref_gz64"], floor=scale, what="softmax backward (fp64 reference)")
    # composed with a drop-in loss: softmax_dim1 -> Entropy_2D -> mean, against torch autograd on the same device
    z2 = T(G2[case + "/z"], dev, grad=True)
    dct.Entropy_2D()(dct.softmax_dim1(z2)).mean().backward()
    z3 = T(G2[case + "/z"], dev).double().requires_grad_()
    q = torch.softmax(z3, 1)
    (-(q * torch.log(q + 1e-16)).sum(1)).mean().backward()
    assert_close(N(z2.grad), N(z3.grad), floor=float(z3.grad.abs().max()), what="softmax -> entropy chain")


@pytest.mark.parametrize("case", cases(G2, "kldiv_"))
def test_kl_div_forward_backward_vs_reference(case, dct, dev, oracle):
    p = T(G2[case + "/p32"], dev, grad=True)
    q = T(G2[case + "/q32"], dev, grad=True)
    gout = T(G2[case + "/gout"], dev)
    m = dct.KL_div(reduce=False)(p, q)
    assert m.requires_grad, "KL_div must stay on the autograd graph (reference loss.py:99-107 is differentiable)"
    sm = float(np.abs(G2[case + "/ref_map64"]).max())
    assert_close(N(m), G2[case + "/ref_map32"], floor=sm, what="KL_div map")
    m.backward(gout)
    for got, name in ((p.grad, "gp"), (q.grad, "gq")):
        scale = float(np.abs(G2[f"{case}/ref_{name}64"]).max())
        assert_close(N(got), G2[f"{case}/ref_{name}32"], rtol=2e-5, floor=scale, what=f"KL_div {name}")
        assert_close(N(got), G2[f"{case}/ref_{name}64"], rtol=2e-5, floor=scale, what=f"KL_div {name} (fp64 reference)")
    # oracle restatement on the same inputs (per element: q/p spans 7 decades, so a max-scaled check alone is blind there)
    ogp, ogq = oracle.kl_div_bwd(G2[case + "/p32"], G2[case + "/q32"], G2[case + "/gout"])
    np.testing.assert_allclose(N(p.grad), ogp, rtol=1e-5, atol=1e-5 * float(np.abs(ogp).max()) * 1e-3)
    np.testing.assert_allclose(N(q.grad), ogq, rtol=1e-5, atol=0)
    # reduce=True (the ctor default): mean and gradients under an upstream scalar of 0.37
    p2 = T(G2[case + "/p32"], dev, grad=True)
    q2 = T(G2[case + "/q32"], dev, grad=True)
    mean = dct.KL_div()(p2, q2)
    assert mean.dim() == 0
    assert abs(mean.item() - float(G2[case + "/ref_mean64"])) <= 1e-5 * max(1.0, abs(float(G2[case + "/ref_mean64"])))
    (0.37 * mean).backward()
    for got, name in ((p2.grad, "gp_mean"), (q2.grad, "gq_mean")):
        assert_close(N(got), G2[f"{case}/ref_{name}32"], rtol=2e-5, floor=float(np.abs(G2[f"{case}/ref_{name}64"]).max()),
                     what=f"KL_div {name}")
    # only p requires a gradient: q's buffer is not produced
    p3 = T(G2[case + "/p32"], dev, grad=True)
    dct.KL_div(reduce=False)(p3, T(G2[case + "/q32"], dev)).backward(gout)
    assert torch.equal(p3.grad, p.grad)


def _vat_net(cin, C, weights, dev):
    net = nn.Sequential(nn.Conv2d(cin, 8, 3, padding=1), nn.Tanh(), nn.Conv2d(8, C, 3, padding=1)).to(dev)
    with torch.no_grad():
        for p, w in zip(net.parameters(), weights):
            p.copy_(torch.from_numpy(w).to(dev))
    return net


def _vat_torch(net, img, d0, xi, eps, ip):
    """SURVEY.md Appendix B "VAT (intended)" in stock torch on the same device (float32)."""
    def l2n(d):
        return d / (d.reshape(d.shape[0], -1).norm(dim=1).view(-1, 1, 1, 1) + 1e-16)
    with torch.no_grad():
        pred = net(img)
    d = l2n(d0)
    for _ in range(ip):
        d = (xi * l2n(d)).requires_grad_()
        y_hat = net(img + d)
        q = F.softmax(pred, 1)
        kl = (q * F.log_softmax(pred, 1)).sum(1) - (q * F.log_softmax(y_hat, 1)).sum(1)
        kl.mean().backward()
        d = d.grad.detach().clone()
        net.zero_grad()
    r = eps * l2n(d)
    return torch.clamp(img + r, 0, 1), r


@pytest.mark.parametrize("case", cases(G2, "vat_"))
def test_vat_generator_vs_reference(case, dct, dev):
    """VATGenerator.__call__ with the start direction injected == the reference's helpers composed as AEGenerator.py:93-119."""
    xi, eps, ip, C = G2[case + "/hyper"]
    ip, C = int(ip), int(C)
    img = T(G2[case + "/img"], dev)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    net = _vat_net(img.shape[1], C, [G2[f"{case}/w{i}"] for i in range(4)], dev)
    gen = dct.VATGenerator(net, xi=float(xi), eplision=float(eps), ip=ip, axises=[1, 2, 3])
    img_adv, r_adv = gen(img, loss_name="kl", d=T(G2[case + "/d0"], dev).clone())
    ref_r, ref_adv = G2[case + "/ref_r_adv32"], G2[case + "/ref_img_adv32"]
    # the direction passes through two conv backward passes (cuDNN vs the reference's CPU kernels): 1e-5 of max |r|
    assert_close(N(r_adv), ref_r, rtol=1e-5, what="r_adv vs reference")
    assert_close(N(img_adv), ref_adv, rtol=1e-5, floor=1.0, what="img_adv vs reference")
    nr = r_adv.reshape(r_adv.shape[0], -1).norm(dim=1)
    assert torch.allclose(nr, torch.full_like(nr, float(eps)), rtol=1e-4)
    assert float(img_adv.min()) >= 0.0 and float(img_adv.max()) <= 1.0
    # same device, stock torch composition of Appendix B
    want_adv, want_r = _vat_torch(net, img, T(G2[case + "/d0"], dev).clone(), float(xi), float(eps), ip)
    assert_close(N(r_adv), N(want_r), rtol=1e-5, what="r_adv vs torch composition")
    assert_close(N(img_adv), N(want_adv), rtol=1e-5, floor=1.0, what="img_adv vs torch composition")
    for p in net.parameters():
        assert p.grad is None or float(p.grad.abs().max()) == 0.0, "the generator must leave no gradient on the net"


def test_vat_generator_seeded_random_direction(dct, dev):
    """Without an injected direction the draw comes from torch's device RNG: same seed, same perturbation (up to the
    atomics of cuDNN's convolution backward, which are not run-to-run deterministic)."""
    torch.manual_seed(3)
    net = nn.Conv2d(1, 4, 3, padding=1).to(dev)
    img = torch.rand(2, 1, 32, 32, device=dev)
    gen = dct.VATGenerator(net, xi=1.0, eplision=0.5, ip=1)
    torch.manual_seed(11); a1, r1 = gen(img)
    torch.manual_seed(11); a2, r2 = gen(img)
    assert_close(N(r1), N(r2), rtol=1e-5, what="r_adv, same seed")
    assert_close(N(a1), N(a2), rtol=1e-5, floor=1.0, what="img_adv, same seed")
    torch.manual_seed(12); _, r3 = gen(img)
    assert float((r3 - r1).abs().max()) > 1e-3 * float(r1.abs().max()), "another seed must give another direction"


@pytest.mark.parametrize("C", [2, 4, 19])
def test_dice_near_ties_vs_aten_same_device(C, dct, dev, oracle, record_property):
    """The pinned Dice arg-max against stock ATen ``softmax(x, 1).argmax(1)`` on THIS device, on the
    adversarial 0..64-ulp near-tie tensor.  Where the two inputs of a pair are distinct but collide after exp / divide
    rounding the answer depends on the exp implementation (ATen CPU = Sleef, ATen CUDA = libdevice, here = the pinned
    polynomial), so the mismatch count is REPORTED, and every mismatch must be such a collision: the two classes that
    disagree hold ATen probabilities within 2 ulp of each other.  Exact ties and everything >= 2^-15 apart must agree."""
    g = torch.Generator().manual_seed(100 + C)
    B, H, W = 2, 128, 128
    x = 3 * torch.randn(B, C, H, W, generator=g)
    mx, am = x.max(1, keepdim=True)
    other = (am + 1 + torch.randint(0, max(C - 1, 1), am.shape, generator=g)) % C
    ulps = torch.randint(0, 65, am.shape, generator=g)
    near = mx.clone()
    for _ in range(64):
        step = torch.nextafter(near, torch.full_like(near, -1e30))
        near = torch.where(ulps > 0, step, near)
        ulps = ulps - 1
    x.scatter_(1, other, near)
    xd = x.to(dev)
    gt = torch.randint(0, C, (B, 1, H, W), generator=g).to(dev)
    counts = dct.dice_counts(xd, gt)                        # [B,C,3] (I, G, P)
    probs = torch.softmax(xd, 1)
    pred_aten = probs.argmax(1)
    pred_spec = torch.from_numpy(oracle.predict(x.numpy(), "dice")).to(dev)
    # the kernel's counts are exactly the counts of the pinned prediction (bit-exact vs the oracle, as elsewhere)
    P_spec = torch.stack([(pred_spec == c).flatten(1).sum(1) for c in range(C)], 1)
    assert torch.equal(counts[:, :, 2], P_spec)
    mism = pred_aten != pred_spec
    n_mism = int(mism.sum())
    record_property("near_tie_pixels", int(B * H * W))
    record_property("aten_vs_pinned_mismatches", n_mism)
    print(f"\nC={C}: pinned arg-max vs ATen on the same device: {n_mism} of {B * H * W} near-tie pixels differ")
    if n_mism:
        pa = probs.gather(1, pred_aten.unsqueeze(1)).squeeze(1)[mism]
        ps = probs.gather(1, pred_spec.unsqueeze(1)).squeeze(1)[mism]
        ulp = torch.abs(pa.view(torch.int32) - ps.view(torch.int32))
        assert int(ulp.max()) <= 2, f"a mismatch that is not a rounding collision: {int(ulp.max())} ulp apart in ATen's own softmax"
    # inputs that are exact ties or at least 2^-15 apart never depend on the exp implementation
    top2 = torch.topk(xd, 2, dim=1).values
    gap = top2[:, 0] - top2[:, 1]
    safe = (gap == 0) | (gap >= 2.0 ** -15)
    assert torch.equal(pred_aten[safe], pred_spec[safe])


@pytest.mark.parametrize("K,C,B,H,W", [(3, 4, 32, 256, 256), (2, 2, 3, 64, 72), (2, 4, 5, 9, 12), (2, 19, 2, 16, 24),
                                       (4, 4, 40, 128, 128)])
def test_fused_dice_counts_overwrite_mode(K, C, B, H, W, dct, dev):
    """DCT_COUNTS_OVERWRITE (``accumulate=False``): the fused launch clears the counters itself -- a buffer full of garbage
    must come back holding exactly what a zeroed buffer accumulates, launch after launch (the flag in the workspace is
    re-armed by every launch), on the fused tile kernel (C <= 4) and on the K counting launches (C = 19, odd sizes)."""
    g = torch.Generator(device=dev).manual_seed(17)
    zs = [3 * torch.randn(B, C, H, W, generator=g, device=dev) for _ in range(K)]
    gt = torch.randint(0, C, (B, 1, H, W), generator=g, device=dev)
    want = torch.zeros(K, B, C, 3, dtype=torch.int64, device=dev)
    lw = dct.jsd_consistency_from_logits([z.clone().requires_grad_() for z in zs], labels=gt, dice_counts=want)
    for rep in range(3):
        got = torch.full((K, B, C, 3), 123456789 + rep, dtype=torch.int64, device=dev)
        zr = [z.clone().requires_grad_() for z in zs]
        lg = dct.jsd_consistency_from_logits(zr, labels=gt, dice_counts=got, accumulate=False)
        assert torch.equal(got, want), f"rep {rep}"
        assert lg.item() == lw.item()
    # and accumulation still accumulates
    acc = want.clone()
    dct.jsd_consistency_from_logits([z.clone().requires_grad_() for z in zs], labels=gt, dice_counts=acc)
    assert torch.equal(acc, 2 * want)
    # evaluation (no gradient wanted): forward kernel + counting launches
    with torch.no_grad():
        ev = torch.full((K, B, C, 3), -7, dtype=torch.int64, device=dev)
        dct.jsd_consistency_from_logits(zs, labels=gt, dice_counts=ev, accumulate=False)
    assert torch.equal(ev, want)


def test_l2_normalize_every_launch_path(dct, dev, oracle):
    """The three implementations behind dct_l2_normalize_f32, each against the oracle: the co-resident one-launch kernel
    (per-sample exchange through tagged words in the workspace; also launched back to back so that every launch must see
    fresh tags), the cluster kernels (B > 256 samples, or a grid that cannot be co-resident) and the grid-wide passes."""
    g = torch.Generator().manual_seed(23)
    for shape, reps in [((32, 1, 256, 256), 4), ((4, 1, 512, 512), 3), ((2, 1, 64, 64), 3), ((300, 1, 64, 64), 2),
                        ((257, 1, 256, 256), 1), ((16, 3, 512, 1024), 1), ((7, 1, 100, 100), 2)]:
        d = torch.randn(*shape, generator=g)
        img = torch.rand(*shape, generator=g)
        want = oracle.l2_normalize(d.numpy())
        dd = d.to(dev)
        for _ in range(reps):   # launches back to back on one stream, no host sync in between
            got = dct.l2_normalize(dd.clone())
            twice = dct.l2_normalize(dd.clone(), scale=0.5, passes=2)
            r, adv = dct.l2_normalize(dd.clone(), scale=10.0, img=img.to(dev))
        assert_close(N(got), want, what=f"l2 {shape}")
        assert_close(N(twice), oracle.l2_normalize(want) * np.float32(0.5), what=f"l2 x2 {shape}")
        wadv, wr = oracle.vat_apply(img.numpy(), want, 10.0)
        assert_close(N(r), wr, what=f"r_adv {shape}")
        assert_close(N(adv), wadv, floor=1.0, what=f"img_adv {shape}")
