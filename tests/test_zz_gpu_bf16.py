"""GPU parity of the bf16 entry points (networks under autocast): dct_jsd_fwdbwd_bf16, dct_kl_logit_bf16,
dct_kl_from_logits_fwdbwd_bf16, dct_ce_fwdbwd_bf16.

The bf16 kernels read and write 2-byte rows but compute in fp32, so on the SAME (bf16-representable) inputs
  * losses equal the fp32 path's to fp32 accuracy (tolerance 1e-5 scaled, the fp32 bar),
  * gradients equal the fp32 path's gradients rounded to bf16: tolerance 1e-2 of the largest gradient (the bar
    BASELINE.json's north_star states for bf16; one bf16 ulp is 2^-8 = 3.9e-3 relative),
  * integer Dice counts are bit-exact (same arg-max of the same values).
The fp32 path is itself pinned by the reference fixtures and the oracle (tests/test_gpu_parity.py); the float64
torch composition below is a second, independent check of the loss.

Gradients are ALSO pinned to the CPU oracle run on the bf16-rounded inputs (``_oracle_*`` below), at the same 1e-2 bar:
the fp32 CUDA path is not the only witness.

History: written at the end of round 1 as non-strict xfail (no GPU minutes left); the driver's round-end run showed all
26 cases passing (GPUTEST_r01: 26 xpassed), so the marker is gone and a regression now fails the suite.
"""
import math

import pytest
import torch

import numpy as np

pytestmark = pytest.mark.gpu

BF16_GRAD_TOL = 1e-2    # of the largest |gradient| (north_star: 1e-2 in bf16)
FP32_LOSS_TOL = 1e-5    # losses are accumulated in fp32 / fixed point from identical inputs


@pytest.fixture(scope="module")
def dct():
    import dct_b200
    assert torch.cuda.is_available()
    return dct_b200


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _grad_close(got_bf16, want_f32, what):
    assert got_bf16.dtype == torch.bfloat16, f"{what}: gradient dtype {got_bf16.dtype}"
    scale = float(want_f32.abs().max())
    err = float((got_bf16.float() - want_f32).abs().max())
    assert err <= BF16_GRAD_TOL * scale + 1e-30, f"{what}: max |grad err| {err:.3e} vs scale {scale:.3e}"


def _grad_close_oracle(got_bf16, want_np, what):
    """bf16 gradient vs the oracle's float32 gradient on the same (bf16-representable) inputs: 1e-2 of max |g|."""
    want = torch.from_numpy(np.ascontiguousarray(want_np)).to(got_bf16.device)
    _grad_close(got_bf16, want, what + " (oracle)")


def _logits(K, C, B, H, W, dev, seed):
    g = torch.Generator(device=dev).manual_seed(seed)
    return [(3 * torch.randn(B, C, H, W, generator=g, device=dev)).to(torch.bfloat16) for _ in range(K)]


def _jsd_mean_f64(zs):
    p = [torch.softmax(z.double(), 1) for z in zs]
    m = sum(p) / len(p)
    ent = lambda q: -(q * torch.log(q + 1e-16)).sum(1)
    return float((ent(m) - sum(ent(q) for q in p) / len(p)).mean())


# H*W % 8 == 0 -> the bf16 tile pipeline; (5, 7) and (9, 12) are not -> promotion to the fp32 kernels inside the wrapper
@pytest.mark.parametrize("K,C,B,H,W", [(2, 4, 3, 64, 72), (3, 4, 2, 40, 40), (2, 2, 2, 128, 64), (2, 19, 2, 32, 40),
                                       (3, 19, 1, 24, 40), (4, 4, 2, 16, 24), (2, 4, 2, 5, 7), (3, 19, 1, 9, 12)])
def test_jsd_consistency_bf16(K, C, B, H, W, dct, dev, oracle):
    zb = _logits(K, C, B, H, W, dev, 11)
    zf = [z.float().requires_grad_() for z in zb]
    lf = dct.jsd_consistency_from_logits(zf, weight=0.7)
    lf.backward()
    zr = [z.clone().requires_grad_() for z in zb]
    lb = dct.jsd_consistency_from_logits(zr, weight=0.7)
    lb.backward()
    assert abs(lb.item() - lf.item()) <= FP32_LOSS_TOL * math.log(K)
    assert abs(lb.item() - 0.7 * _jsd_mean_f64(zb)) <= FP32_LOSS_TOL * math.log(K)
    omean, _, ogz = oracle.jsd_logits_fwdbwd([z.float().cpu().numpy() for z in zb], 0.7, want_map=False)
    assert abs(lb.item() - 0.7 * omean) <= FP32_LOSS_TOL * math.log(K)
    for k in range(K):
        _grad_close(zr[k].grad, zf[k].grad, f"view {k}")
        _grad_close_oracle(zr[k].grad, ogz[k], f"view {k}")


@pytest.mark.parametrize("K,C,B,H,W", [(3, 4, 4, 64, 64), (2, 2, 2, 96, 80), (2, 4, 1, 8, 8), (2, 19, 2, 32, 40)])
def test_jsd_consistency_bf16_with_dice_counts(K, C, B, H, W, dct, dev, oracle):
    zb = _logits(K, C, B, H, W, dev, 12)
    g = torch.Generator(device=dev).manual_seed(5)
    gt = torch.randint(0, C, (B, 1, H, W), generator=g, device=dev)
    cf = torch.zeros(K, B, C, 3, dtype=torch.int64, device=dev)
    zf = [z.float().requires_grad_() for z in zb]
    lf = dct.jsd_consistency_from_logits(zf, labels=gt, dice_counts=cf)
    lf.backward()
    cb = torch.zeros(K, B, C, 3, dtype=torch.int64, device=dev)
    zr = [z.clone().requires_grad_() for z in zb]
    lb = dct.jsd_consistency_from_logits(zr, labels=gt, dice_counts=cb)
    lb.backward()
    assert torch.equal(cb, cf), "Dice counts of the bf16 launch differ from the fp32 launch on the same values"
    for k in range(K):   # and from a plain count (bf16 values tie often: first index wins, as argmax softmax does)
        assert torch.equal(cb[k, :, :, 1].sum(1), torch.full((B,), H * W, device=dev))
        assert torch.equal(cb[k, :, :, 2].sum(1), torch.full((B,), H * W, device=dev))
    assert abs(lb.item() - lf.item()) <= FP32_LOSS_TOL * math.log(K)
    _, _, ogz = oracle.jsd_logits_fwdbwd([z.float().cpu().numpy() for z in zb], 1.0, want_map=False)
    for k in range(K):
        _grad_close(zr[k].grad, zf[k].grad, f"view {k}")
        _grad_close_oracle(zr[k].grad, ogz[k], f"view {k}")
        oc, bad = oracle.dice_counts(zb[k].float().cpu().numpy(), gt.cpu().numpy())
        assert bad == 0 and np.array_equal(cb[k].cpu().numpy(), oc), "Dice counts of the bf16 launch differ from the oracle"


@pytest.mark.parametrize("C,B,H,W", [(4, 3, 64, 72), (2, 2, 128, 64), (19, 2, 32, 40), (4, 2, 5, 7)])
def test_kl_div_with_logit_bf16(C, B, H, W, dct, dev, oracle):
    qb, pb = _logits(2, C, B, H, W, dev, 13)
    g = torch.Generator(device=dev).manual_seed(6)
    up = torch.rand(B, H, W, generator=g, device=dev)   # a per-pixel upstream, as `.mean()` or a weighting would send
    qf, pf = qb.float().requires_grad_(), pb.float().requires_grad_()
    mf = dct.kl_div_with_logit(qf, pf)
    (mf * up).sum().backward()
    qr, pr = qb.clone().requires_grad_(), pb.clone().requires_grad_()
    mb = dct.kl_div_with_logit(qr, pr)
    (mb * up).sum().backward()
    assert mb.dtype == torch.float32 and mb.shape == (B, H, W)
    assert float((mb - mf).detach().abs().max()) <= FP32_LOSS_TOL * max(1.0, float(mf.detach().abs().max()))
    want = (torch.softmax(qb.double(), 1) * (torch.log_softmax(qb.double(), 1) - torch.log_softmax(pb.double(), 1))).sum(1)
    assert float((mb.detach().double() - want).abs().max()) <= 2e-5 * max(1.0, float(want.abs().max()))
    _grad_close(pr.grad, pf.grad, "p_logit")
    _grad_close(qr.grad, qf.grad, "q_logit")
    omap, ogp, ogq = oracle.kl_logit(qb.float().cpu().numpy(), pb.float().cpu().numpy(), up.cpu().numpy())
    assert float(np.abs(mb.detach().cpu().numpy() - omap).max()) <= 2e-5 * max(1.0, float(np.abs(omap).max()))
    _grad_close_oracle(pr.grad, ogp, "p_logit")
    _grad_close_oracle(qr.grad, ogq, "q_logit")


@pytest.mark.parametrize("C,B,H,W", [(4, 3, 64, 72), (2, 2, 128, 64), (19, 2, 32, 40), (4, 2, 5, 7)])
def test_kl_consistency_from_logits_bf16(C, B, H, W, dct, dev, oracle):
    (ab,) = _logits(1, C, B, H, W, dev, 14)
    # the target must be a simplex to 1e-5 (the reference's assert, kept by the kernel): a rounded bf16 softmax is not, so
    # use probabilities that bf16 holds exactly -- half the mass on each of two random classes (all of it when they agree)
    g = torch.Generator(device=dev).manual_seed(8)
    c1 = torch.randint(0, C, (B, 1, H, W), generator=g, device=dev)
    c2 = torch.randint(0, C, (B, 1, H, W), generator=g, device=dev)
    real_f = torch.zeros(B, C, H, W, device=dev).scatter_add_(1, c1, torch.full_like(c1, 0.5, dtype=torch.float32))
    real_f.scatter_add_(1, c2, torch.full_like(c2, 0.5, dtype=torch.float32))
    real_b = real_f.to(torch.bfloat16)
    assert torch.equal(real_b.float(), real_f)
    af = ab.float().requires_grad_()
    lf = dct.kl_consistency_from_logits(af, real_b.float(), weight=1.3)
    lf.backward()
    ar = ab.clone().requires_grad_()
    lb = dct.kl_consistency_from_logits(ar, real_b, weight=1.3)
    lb.backward()
    assert abs(lb.item() - lf.item()) <= FP32_LOSS_TOL * max(1.0, abs(lf.item()))
    _grad_close(ar.grad, af.grad, "adv_logit")
    # oracle: softmax -> KL_Divergence_2D(reduce=True) -> backward through the softmax, upstream 1.3 / N
    a_np, y_np = ab.float().cpu().numpy(), real_f.cpu().numpy()
    p_np = oracle.softmax(a_np)
    n = B * H * W
    assert abs(lb.item() - 1.3 * float(oracle.kl_fwd(p_np, y_np).mean(dtype=np.float64))) <= 2e-5 * max(1.0, abs(lf.item()))
    gp, _ = oracle.kl_bwd(p_np, y_np, np.full((B, H, W), 1.3 / n, dtype=np.float32))
    _grad_close_oracle(ar.grad, oracle.softmax_bwd(p_np, gp), "adv_logit")


@pytest.mark.parametrize("C,B,H,W,weighted", [(4, 3, 64, 72, False), (4, 2, 40, 40, True), (2, 2, 128, 64, False),
                                              (19, 2, 32, 40, True), (4, 2, 5, 7, False)])
def test_supervised_from_logits_bf16(C, B, H, W, weighted, dct, dev, oracle):
    (zb,) = _logits(1, C, B, H, W, dev, 15)
    g = torch.Generator(device=dev).manual_seed(7)
    gt = torch.randint(0, C, (B, 1, H, W), generator=g, device=dev)
    if C == 19:
        gt[:, :, ::5, ::3] = 255      # ignore_index
    w = (0.5 + torch.rand(C, generator=g, device=dev)) if weighted else None
    dice = C <= 4
    cf = torch.zeros(B, C, 3, dtype=torch.int64, device=dev) if dice else None
    cb = torch.zeros(B, C, 3, dtype=torch.int64, device=dev) if dice else None
    zf = zb.float().requires_grad_()
    lf = dct.supervised_from_logits(zf, gt, weight=w, dice_counts=cf)
    lf.backward()
    zr = zb.clone().requires_grad_()
    lb = dct.supervised_from_logits(zr, gt, weight=w, dice_counts=cb)
    lb.backward()
    want = torch.nn.functional.cross_entropy(zb.double(), gt.squeeze(1), weight=None if w is None else w.double(),
                                             ignore_index=255)
    assert abs(lb.item() - lf.item()) <= FP32_LOSS_TOL * max(1.0, abs(lf.item()))
    assert abs(lb.item() - float(want)) <= 2e-5 * max(1.0, abs(float(want)))
    _grad_close(zr.grad, zf.grad, "logits")
    oloss, ogz, _ = oracle.cross_entropy(zb.float().cpu().numpy(), gt.cpu().numpy(),
                                         weight=None if w is None else w.cpu().numpy(), ignore_index=255)
    assert abs(lb.item() - oloss) <= 2e-5 * max(1.0, abs(oloss))
    _grad_close_oracle(zr.grad, ogz, "logits")
    if dice:
        assert torch.equal(cb, cf)
        oc, bad = oracle.dice_counts(zb.float().cpu().numpy(), gt.cpu().numpy())
        assert bad == 0 and np.array_equal(cb.cpu().numpy(), oc)


def test_mixed_precision_views_are_promoted(dct, dev):
    """One bf16 view next to a float32 view: promoted to the fp32 kernels, each gradient in its input's dtype."""
    zb = _logits(2, 4, 2, 32, 32, dev, 16)
    a, b = zb[0].clone().requires_grad_(), zb[1].float().requires_grad_()
    loss = dct.jsd_consistency_from_logits([a, b])
    loss.backward()
    af, bf = zb[0].float().requires_grad_(), zb[1].float().requires_grad_()
    lf = dct.jsd_consistency_from_logits([af, bf])
    lf.backward()
    assert abs(loss.item() - lf.item()) <= FP32_LOSS_TOL
    assert a.grad.dtype == torch.bfloat16 and b.grad.dtype == torch.float32
    _grad_close(a.grad, af.grad, "bf16 view")
    assert float((b.grad - bf.grad).abs().max()) <= 1e-6 * float(bf.grad.abs().max())
