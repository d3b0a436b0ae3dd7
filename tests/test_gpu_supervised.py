"""GPU parity tests of the supervised branch: the cross-entropy kernels (through the C ABI, via
the CrossEntropyLoss2d drop-in and the fused supervised_from_logits) vs the committed reference fixtures
(tests/golden/reference_golden_sup.npz) and vs the CPU oracle on seeded inputs.  Losses and gradients within
1e-5 scaled (tests/util.py); the fused Dice counts bit-exact.  Nothing here reads /root/reference.
"""
import os

import numpy as np
import pytest
import torch

from test_oracle_sup_vs_golden import ce_case_args
from util import assert_close, cases

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden_sup.npz"))


@pytest.fixture(scope="module")
def dct():
    import dct_b200
    assert torch.cuda.is_available()
    assert dct_b200._lib.lib().dct_device_check(0) == 0
    return dct_b200


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def T(a, dev, grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    return t.requires_grad_() if grad else t


def N(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("case", [c for c in cases(G, "ce_") if not c.endswith("all_ignored")])
def test_cross_entropy_dropin_vs_reference(case, dct, dev):
    x, gt = G[case + "/x"], G[case + "/gt"]
    weight, reduction, up, gout = ce_case_args(G, case)
    kw = {"sum": dict(size_average=False), "none": dict(reduce=False)}.get(case.split("_", 2)[2], {})
    crit = dct.get_loss_fn("cross_entropy") if (weight is None and not kw) else \
        dct.CrossEntropyLoss2d(weight=None if weight is None else weight.tolist(), **kw)
    crit = crit.to(dev)
    z = T(x, dev, grad=True)
    out = crit(z, T(gt, dev).squeeze(1))
    ref_l, ref_g = G[case + "/ref_loss32"], G[case + "/ref_gz32"]
    assert out.dtype == torch.float32 and tuple(out.shape) == tuple(ref_l.shape)
    assert_close(N(out), ref_l, floor=1.0, what="loss")
    if out.dim() == 0:
        (up * out).backward()
    else:
        out.backward(T(gout, dev))
    assert_close(N(z.grad), ref_g, floor=float(np.abs(ref_g).max()), what="d/dlogits")
    # and against the fp64 run of the reference
    assert_close(N(z.grad), G[case + "/ref_gz64"], floor=float(np.abs(ref_g).max()), what="d/dlogits vs fp64")
    # no-grad forward (the _eval_loop call, cotraining_totalloss.py:294) uses the forward-only kernel
    with torch.no_grad():
        out2 = crit(T(x, dev), T(gt, dev).squeeze(1))
    assert_close(N(out2), ref_l, floor=1.0, what="loss (no_grad)")


@pytest.mark.parametrize("case", [c for c in cases(G, "ce_") if c.endswith(("plain", "confident"))])
def test_supervised_fused_with_dice_vs_reference(case, dct, dev):
    x, gt = G[case + "/x"], G[case + "/gt"]
    B, C = x.shape[0], x.shape[1]
    z = T(x, dev, grad=True)
    counts = torch.zeros(B, C, 3, dtype=torch.int64, device=dev)
    loss = dct.supervised_from_logits(z, T(gt, dev), dice_counts=counts)
    assert_close(loss.item(), G[case + "/ref_loss32"], floor=1.0, what="fused loss")
    (0.37 * loss).backward()
    ref_g = G[case + "/ref_gz32"]
    assert_close(N(z.grad), ref_g, floor=float(np.abs(ref_g).max()), what="fused grad")
    m = dct.DiceMeter(method="2d", C=C)
    m.add_counts(counts)
    assert np.array_equal(N(m.log), G[case + "/ref_dice2d"]), "Dice rows from the fused counts differ"


@pytest.mark.parametrize("C", [2, 4, 19, 5])
def test_cross_entropy_all_ignored(C, dct, dev):
    x = G[f"ce_C{C}_all_ignored/x"]
    z = T(x, dev, grad=True)
    gt = torch.full((x.shape[0],) + x.shape[2:], 255, dtype=torch.int64, device=dev)
    out = dct.CrossEntropyLoss2d()(z, gt)
    assert torch.isnan(out).item() and np.isnan(G[f"ce_C{C}_all_ignored/ref_loss32"])
    out.backward()
    ref = np.nan_to_num(G[f"ce_C{C}_all_ignored/ref_gz32"], nan=0.0)
    assert np.array_equal(np.nan_to_num(N(z.grad), nan=0.0), ref)


SHAPES = [  # (C, B, H, W): tile shapes, ragged tails, odd sizes (fallback kernel), one 19-class Cityscapes-like
    (4, 4, 256, 256), (2, 3, 96, 100), (3, 2, 40, 52), (4, 2, 33, 31), (19, 2, 64, 128), (19, 1, 37, 41), (7, 2, 24, 24),
    (4, 1, 1, 4), (2, 5, 1, 1),
]


@pytest.mark.parametrize("C,B,H,W", SHAPES)
@pytest.mark.parametrize("mode", ["plain", "weighted_ignore", "fused_dice"])
def test_cross_entropy_vs_oracle(C, B, H, W, mode, dct, dev, oracle):
    g = torch.Generator().manual_seed(1234 + 31 * C + H)
    x = 3 * torch.randn(B, C, H, W, generator=g)
    gt = torch.randint(0, C, (B, 1, H, W), generator=g)
    weight = None
    if mode == "weighted_ignore":
        gt[torch.rand(B, 1, H, W, generator=g) < 0.1] = 255
        weight = (0.25 + torch.rand(C, generator=g))
    z = x.to(dev).requires_grad_()
    counts = torch.zeros(B, C, 3, dtype=torch.int64, device=dev) if mode == "fused_dice" else None
    loss = dct.supervised_from_logits(z, gt.to(dev), weight=weight, dice_counts=counts)
    loss.backward()
    ol, og, bad = oracle.cross_entropy(x.numpy(), gt.numpy(), None if weight is None else weight.numpy(), 255, "mean")
    assert bad == 0
    assert_close(loss.item(), ol, floor=1.0, what="loss")
    assert_close(N(z.grad), og, floor=float(np.abs(og).max()), what="grad")
    if counts is not None:
        oc, obad = oracle.dice_counts(x.numpy(), gt.numpy())
        assert obad == 0 and np.array_equal(N(counts), oc), "fused Dice counts differ from the oracle"
    # the sum is bit-reproducible run to run (order-independent fixed-point accumulation / fixed-order tree)
    z2 = x.to(dev).requires_grad_()
    loss2 = dct.supervised_from_logits(z2, gt.to(dev), weight=weight,
                                       dice_counts=None if counts is None else torch.zeros_like(counts))
    assert loss2.item() == loss.item()


@pytest.mark.parametrize("C,B,H,W", SHAPES)
@pytest.mark.parametrize("variant", ["plain", "weighted_ignore_ties"])
def test_cross_entropy_fused_with_confusion_vs_oracle(C, B, H, W, variant, dct, dev, oracle):
    """cotraining_city.py:236-241: CE + IoU.add of the same (pred, gt) from one launch (dct_ce_fwdbwd_conf_f32): loss and
    gradient within 1e-5, the confusion matrix bit-exact (raw arg-max, first index on ties, 255 dropped), accumulated."""
    g = torch.Generator().manual_seed(4321 + 17 * C + W)
    x = 3 * torch.randn(B, C, H, W, generator=g)
    gt = torch.randint(0, C, (B, 1, H, W), generator=g)
    weight = None
    if variant == "weighted_ignore_ties":
        x = torch.round(x)                                   # integer-valued logits: plenty of exact ties
        gt[torch.rand(B, 1, H, W, generator=g) < 0.1] = 255
        weight = (0.25 + torch.rand(C, generator=g))
    meter = dct.IoU(C, ignore_index=255)
    conf = meter.device_counts(dev)
    z = x.to(dev).requires_grad_()
    loss = dct.supervised_from_logits(z, gt.to(dev), weight=weight, confusion=conf)
    loss.backward()
    ol, og, bad = oracle.cross_entropy(x.numpy(), gt.numpy(), None if weight is None else weight.numpy(), 255, "mean")
    assert bad == 0
    assert_close(loss.item(), ol, floor=1.0, what="loss")
    assert_close(N(z.grad), og, floor=float(np.abs(og).max()), what="grad")
    oc = oracle.confusion(x.numpy(), gt.numpy())
    assert np.array_equal(N(conf), oc), "fused confusion counts differ from the oracle"
    # the meter reads the same accumulator; a second launch accumulates; the stand-alone kernel agrees
    dct.supervised_from_logits(x.to(dev).requires_grad_(), gt.to(dev), weight=weight, confusion=conf)
    assert np.array_equal(meter.conf_metric.conf64, 2 * oc)
    ref_meter = dct.IoU(C, ignore_index=255)
    ref_meter.add(predicted=x.to(dev), target=gt.to(dev))
    assert np.array_equal(ref_meter.conf_metric.conf64, oc)
    # evaluation (no gradient wanted): loss + counts through the forward kernels
    conf2 = torch.zeros(C, C, dtype=torch.int64, device=dev)
    with torch.no_grad():
        l2 = dct.supervised_from_logits(x.to(dev), gt.to(dev), weight=weight, confusion=conf2)
    assert_close(l2.item(), ol, floor=1.0, what="eval loss")
    assert np.array_equal(N(conf2), oc)


def test_fused_confusion_nan_scores_follow_torch_max(dct, dev, oracle):
    g = torch.Generator().manual_seed(99)
    C, B, H, W = 19, 2, 64, 128
    x = 3 * torch.randn(B, C, H, W, generator=g)
    x[torch.rand(B, C, H, W, generator=g) < 0.01] = float("nan")
    gt = torch.randint(0, C, (B, 1, H, W), generator=g)
    conf = torch.zeros(C, C, dtype=torch.int64, device=dev)
    dct.supervised_from_logits(x.to(dev).requires_grad_(), gt.to(dev), confusion=conf)
    assert np.array_equal(N(conf), oracle.confusion(x.numpy(), gt.numpy()))
    pred = x.to(dev).max(1)[1]                               # the reference's own arg-max on the same device
    want = torch.bincount((gt.to(dev).view(-1) * C + pred.view(-1)), minlength=C * C).view(C, C)
    assert torch.equal(conf, want)


def test_label_hist_and_bad_labels(dct, dev):
    import ctypes
    h = dct._lib.lib()
    C = 4
    lab = torch.tensor([0, 1, 1, 3, 255, 255, 7, -2, 2, 2, 2], dtype=torch.int64, device=dev)
    hist = torch.empty(C + 2, dtype=torch.int64, device=dev)
    rc = h.dct_label_hist_i64(lab.data_ptr(), lab.numel(), C, 255, hist.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    assert hist.tolist() == [1, 2, 3, 1, 2, 2]
    # a label outside [0,C) other than ignore_index raises like PyTorch's "Target out of bounds"
    x = torch.zeros(1, C, 2, 2, device=dev, requires_grad=True)
    gt = torch.tensor([[[0, 1], [255, 9]]], dtype=torch.int64, device=dev)
    old = dct.set_check_mode("eager")
    try:
        with pytest.raises(AssertionError):
            dct.CrossEntropyLoss2d()(x, gt)
        # with the fused meter the ignore label is out of range too (class2one_hot's assert)
        gt2 = torch.tensor([[[0, 1], [255, 2]]], dtype=torch.int64, device=dev)
        with pytest.raises(AssertionError):
            dct.supervised_from_logits(x, gt2, dice_counts=torch.zeros(1, C, 3, dtype=torch.int64, device=dev))
        dct.CrossEntropyLoss2d()(x, gt2)  # fine without the meter
    finally:
        dct.set_check_mode(old)
    assert ctypes.c_int(h.dct_ce_fwd_f32(None, None, 4, 1, 4, None, 255, None, None, None, None, None)).value == -1


def test_full_size_properties(dct, dev):
    """BASELINE sizes (size-independent properties): c2 labeled batch and one Cityscapes-sized batch."""
    for C, B, H, W in [(4, 32, 256, 256), (19, 4, 512, 1024)]:
        g = torch.Generator(device=dev).manual_seed(7)
        x = 3 * torch.randn(B, C, H, W, device=dev, generator=g)
        gt = torch.randint(0, C, (B, 1, H, W), device=dev, generator=g)
        z = x.clone().requires_grad_()
        counts = torch.zeros(B, C, 3, dtype=torch.int64, device=dev)
        loss = dct.supervised_from_logits(z, gt, dice_counts=counts)
        loss.backward()
        # gradients of a softmax cross-entropy sum to zero over the class axis; total mass of |grad| <= 2/N per pixel
        assert float(z.grad.sum(1).abs().max()) <= 1e-6 / (B * H * W) * 10
        # linearity in the upstream: loss' = 3 * loss -> grad' = 3 * grad exactly up to one rounding
        z3 = x.clone().requires_grad_()
        (3.0 * dct.supervised_from_logits(z3, gt)).backward()
        assert torch.allclose(z3.grad, 3.0 * z.grad, rtol=1e-6, atol=0)
        # agrees with the stock ATen composition on the same device (not the oracle: full size)
        ref = torch.nn.functional.cross_entropy(x, gt.squeeze(1))
        assert abs(loss.item() - ref.item()) <= 1e-5 * max(1.0, abs(ref.item()))
        # Dice counts: every pixel is counted once in G and once in P
        assert int(counts[..., 1].sum()) == B * H * W and int(counts[..., 2].sum()) == B * H * W
        pred = x.argmax(1)
        inter = torch.stack([((pred == c) & (gt.squeeze(1) == c)).flatten(1).sum(1) for c in range(C)], 1)
        assert torch.equal(counts[..., 0], inter)
