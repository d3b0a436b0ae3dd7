"""GPU parity: class maps, one-hot helpers, functional Dice, ensemble voting and kappa (SURVEY.md 8a11, 8b
"Functional Dice", 8f.4) through the C ABI vs the reference fixtures (tests/golden/reference_golden_aux.npz) and
the numpy oracle (oracle/oracle_aux.py) on seeded inputs.  Integer / one-hot / class results bit-exact; the
float32 soft vote and Dice values bit-exact (same operation order, IEEE divide); kappa within 1e-12.
Nothing here reads /root/reference.
"""
import os

import numpy as np
import pytest
import torch

import oracle_aux as A

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden_aux.npz"))
CS = [2, 4, 19, 5]


@pytest.fixture(scope="module")
def dct():
    import dct_b200
    assert torch.cuda.is_available()
    return dct_b200


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def N(t):
    return t.detach().cpu().numpy()


@pytest.mark.parametrize("C", CS)
def test_class_maps_vs_reference(C, dct, dev):
    k = f"aux_C{C}"
    U = dct.utils
    x, p, gt = T(G[k + "/x"], dev), T(G[k + "/p"], dev), T(G[k + "/gt"], dev)
    assert np.array_equal(N(U.pred2class(x)), G[k + "/pred2class_x"])
    assert np.array_equal(N(U.probs2class(p)), G[k + "/probs2class_p"])
    assert np.array_equal(N(U.pred2png(x)), G[k + "/pred2class_x"].astype(np.uint8))
    oh = U.class2one_hot(gt, C)
    assert oh.dtype == torch.int32 and np.array_equal(N(oh), G[k + "/class2one_hot_gt"])
    assert np.array_equal(N(U.probs2one_hot(p)), G[k + "/probs2one_hot_p"])
    assert np.array_equal(N(U.predlogit2one_hot(x)), G[k + "/predlogit2one_hot_x"])
    with pytest.raises(AssertionError):   # logits are not a simplex (utils.py:180)
        U.probs2class(x)
    with pytest.raises(AssertionError):   # label outside [0,C) (utils.py:190)
        bad = gt.clone(); bad[0, 0, 0] = C
        U.class2one_hot(bad, C)


@pytest.mark.parametrize("C", CS)
def test_functional_dice_vs_reference(C, dct, dev):
    k = f"aux_C{C}"
    U = dct.utils
    lab, pred = T(G[k + "/class2one_hot_gt"], dev), T(G[k + "/probs2one_hot_p"], dev)
    assert np.array_equal(N(U.dice_coef(lab, pred)), G[k + "/dice_coef"])
    assert np.array_equal(N(U.dice_batch(lab, pred)), G[k + "/dice_batch"])
    assert np.array_equal(N(U.intersection(lab, pred)), G[k + "/intersection"])
    assert np.array_equal(N(U.onehot_dice_counts(lab, pred)), A.onehot_dice_counts(G[k + "/class2one_hot_gt"], G[k + "/probs2one_hot_p"]))
    assert U.one_hot(lab) is True
    broken = lab.clone(); broken[0, 0, 1, 1] = 1 - broken[0, 0, 1, 1]
    assert U.one_hot(broken) is False
    two = lab.clone(); two[0, :, 2, 2] = 0; two[0, 0, 2, 2] = 2
    assert U.one_hot(two) is False
    with pytest.raises(AssertionError):   # meta_dice asserts one_hot(label) (utils.py:223)
        U.dice_coef(broken, pred)
    assert U.sset(lab, [0, 1]) and not U.sset(two, [0, 1]) and U.uniq(two) == {0, 1, 2}
    assert U.simplex(T(G[k + "/p"], dev)) and not U.simplex(T(G[k + "/x"], dev))


@pytest.mark.parametrize("K", [2, 3, 4])
@pytest.mark.parametrize("C", CS)
def test_voting_and_kappa_vs_reference(C, K, dct, dev):
    k = f"aux_C{C}_K{K}"
    E = dct.ensemble
    views = [T(G[f"{k}/view{j}"], dev) for j in range(K)]
    soft = E.Ensembleway("soft")(views)
    assert np.array_equal(N(soft), G[k + "/soft"])
    assert np.array_equal(N(E.vote_class(views)), G[k + "/soft_class"])
    assert np.array_equal(N(E.vote_class(views, uint8=True)), G[k + "/soft_class"].astype(np.uint8))
    hard = E.Ensembleway("hard")(views)
    assert hard.dtype == torch.float32 and np.array_equal(N(hard), G[k + "/hard"])
    assert np.array_equal(N(E.vote_class(views, hard=True)), G[k + "/hard"].argmax(1))
    considered = [int(c) for c in G[k + "/kappa_considered"]]
    preds = [dct.utils.pred2class(v) for v in views]
    km = E.KappaMetrics(num_classes=C)
    km.add(predicts=preds, target=dct.utils.pred2class(soft), considered_classes=considered)
    assert np.allclose(np.asarray(km.kappa[0]), G[k + "/kappa_vs_vote"], rtol=0, atol=1e-12)
    assert km.value().shape == (K,)
    gt = T(G[f"aux_C{C}/gt"], dev)
    k2 = E.Kappa2Annotator(num_classes=C); k2.add(preds[0], preds[1], gt=gt, considered_classes=considered)
    assert abs(k2.kappa[0] - float(G[k + "/kappa2"])) <= 1e-12
    k2 = E.Kappa2Annotator(); k2.add(preds[0], preds[1], gt=gt, considered_classes=None)
    assert abs(k2.kappa[0] - float(G[k + "/kappa2_all"])) <= 1e-12


@pytest.mark.parametrize("shape", [(2, 3, 7, 9), (1, 4, 1, 1), (3, 19, 33, 17), (2, 64, 8, 8)])
def test_odd_shapes_vs_oracle(shape, dct, dev):
    """Ragged sizes (HW not a multiple of 4: scalar kernels), single pixel, C at the ABI maximum, NaN / inf / ties."""
    rng = np.random.default_rng(7)
    B, C, H, W = shape
    x = (3 * rng.standard_normal(shape)).astype(np.float32)
    x.reshape(B, C, -1)[0, :, 0] = 0.25                      # full tie
    if H * W > 2:
        x.reshape(B, C, -1)[0, C - 1, 1] = np.nan            # NaN is maximal for torch.max
        x.reshape(B, C, -1)[0, 0, 2] = np.inf
    U, E = dct.utils, dct.ensemble
    assert np.array_equal(N(U.pred2class(T(x, dev))), A.pred2class(x))
    ref_t = torch.from_numpy(x).max(1)[1].numpy()            # ATen on the host agrees with the restatement
    assert np.array_equal(A.pred2class(x), ref_t)
    gt = rng.integers(0, C, (B, H, W))
    oh = A.class2one_hot(gt, C)[0]
    assert np.array_equal(N(U.class2one_hot(T(gt, dev), C)), oh)
    views = [np.where(np.isfinite(x), x, 0).astype(np.float32) + rng.standard_normal(shape).astype(np.float32) for _ in range(3)]
    tv = [T(v, dev) for v in views]
    assert np.array_equal(N(E.soft_vote(tv)), A.soft_vote(views))
    hard, win = A.hard_vote(views)
    assert np.array_equal(N(E.hard_vote(tv)), hard)
    assert np.array_equal(N(E.vote_class(tv, hard=True)), win)
    pred_oh = A.class2one_hot(win, C)[0]
    assert np.array_equal(N(U.onehot_dice_counts(T(oh, dev), T(pred_oh, dev))), A.onehot_dice_counts(oh, pred_oh))
    assert np.array_equal(N(U.dice_coef(T(oh, dev), T(pred_oh, dev))), A.dice_from_counts(A.onehot_dice_counts(oh, pred_oh)))


def test_full_size_properties(dct, dev):
    """BASELINE c4 size (B=16, C=19, 512x1024): size-independent properties instead of an oracle run.
    one-hot of the arg-max sums to 1 per pixel; Dice of a tensor with itself is exactly 1 where the class occurs;
    the hard vote of K copies of one view is that view's one-hot; the soft vote of K copies is the view itself
    (K a power of two: exact)."""
    B, C, H, W = 16, 19, 512, 1024
    g = torch.Generator(device=dev).manual_seed(1234)
    x = 3 * torch.randn(B, C, H, W, device=dev, generator=g)
    U, E = dct.utils, dct.ensemble
    oh = U.predlogit2one_hot(x)
    assert U.one_hot(oh)
    cls = U.pred2class(x)
    assert torch.equal(oh.argmax(1), cls) and torch.equal(cls, x.max(1)[1])
    counts = U.onehot_dice_counts(oh, oh)
    assert torch.equal(counts[..., 0], counts[..., 1]) and torch.equal(counts[..., 1], counts[..., 2])
    assert int(counts[..., 0].sum()) == B * H * W
    assert torch.equal(E.hard_vote([x, x, x]).to(torch.int32), oh)
    assert torch.equal(E.soft_vote([x, x, x, x]), x)
    assert torch.equal(U.class2one_hot(cls, C), oh)
