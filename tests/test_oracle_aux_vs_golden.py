"""CPU: pin the numpy restatement of the class-map / one-hot / ensemble helpers (oracle/oracle_aux.py) against
the reference.

Fixtures: tests/golden/reference_golden_aux.npz, written by oracle/make_golden_aux.py from the unmodified
reference (generalframework/utils/utils.py:73-235, Summary.py:88-120, generalframework/metrics/kappa.py).
Bar: bit-exact for every integer / one-hot / class result and for the float32 Dice and soft-vote values;
kappa (float64) within 1e-12.
"""
import numpy as np
import pytest

import oracle_aux as A

CS = [2, 4, 19, 5]


@pytest.mark.parametrize("C", CS)
def test_class_maps_and_one_hot(C, golden_aux):
    G, k = golden_aux, f"aux_C{C}"
    x, p, gt = G[k + "/x"], G[k + "/p"], G[k + "/gt"]
    assert np.array_equal(A.pred2class(x), G[k + "/pred2class_x"])
    assert np.array_equal(A.pred2class(p), G[k + "/probs2class_p"])
    assert A.simplex_violations(p) == 0 and A.simplex_violations(x) > 0
    oh, bad = A.class2one_hot(gt, C)
    assert bad == 0 and oh.dtype == np.int32 and np.array_equal(oh, G[k + "/class2one_hot_gt"])
    assert np.array_equal(A.probs2one_hot(p), G[k + "/probs2one_hot_p"])
    assert A.class2one_hot(np.array([[0, C], [-1, 1]]), C)[1] == 2


@pytest.mark.parametrize("C", CS)
def test_predlogit2one_hot_uses_the_dice_spec(C, golden_aux, oracle):
    G, k = golden_aux, f"aux_C{C}"
    pred = oracle.predict(G[k + "/x"], "dice")
    assert np.array_equal(A.class2one_hot(pred, C)[0], G[k + "/predlogit2one_hot_x"])


@pytest.mark.parametrize("C", CS)
def test_functional_dice(C, golden_aux):
    G, k = golden_aux, f"aux_C{C}"
    lab, pred = G[k + "/class2one_hot_gt"], G[k + "/probs2one_hot_p"]
    counts = A.onehot_dice_counts(lab, pred)
    assert np.array_equal(A.dice_from_counts(counts), G[k + "/dice_coef"])
    assert np.array_equal(A.dice_from_counts(counts, batch_sum=True).reshape(-1), G[k + "/dice_batch"])
    assert np.array_equal(lab & pred, G[k + "/intersection"])
    assert bool(G[k + "/one_hot_true"]) and A.one_hot_violations(lab) == 0
    broken = lab.copy(); broken[0, 0, 1, 1] = 1 - broken[0, 0, 1, 1]
    assert not bool(G[k + "/one_hot_broken"]) and A.one_hot_violations(broken) > 0
    two = lab.copy(); two[0, :, 2, 2] = 0; two[0, 0, 2, 2] = 2
    assert not bool(G[k + "/one_hot_value2"]) and A.one_hot_violations(two) > 0


@pytest.mark.parametrize("K", [2, 3, 4])
@pytest.mark.parametrize("C", CS)
def test_voting_and_kappa(C, K, golden_aux):
    G, k = golden_aux, f"aux_C{C}_K{K}"
    views = [G[f"{k}/view{j}"] for j in range(K)]
    soft = A.soft_vote(views)
    assert np.array_equal(soft, G[k + "/soft"])
    assert np.array_equal(A.pred2class(soft), G[k + "/soft_class"])
    hard, win = A.hard_vote(views)
    assert np.array_equal(hard, G[k + "/hard"])
    assert np.array_equal(A.class2one_hot(win, C)[0].astype(np.float32), hard)
    # kappa of every view against the soft vote over the considered target classes (Summary.py:171-172)
    considered = G[k + "/kappa_considered"]
    target = A.pred2class(soft)
    mask = np.isin(target, considered)
    for j in range(K):
        got = A.cohen_kappa(A.pred2class(views[j])[mask], target[mask], C)
        assert abs(got - G[k + "/kappa_vs_vote"][j]) <= 1e-12
    gt = golden_aux[f"aux_C{C}/gt"]
    p0, p1 = A.pred2class(views[0]), A.pred2class(views[1])
    m = np.isin(gt, considered)
    assert abs(A.cohen_kappa(p0[m], p1[m], C) - float(G[k + "/kappa2"])) <= 1e-12
    assert abs(A.cohen_kappa(p0, p1, C) - float(G[k + "/kappa2_all"])) <= 1e-12
