"""CPU: pin the oracle's supervised-branch restatement (cross-entropy) against the reference.

Fixtures: tests/golden/reference_golden_sup.npz, written by oracle/make_golden_sup.py from the unmodified
reference (CrossEntropyLoss2d, generalframework/loss/loss.py:12-25, and its autograd gradients).
Tolerance: 1e-5 scaled in fp32 (tests/util.py), 1e-11 for the fp64 twin.
"""
import os

import numpy as np
import pytest

from util import assert_close, cases


def _names(prefix):
    g = np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_golden_sup.npz"))
    return [c for c in cases(g, prefix) if not c.endswith("all_ignored")]


def ce_case_args(G, case):
    """(weight, reduction, upstream, gout) of a fixture case (see make_golden_sup.py)."""
    name = case.split("_", 2)[2]
    w = G[case + "/weight"]
    weight = None if w.size == 0 else w
    reduction = {"sum": "sum", "none": "none"}.get(name, "mean")
    return weight, reduction, 0.37, G[case + "/gout"]


@pytest.mark.parametrize("case", _names("ce_"))
def test_cross_entropy_oracle(case, golden_sup, oracle):
    G = golden_sup
    x, gt = G[case + "/x"], G[case + "/gt"]
    weight, reduction, up, gout = ce_case_args(G, case)
    for suf, dt, tol in (("32", np.float32, 1e-5), ("64", np.float64, 1e-11)):
        loss, gz, bad = oracle.cross_entropy(x.astype(dt), gt, None if weight is None else weight.astype(dt),
                                             255, reduction, upstream=up, gout=gout)
        assert bad == 0
        ref_l, ref_g = G[case + "/ref_loss" + suf], G[case + "/ref_gz" + suf]
        assert_close(loss, ref_l, rtol=tol, floor=1.0, what="loss" + suf)
        assert_close(gz, ref_g, rtol=tol, floor=float(np.abs(ref_g).max()), what="grad" + suf)
    if (case + "/ref_dice2d") in G.files:
        assert np.array_equal(oracle.dice(x, gt, "2d"), G[case + "/ref_dice2d"])


@pytest.mark.parametrize("C", [2, 4, 19, 5])
def test_cross_entropy_all_ignored(C, golden_sup, oracle):
    G = golden_sup
    x = G[f"ce_C{C}_all_ignored/x"]
    gt = np.full((x.shape[0],) + x.shape[2:], 255, dtype=np.int64)
    with np.errstate(invalid="ignore", divide="ignore"):
        loss, gz, bad = oracle.cross_entropy(x, gt)
    assert bad == 0 and np.isnan(loss) and np.isnan(G[f"ce_C{C}_all_ignored/ref_loss32"])
    # the reference's gradients are 0 * (1/0-weighted) = NaN-free zeros only where ATen masks them; compare
    ref = G[f"ce_C{C}_all_ignored/ref_gz32"]
    assert np.array_equal(np.nan_to_num(gz, nan=0.0), np.nan_to_num(ref, nan=0.0))


def test_cross_entropy_bad_label_is_counted(oracle):
    x = np.zeros((1, 3, 2, 2), np.float32)
    gt = np.array([[[0, 3], [255, -1]]], dtype=np.int64)
    _, _, bad = oracle.cross_entropy(x, gt)
    assert bad == 2   # 3 and -1; 255 is the ignore_index
