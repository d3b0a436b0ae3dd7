"""CPU restatement (numpy) of the reference's class-map / one-hot / ensemble helpers -- TEST INFRASTRUCTURE ONLY.

Integer and byte work, restated in plain numpy in the reference's own order; every function cites the reference
lines it follows.  Pinned by ``tests/golden/reference_golden_aux.npz`` (``oracle/make_golden_aux.py`` runs the
unmodified reference functions, including the ``Ensembleway`` class lifted at run time out of ``Summary.py``).
Only ``tests/`` imports this module; the product path (``dct_b200``) never does.
"""
import numpy as np


def pred2class(x):
    """``pred.max(1)[1]`` -- generalframework/utils/utils.py:73-80 (first index on ties, NaN maximal)."""
    x = np.asarray(x)
    nan = np.isnan(x)
    cls = np.argmax(np.where(nan, np.inf, x), axis=1)
    first_nan = np.argmax(nan, axis=1)
    return np.where(nan.any(axis=1), first_nan, cls).astype(np.int64)


def simplex_violations(p):
    """#pixels failing ``torch.allclose(t.sum(1), 1)`` -- utils.py:142-151 (rtol 1e-5, atol 1e-8)."""
    s = np.asarray(p, dtype=np.float32).sum(axis=1, dtype=np.float32)
    return int((~(np.abs(s - np.float32(1)) <= np.float32(1e-8) + np.float32(1e-5))).sum())


def class2one_hot(seg, C):
    """``stack([seg == c for c in range(C)], 1).int32`` -- utils.py:187-198.  Returns (one-hot, #labels outside [0,C))."""
    seg = np.asarray(seg)
    if seg.ndim == 2:
        seg = seg[None]
    bad = int(((seg < 0) | (seg >= C)).sum())
    return np.stack([seg == c for c in range(C)], axis=1).astype(np.int32), bad


def probs2one_hot(p):
    """``class2one_hot(probs2class(probs), C)`` -- utils.py:201-207."""
    return class2one_hot(pred2class(p), np.asarray(p).shape[1])[0]


def one_hot_violations(t):
    """Counts what ``one_hot`` rejects (utils.py:154-161): values outside {0,1} plus pixels whose class column does
    not sum to exactly one (the kernel's DCT_FLAG_ONEHOT convention)."""
    t = np.asarray(t)
    return int(((t != 0) & (t != 1)).sum()) + int((t.sum(axis=1) != 1).sum())


def onehot_dice_counts(label, pred):
    """``einsum('bcwh->bc', ...)`` of ``label & pred``, ``label`` and ``pred`` -- utils.py:221-231.  int64 [B,C,3]."""
    label, pred = np.asarray(label).astype(np.int64), np.asarray(pred).astype(np.int64)
    ax = tuple(range(2, label.ndim))
    return np.stack([(label & pred).sum(axis=ax), label.sum(axis=ax), pred.sum(axis=ax)], axis=-1)


def dice_from_counts(counts, batch_sum=False):
    """``(2*inter + 1e-8) / (sum_sizes + 1e-8)`` in float32 -- utils.py:229."""
    c = np.asarray(counts, dtype=np.int64)
    if batch_sum:
        c = c.sum(axis=0, keepdims=True)
    inter = c[..., 0].astype(np.float32)
    sizes = (c[..., 1] + c[..., 2]).astype(np.float32)
    return (np.float32(2) * inter + np.float32(1e-8)) / (sizes + np.float32(1e-8))


def soft_vote(views):
    """``torch.stack(predicts, 0).mean(0)`` -- Summary.py:101-107: sequential fp32 sum over the views, then / K."""
    acc = np.asarray(views[0], dtype=np.float32).copy()
    for v in views[1:]:
        acc = acc + np.asarray(v, dtype=np.float32)
    return acc / np.float32(len(views))


def hard_vote(views):
    """Per-pixel ``np.bincount(votes).argmax()`` over the views' arg-max maps, one-hot float -- Summary.py:109-120
    (voted per image; the reference's concatenation along the batch axis is the B = 1 case).  Returns (one-hot, class)."""
    C = np.asarray(views[0]).shape[1]
    votes = np.stack([pred2class(v) for v in views], axis=0)            # [K,B,H,W]
    counts = np.stack([(votes == c).sum(axis=0) for c in range(C)], 0)  # [C,B,H,W]
    win = counts.argmax(axis=0)                                         # smallest class on ties
    return class2one_hot(win, C)[0].astype(np.float32), win.astype(np.int64)


def cohen_kappa(y1, y2, num_classes):
    """sklearn.metrics.cohen_kappa_score(y1, y2) from the C x C agreement table (metrics/kappa.py:28,57)."""
    y1, y2 = np.asarray(y1).ravel(), np.asarray(y2).ravel()
    conf = np.zeros((num_classes, num_classes), dtype=np.float64)
    np.add.at(conf, (y1, y2), 1.0)
    s0, s1 = conf.sum(axis=0), conf.sum(axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        expected = np.outer(s0, s1) / s0.sum()
        w = 1.0 - np.eye(num_classes)
        return float(1.0 - (w * conf).sum() / (w * expected).sum())
