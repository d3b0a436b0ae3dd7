"""Evaluation-time ensemble path: voting and the kappa diversity meters.

``Ensembleway`` mirrors the class inside the reference's evaluation script (``Summary.py:88-120``):
soft voting is ``torch.stack(preds).mean(0)``; hard voting copies every arg-max map to the host and runs
``np.apply_along_axis(lambda x: np.bincount(x).argmax(), ...)`` -- one Python call per pixel.  Both are one
kernel here (``dct_vote_f32``).  ``KappaMetrics`` / ``Kappa2Annotator`` (``generalframework/metrics/kappa.py:9-61``)
call ``sklearn.metrics.cohen_kappa_score`` on host copies of the class maps; here the C x C agreement counts
come from the confusion kernel (``dct_confusion_labels_i64``, exact integers) and Cohen's kappa is finished
from those counts in float64 with sklearn's formula.
"""
from typing import List, Optional

import numpy as np
import torch
from torch import Tensor

from . import _lib, _runtime
from .metrics import Metric


def _vote(predicts: List[Tensor], hard: bool, want_out=True, want_cls=False, want_u8=False):
    assert isinstance(predicts, list), type(predicts)
    assert 1 <= len(predicts) <= _lib.MAX_VIEWS
    views = []
    for p in predicts:
        _runtime.require_cuda(p, "Ensembleway")
        assert p.shape == predicts[0].shape and p.dim() == 4
        views.append(p.detach().to(torch.float32).contiguous())
    b, c = views[0].shape[0], views[0].shape[1]
    hw = views[0].numel() // (b * c)
    dev = views[0].device
    out = torch.empty_like(views[0]) if want_out else None
    cls = torch.empty((b,) + tuple(views[0].shape[2:]), dtype=torch.int64, device=dev) if want_cls else None
    u8 = torch.empty((b,) + tuple(views[0].shape[2:]), dtype=torch.uint8, device=dev) if want_u8 else None
    _lib.check(_lib.lib().dct_vote_f32(_lib.ptr_array(views), len(views), c, b, hw, int(hard),
                                       None if out is None else out.data_ptr(),
                                       None if cls is None else cls.data_ptr(),
                                       None if u8 is None else u8.data_ptr(), _runtime.stream_ptr(dev)), "dct_vote_f32")
    return out, cls, u8


def soft_vote(predicts: List[Tensor]) -> Tensor:
    """``torch.stack(predicts, 0).mean(0)`` (Summary.py:101-107)."""
    return _vote(predicts, hard=False)[0]


def hard_vote(predicts: List[Tensor]) -> Tensor:
    """float one-hot ``[B,C,H,W]`` of the per-pixel majority class, smallest class on ties (Summary.py:109-120).
    The reference concatenates the views along the batch axis (it is written for B = 1); every image of the
    batch is voted on its own here."""
    return _vote(predicts, hard=True)[0]


def vote_class(predicts: List[Tensor], hard: bool = False, uint8: bool = False) -> Tensor:
    """``pred2class(ensemble(predicts))`` without materialising the voted tensor (Summary.py:162-165)."""
    _, cls, u8 = _vote(predicts, hard, want_out=False, want_cls=not uint8, want_u8=uint8)
    return u8 if uint8 else cls


class Ensembleway(object):
    """Drop-in for ``Summary.py``'s ``Ensembleway('soft' | 'hard')``."""

    def __init__(self, ensembleway: str) -> None:
        super().__init__()
        assert ensembleway in ('soft', 'hard'), ensembleway
        self.ensembleway = ensembleway

    def __call__(self, predicts):
        if self.ensembleway == 'soft':
            return self._softVoting(predicts)
        return self._hardVoting(predicts)

    _softVoting = staticmethod(soft_vote)
    _hardVoting = staticmethod(hard_vote)


# ---------------------------------------------------------------------------------------------------------- kappa
def agreement_counts(a: Tensor, b: Tensor, num_classes: int) -> Tensor:
    """int64 ``[C,C]``: ``n[i][j] = #{pixels: b == i and a == j}`` for integer class maps of equal shape."""
    _runtime.require_cuda(a, "agreement_counts")
    _runtime.require_cuda(b, "agreement_counts")
    assert a.shape == b.shape
    a64, b64 = a.detach().to(torch.int64).contiguous().view(-1), b.detach().to(torch.int64).contiguous().view(-1)
    conf = torch.zeros(num_classes, num_classes, dtype=torch.int64, device=a.device)
    st = _runtime.state(a.device)
    _lib.check(_lib.lib().dct_confusion_labels_i64(a64.data_ptr(), b64.data_ptr(), a64.numel(), num_classes,
                                                   conf.data_ptr(), _runtime.flags_ptr(st),
                                                   _runtime.stream_ptr(a.device)), "dct_confusion_labels_i64")
    _runtime.after_call(st)
    return conf


def cohen_kappa_from_counts(conf: np.ndarray) -> float:
    """Cohen's kappa of a C x C agreement table, sklearn's arithmetic (``cohen_kappa_score``, weights=None):
    ``1 - sum(w * n) / sum(w * expected)`` with ``w = 1 - I`` and ``expected = outer(rows, cols) / total``.
    Classes absent from both raters contribute zero rows and columns and do not change the value."""
    conf = np.asarray(conf, dtype=np.float64)
    n_classes = conf.shape[0]
    sum0, sum1 = conf.sum(axis=0), conf.sum(axis=1)
    with np.errstate(invalid="ignore", divide="ignore"):
        expected = np.outer(sum0, sum1) / np.sum(sum0)
        w = np.ones((n_classes, n_classes)) - np.eye(n_classes)
        k = np.sum(w * conf) / np.sum(w * expected)
    return float(1.0 - k)


def _num_classes(maps: List[Tensor], given: Optional[int]) -> int:
    if given is not None:
        return int(given)
    return int(max(int(m.max().item()) for m in maps)) + 1


class KappaMetrics(Metric):
    """Drop-in for ``KappaMetrics`` (metrics/kappa.py:9-38): kappa of every model's class map against the
    (voted) target over the pixels whose target class is in ``considered_classes``."""

    def __init__(self, num_classes: Optional[int] = None) -> None:
        super().__init__()
        self.kappa = []
        self.num_classes = num_classes

    def add(self, predicts: List[Tensor], target: Tensor, considered_classes: List[int]):
        for predict in predicts:
            assert predict.shape == target.shape
        c = _num_classes(list(predicts) + [target], self.num_classes)
        keep = np.zeros(c, dtype=bool)
        keep[[k for k in considered_classes if 0 <= k < c]] = True
        row = []
        for predict in predicts:
            conf = agreement_counts(predict, target, c).cpu().numpy()   # rows = target class
            conf[~keep, :] = 0                                          # mask = target in considered_classes
            row.append(cohen_kappa_from_counts(conf.T))                 # cohen_kappa_score(predict, target)
        self.kappa.append(row)

    def reset(self):
        self.kappa = []

    def value(self):
        return torch.from_numpy(np.nanmean(torch.Tensor(self.kappa).numpy(), 0)).float()

    def summary(self):
        return {f'kappa{i}': self.value()[i].item() for i in range(len(self.value()))}

    def detailed_summary(self):
        return {f'kappa{i}': self.value()[i].item() for i in range(len(self.value()))}


class KappaMetrics2(KappaMetrics):
    """The ``metrics2`` flavour (generalframework/metrics2/kappa.py:31-32): ``value()`` is the plain mean over the
    log (a NaN kappa -- an empty or single-class mask -- propagates instead of being skipped by ``nanmean``)."""

    def value(self):
        return torch.Tensor(self.kappa).mean(0)


class Kappa2Annotator(KappaMetrics):
    """Drop-in for ``Kappa2Annotator`` (metrics/kappa.py:41-61): agreement of two predictions over the pixels
    whose ground truth is in ``considered_classes``."""

    def __init__(self, num_classes: Optional[int] = None) -> None:
        super().__init__(num_classes)

    def add(self, predict1: Tensor, predict2: Tensor, gt: Tensor = None, considered_classes=[1, 2, 3]):
        assert predict1.shape == predict2.shape
        c = _num_classes([predict1, predict2], self.num_classes)
        p2 = predict2.detach().to(torch.int64)
        if considered_classes is not None:
            # pixels outside the considered ground-truth classes get the label -1, which the counting kernel skips
            keep = torch.zeros_like(p2, dtype=torch.bool)
            g = gt.detach().reshape(p2.shape)
            for k in considered_classes:
                keep |= (g == k)
            p2 = torch.where(keep, p2, torch.full_like(p2, -1))
        conf = agreement_counts(predict1, p2, c).cpu().numpy()          # rows = predict2
        self.kappa.append(cohen_kappa_from_counts(conf.T))              # cohen_kappa_score(y1=predict1, y2=predict2)

    def value(self, **kwargs):
        return torch.Tensor(self.kappa).mean()
