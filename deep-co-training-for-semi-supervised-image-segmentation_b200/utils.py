"""Device-side versions of the tensor helpers the hot path and its callers go through in the reference
(``generalframework/utils/utils.py:73-80,130-235``), same names and error behaviour.

Predicates return Python bools like the reference (one 16-byte D2H of the flag vector each -- the reference
copies the whole tensor to the host for ``torch.unique``); converters stay on the device.  Assertions follow
the package check mode (``set_check_mode``): 'eager' raises ``AssertionError`` at the call like the reference,
'deferred' leaves the flag for ``raise_if_flagged()``, 'off' mirrors ``python -O``.
"""
from functools import partial
from typing import Iterable, Set

import torch
from torch import Tensor

from . import _lib, _runtime
from .metrics import dice_counts, dice_from_counts


def _bchw(t: Tensor):
    b, c = t.shape[0], t.shape[1]
    return b, c, t.numel() // max(b * c, 1)


def _flag_count(fn_name: str, flag: int, device, *args) -> int:
    """Run one predicate kernel against a private flag vector and read the count back."""
    flags = torch.zeros(_lib.NUM_FLAGS, dtype=torch.int32, device=device)
    _lib.check(getattr(_lib.lib(), fn_name)(*args, flags.data_ptr(), _runtime.stream_ptr(device)), fn_name)
    return int(flags[flag].item())


# ----------------------------------------------------------------------------------------------- predicates
def uniq(a: Tensor) -> Set:
    """``set(torch.unique(a.cpu()).numpy())`` (utils.py:130-131) -- unique on the device, only the set is copied."""
    _runtime.require_cuda(a, "uniq")
    return set(torch.unique(a).cpu().numpy())


def sset(a: Tensor, sub: Iterable) -> bool:
    """``uniq(a).issubset(sub)`` (utils.py:134-135)."""
    return uniq(a).issubset(sub)


def eq(a: Tensor, b) -> bool:
    return bool(torch.eq(a, b).all())


def simplex(t: Tensor, axis=1) -> bool:
    """``allclose(t.sum(axis), 1)`` (utils.py:142-151) evaluated by the entropy kernel's simplex flag."""
    _runtime.require_cuda(t, "simplex")
    if axis != 1:
        t = t.transpose(1, axis)
    t = t.to(torch.float32).contiguous()
    b, c, hw = _bchw(t)
    st = _runtime.state(t.device)
    flags = torch.zeros(_lib.NUM_FLAGS, dtype=torch.int32, device=t.device)
    _lib.check(_lib.lib().dct_entropy_fwd_f32(t.data_ptr(), c, b, hw, None, None, flags.data_ptr(),
                                              st.workspace.data_ptr(), _runtime.stream_ptr(t.device)),
               "dct_entropy_fwd_f32")
    return int(flags[_lib.FLAG_SIMPLEX].item()) == 0


def _as_onehot_i32(t: Tensor, what: str) -> Tensor:
    _runtime.require_cuda(t, what)
    if t.dtype != torch.int32:
        if t.is_floating_point() and not bool((t == t.round()).all()):
            raise AssertionError(f"{what}: non-integer values cannot be one-hot")
        t = t.to(torch.int32)
    return t.contiguous()


def one_hot(t: Tensor, axis=1) -> bool:
    """``simplex(t, axis) and sset(t, [0, 1])`` (utils.py:154-161): one counting pass, no host copy of ``t``."""
    if axis != 1:
        t = t.transpose(1, axis)
    t = _as_onehot_i32(t, "one_hot")
    b, c, hw = _bchw(t)
    return _flag_count("dct_onehot_dice_counts_i32", _lib.FLAG_ONEHOT, t.device, t.data_ptr(), None, c, b, hw, None) == 0


def intersection(a: Tensor, b: Tensor) -> Tensor:
    """``a & b`` after the {0,1} checks (utils.py:164-168)."""
    assert a.shape == b.shape
    assert sset(a, [0, 1])
    assert sset(b, [0, 1])
    return a & b


def union(a: Tensor, b: Tensor) -> Tensor:
    assert a.shape == b.shape
    assert sset(a, [0, 1])
    assert sset(b, [0, 1])
    return a | b


# ----------------------------------------------------------------------------------------------- converters
def _classmap(x: Tensor, mode: int, want_cls=False, want_u8=False, want_onehot=False):
    _runtime.require_cuda(x, "classmap")
    assert x.dim() == 4, x.shape
    x = x.to(torch.float32).contiguous()
    b, c, hw = _bchw(x)
    sp = x.shape[2:]
    dev = x.device
    cls = torch.empty((b,) + tuple(sp), dtype=torch.int64, device=dev) if want_cls else None
    u8 = torch.empty((b,) + tuple(sp), dtype=torch.uint8, device=dev) if want_u8 else None
    oh = torch.empty(x.shape, dtype=torch.int32, device=dev) if want_onehot else None
    st = _runtime.state(dev)
    _lib.check(_lib.lib().dct_classmap_f32(x.data_ptr(), c, b, hw, mode,
                                           None if cls is None else cls.data_ptr(),
                                           None if u8 is None else u8.data_ptr(),
                                           None if oh is None else oh.data_ptr(),
                                           _runtime.flags_ptr(st), _runtime.stream_ptr(dev)), "dct_classmap_f32")
    _runtime.after_call(st)
    return cls, u8, oh


def pred2class(pred: Tensor) -> Tensor:
    """``pred.max(1)[1]`` for logits or probabilities (utils.py:73-80)."""
    assert pred.shape.__len__() == 4, pred.shape
    return _classmap(pred, 0, want_cls=True)[0]


def pred2png(pred: Tensor) -> Tensor:
    """The uint8 class plane ``save_images`` writes (``seg.cpu().numpy().astype(np.uint8)``, utils.py:238-250),
    produced on the device in the same pass as the arg-max (1 byte/pixel crosses PCIe instead of 8)."""
    assert pred.shape.__len__() == 4, pred.shape
    return _classmap(pred, 0, want_u8=True)[1]


def probs2class(probs: Tensor) -> Tensor:
    """``probs.argmax(1)`` with the simplex assert (utils.py:178-184)."""
    b, _, w, h = probs.shape
    res = _classmap(probs, 2, want_cls=True)[0]
    assert res.shape == (b, w, h)
    return res


def class2one_hot(seg: Tensor, C: int) -> Tensor:
    """int32 one-hot ``[B,C,W,H]`` of an integer map (utils.py:187-198); labels outside [0,C) assert."""
    _runtime.require_cuda(seg, "class2one_hot")
    if len(seg.shape) == 2:
        seg = seg.unsqueeze(dim=0)
    b, w, h = seg.shape
    if seg.is_floating_point():   # Ensembleway._hardVoting hands a float class map (Summary.py:118-119)
        assert bool((seg == seg.round()).all())
    lab = seg.to(torch.int64).contiguous()
    res = torch.empty((b, C, w, h), dtype=torch.int32, device=seg.device)
    st = _runtime.state(seg.device)
    _lib.check(_lib.lib().dct_onehot_from_labels_i64(lab.data_ptr(), C, b, w * h, res.data_ptr(),
                                                     _runtime.flags_ptr(st), _runtime.stream_ptr(seg.device)),
               "dct_onehot_from_labels_i64")
    _runtime.after_call(st)
    assert res.shape == (b, C, w, h)
    return res


def probs2one_hot(probs: Tensor) -> Tensor:
    """``class2one_hot(probs2class(probs), C)`` in one pass (utils.py:201-207)."""
    res = _classmap(probs, 2, want_onehot=True)[2]
    assert res.shape == probs.shape
    return res


def predlogit2one_hot(logit: Tensor) -> Tensor:
    """``class2one_hot(probs2class(softmax(logit)), C)`` in one pass (utils.py:210-217), arg-max of the softmax
    under the pinned Dice arithmetic."""
    res = _classmap(logit, 1, want_onehot=True)[2]
    assert res.shape == logit.shape
    return res


# ----------------------------------------------------------------------------------------------- functional Dice
def onehot_dice_counts(label: Tensor, pred: Tensor) -> Tensor:
    """int64 ``[B,C,3]`` (sum label&pred, sum label, sum pred) of two int32 one-hot ``[B,C,W,H]`` tensors, with
    ``one_hot`` asserted on both in the same pass."""
    assert label.shape == pred.shape
    label, pred = _as_onehot_i32(label, "meta_dice"), _as_onehot_i32(pred, "meta_dice")
    b, c, hw = _bchw(label)
    counts = torch.empty(b, c, 3, dtype=torch.int64, device=label.device)
    st = _runtime.state(label.device)
    _lib.check(_lib.lib().dct_onehot_dice_counts_i32(label.data_ptr(), pred.data_ptr(), c, b, hw, counts.data_ptr(),
                                                     _runtime.flags_ptr(st), _runtime.stream_ptr(label.device)),
               "dct_onehot_dice_counts_i32")
    _runtime.after_call(st)
    return counts


def meta_dice(sum_str: str, label: Tensor, pred: Tensor, smooth: float = 1e-8) -> Tensor:
    """``(2*inter + smooth) / (sum_sizes + smooth)`` over ``bcwh->bc`` or ``bcwh->c`` (utils.py:221-231)."""
    if sum_str not in ("bcwh->bc", "bcwh->c"):
        raise ValueError(f"meta_dice: unsupported einsum '{sum_str}' (the reference uses 'bcwh->bc' and 'bcwh->c')")
    if smooth != 1e-8:
        raise ValueError("meta_dice: the kernel implements the reference's smooth=1e-8")
    counts = onehot_dice_counts(label, pred)
    if sum_str == "bcwh->c":
        return dice_from_counts(counts, batch_sum=True).reshape(-1)
    return dice_from_counts(counts, batch_sum=False)


dice_coef = partial(meta_dice, "bcwh->bc")
dice_batch = partial(meta_dice, "bcwh->c")  # used for 3d dice


def meta_dice_from_scores(pred_logit: Tensor, gt: Tensor, batch: bool = False) -> Tensor:
    """``dice_coef(*toOneHot(pred_logit, gt))`` (``dice_batch`` if ``batch``) without the one-hot tensors
    (dice_meter.py:12-33): one counting launch + one divide launch."""
    return dice_from_counts(dice_counts(pred_logit, gt), batch_sum=batch)
