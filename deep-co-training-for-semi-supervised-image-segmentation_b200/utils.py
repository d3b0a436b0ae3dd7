"""Device-side versions of the tensor helpers the hot path goes through in the reference
(``generalframework/utils/utils.py:130-235``).  Predicates return Python bools like the
reference (one small D2H each); converters stay on the device.
"""
import torch
from torch import Tensor

from . import _lib, _runtime
from .metrics import dice_counts, dice_from_counts


def simplex(t: Tensor, axis=1) -> bool:
    """``allclose(t.sum(axis), 1)`` (utils.py:142-151) evaluated by the entropy kernel's simplex flag."""
    _runtime.require_cuda(t, "simplex")
    if axis != 1:
        t = t.transpose(1, axis)
    t = t.to(torch.float32).contiguous()
    b, c = t.shape[0], t.shape[1]
    hw = t.numel() // max(b * c, 1)
    flags = torch.zeros(_lib.NUM_FLAGS, dtype=torch.int32, device=t.device)
    st = _runtime.state(t.device)
    _lib.check(_lib.lib().dct_entropy_fwd_f32(t.data_ptr(), c, b, hw, None, None, flags.data_ptr(),
                                              st.workspace.data_ptr(), _runtime.stream_ptr(t.device)),
               "dct_entropy_fwd_f32")
    return int(flags[_lib.FLAG_SIMPLEX].item()) == 0


def probs2class(probs: Tensor) -> Tensor:
    """``probs.argmax(1)`` (utils.py:178-184)."""
    b, _, w, h = probs.shape
    assert simplex(probs, 1)
    res = probs.argmax(dim=1)
    assert res.shape == (b, w, h)
    return res


def class2one_hot(seg: Tensor, C: int) -> Tensor:
    """int32 one-hot ``[B,C,W,H]`` of an integer map (utils.py:187-198)."""
    if len(seg.shape) == 2:
        seg = seg.unsqueeze(dim=0)
    assert bool(((seg >= 0) & (seg < C)).all())
    b, w, h = seg.shape
    res = torch.stack([seg == c for c in range(C)], dim=1).type(torch.int32)
    assert res.shape == (b, C, w, h)
    return res


def meta_dice_from_scores(pred_logit: Tensor, gt: Tensor, batch: bool = False) -> Tensor:
    """``dice_coef(*toOneHot(pred_logit, gt))`` (``dice_batch`` if ``batch``) without the one-hot tensors
    (dice_meter.py:12-33): one counting launch + one divide launch."""
    return dice_from_counts(dice_counts(pred_logit, gt), batch_sum=batch)
