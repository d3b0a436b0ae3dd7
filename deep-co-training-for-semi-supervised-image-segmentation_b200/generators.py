"""Adversarial-example generators of the consistency path.

Mirrors ``generalframework/utils/AEGenerator.py``:

  FSGMGenerator   :9-51   (net fwd+bwd stays in PyTorch/cuDNN; the sign/scale/add tail is one kernel)
  VATGenerator    :54-119 (statics ``_l2_normalize`` :68-76 and ``kl_div_with_logit`` :78-91)

``VATGenerator.__call__`` is un-runnable in the reference at this commit (it passes an
attribute ``self.axises`` that ``__init__`` never sets, and every caller passes an ``axises=``
kwarg ``__init__`` rejects -- SURVEY.md section 2).  Here it implements the algorithm that code
describes, accepts the kwargs its callers pass, and keeps the perturbation on the device.
"""
from typing import Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import _lib, _runtime
from .loss import kl_div_with_logit as _kl_div_with_logit


def l2_normalize(d: Tensor, scale: float = 1.0, out: Tensor = None, img: Tensor = None, passes: int = 1) -> Tensor:
    """``scale * d / (||d_b||_2 + 1e-16)`` per sample.  ``out`` defaults to ``d`` itself (in place).
    ``passes=2`` normalises twice before scaling (``xi * _l2_normalize(_l2_normalize(d))``) in one launch.

    With ``img`` returns ``(out, clamp(img + out, 0, 1))`` from the same launch.
    """
    _runtime.require_cuda(d, "l2_normalize")
    if d.dtype != torch.float32:
        raise TypeError(f"l2_normalize: float32 expected, got {d.dtype}")
    assert d.is_contiguous(), "l2_normalize works in place and needs a contiguous tensor"
    if out is None:
        out = d
    assert out.is_contiguous() and out.shape == d.shape and out.dtype == d.dtype
    b = d.shape[0]
    m = d.numel() // max(b, 1)
    st = _runtime.state(d.device)
    adv = None
    if img is not None:
        img = img.contiguous()
        assert img.shape == d.shape and img.dtype == torch.float32
        adv = torch.empty_like(img)
    _lib.check(_lib.lib().dct_l2_normalize_f32(d.data_ptr(), out.data_ptr(), b, m, int(passes), float(scale),
                                               None if img is None else img.data_ptr(),
                                               None if adv is None else adv.data_ptr(),
                                               st.workspace.data_ptr(), _runtime.stream_ptr(d.device)),
               "dct_l2_normalize_f32")
    return out if img is None else (out, adv)


def fgsm_perturb(image: Tensor, data_grad: Tensor, epsilon: float) -> Tuple[Tensor, Tensor]:
    """``noise = epsilon * sign(grad); adv = image + noise`` in one launch -> ``(adv, noise)``."""
    _runtime.require_cuda(image, "fgsm_perturb")
    img = image.detach().contiguous()
    g = data_grad.detach().contiguous()
    assert img.shape == g.shape and img.dtype == torch.float32 and g.dtype == torch.float32
    adv = torch.empty_like(img)
    noise = torch.empty_like(img)
    _lib.check(_lib.lib().dct_fgsm_f32(img.data_ptr(), g.data_ptr(), float(epsilon), adv.data_ptr(), noise.data_ptr(),
                                       img.numel(), _runtime.stream_ptr(img.device)), "dct_fgsm_f32")
    return adv, noise


class FSGMGenerator(object):
    """Drop-in for ``FSGMGenerator`` (AEGenerator.py:9-51)."""

    def __init__(self, net: nn.Module, eplision: float = 0.05) -> None:
        super().__init__()
        self.net = net
        self.eplision = eplision

    def __call__(self, img: Tensor, gt: Tensor, criterion: nn.Module) -> Tuple[Tensor, Tensor, Tensor]:
        assert img.shape.__len__() == 4
        assert img.shape[0] >= gt.shape[0]
        img.requires_grad = True
        if img.grad is not None:
            img.grad.zero_()
        self.net.zero_grad()
        pred = self.net(img)
        if img.shape[0] > gt.shape[0]:
            gt = torch.cat((gt, pred.max(1)[1][gt.shape[0]:].unsqueeze(1)), dim=0)
        loss = criterion(pred, gt.squeeze(1))
        loss.backward()
        adv_img, noise = self.adversarial_fgsm(img, img.grad, epsilon=self.eplision)
        self.net.zero_grad()
        img.grad.zero_()
        return adv_img.detach(), noise.detach(), F.softmax(pred, 1)

    @staticmethod
    def adversarial_fgsm(image: Tensor, data_grad: Tensor, epsilon: float = 0.01) -> Tuple[Tensor, Tensor]:
        return fgsm_perturb(image, data_grad, epsilon)


class VATGenerator(object):
    """Working VAT generator with the reference's interface (AEGenerator.py:54-119).

    ``VATGenerator(net, xi=1e-6, eplision=10, ip=1, axises=None)(img, loss_name='kl') -> (img_adv, r_adv)``
    """

    def __init__(self, net: nn.Module, xi=1e-6, eplision=10, ip=1, axises=None) -> None:
        super(VATGenerator, self).__init__()
        self.xi = xi
        self.eps = eplision
        self.ip = ip
        self.net = net
        self.axises = axises  # accepted for the callers (vattrainer.py:142-143, cotraining_city.py:380,394)

    @staticmethod
    def _l2_normalize(d) -> Tensor:
        return l2_normalize(d)

    @staticmethod
    def kl_div_with_logit(q_logit, p_logit, *unused):
        return _kl_div_with_logit(q_logit, p_logit)

    def __call__(self, img: Tensor, loss_name='kl', d: Tensor = None) -> Tuple[Tensor, Tensor]:
        """``d``: optional start direction replacing the N(0,1) draw of AEGenerator.py:97 (seeded parity tests; the
        tensor is consumed -- normalised in place like the reference's own ``d``)."""
        with torch.no_grad():
            pred = self.net(img)
        if d is None:
            d = torch.randn(img.shape, dtype=torch.float32, device=img.device)
        else:
            assert d.shape == img.shape and d.device == img.device and d.dtype == torch.float32
        self.net.zero_grad()
        for it in range(self.ip):
            # first iteration: d = _l2_normalize(d) (:98) and xi * _l2_normalize(d) (:103) in one launch
            d = l2_normalize(d, scale=self.xi, passes=2 if it == 0 else 1)
            d.requires_grad = True
            y_hat = self.net(img + d)
            delta_kl = _kl_div_with_logit(pred.detach(), y_hat)  # [B,H,W]
            delta_kl.mean().backward()
            d = d.grad.detach().clone()
            self.net.zero_grad()
        # r_adv = eps * normalize(d); img_adv = clamp(img + r_adv, 0, 1): one launch
        r_adv, img_adv = l2_normalize(d, scale=self.eps, img=img.detach())
        return img_adv.detach(), r_adv.detach()
