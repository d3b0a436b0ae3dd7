"""The north star's contract, end to end on the GPU: the reference's OWN trainer loop with its call sites unchanged.

``CoTrainer._train_loop`` (generalframework/trainer/cotraining_totalloss.py:158-271) runs twice from the same seed on
synthetic loaders -- stock, and after ``dct_b200.install()`` -- through oracle/ref_trainer.py (which builds the trainer as
train_ACDC_cotraining.py:38-63 does and records every ``totalLoss`` and every Dice row without editing the reference).

  iteration 1 (identical weights): total loss within 1e-5, every Dice row bit-exact (same logits -> same integer counts
               -> the same fp32 divide);
  later iterations: the two runs' weights have drifted by the 1e-7-level difference of the gradients passed through Adam,
               so losses are compared at 1e-3 and Dice rows at 2e-3 (a handful of boundary pixels), and the drift is printed.

Needs the staged reference (tools/stage_reference.sh -> baseline/_ref, git-ignored but shipped to the GPU box); skipped
when it is absent.
"""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import ref_trainer  # noqa: E402

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not ref_trainer.available(), reason="reference tree not staged (tools/stage_reference.sh)")]


@pytest.mark.parametrize("K,arch,C,B,H,W,adv", [(2, "enet", 4, 4, 256, 256, True), (3, "enet", 4, 2, 128, 128, True),
                                                (2, "unet", 4, 2, 64, 64, False)])
def test_reference_train_loop_stock_vs_dropins(K, arch, C, B, H, W, adv):
    dev = torch.device("cuda:0")
    iters = 3
    kw = dict(iters=iters, K=K, arch=arch, C=C, B=B, H=H, W=W, train_jsd=True, train_adv=adv)
    stock = ref_trainer.run_train_loop(dev, False, **kw)
    ours = ref_trainer.run_train_loop(dev, True, **kw)
    assert stock["meter_class"].startswith("generalframework.") and stock["jsd_class"].startswith("generalframework.")
    assert "b200" in ours["meter_class"] and "b200" in ours["jsd_class"], "install() did not reach the trainer's call sites"
    assert len(stock["total_loss"]) == len(ours["total_loss"]) == iters
    per_it = len(stock["dice_rows"]) // iters           # K labeled + K unlabeled meter adds per iteration
    assert per_it == 2 * K and len(ours["dice_rows"]) == len(stock["dice_rows"])
    rel = np.abs(ours["total_loss"] - stock["total_loss"]) / np.abs(stock["total_loss"])
    drift = [max(float(np.abs(a - b).max()) for a, b in zip(ours["dice_rows"][i * per_it:(i + 1) * per_it],
                                                            stock["dice_rows"][i * per_it:(i + 1) * per_it]))
             for i in range(iters)]
    print(f"\n{arch} K={K} {H}x{W} B={B}: total loss stock {stock['total_loss']} drop-in {ours['total_loss']} rel {rel}; "
          f"max |Dice row diff| per iteration {drift}; it/s stock {stock['it_per_s']:.2f} drop-in {ours['it_per_s']:.2f}")
    assert rel[0] <= 1e-5, f"iteration 1 total loss differs: {rel[0]:.2e}"
    for a, b in zip(ours["dice_rows"][:per_it], stock["dice_rows"][:per_it]):
        assert np.array_equal(a, b), "iteration 1 Dice rows must be bit-exact"
    assert float(rel.max()) <= 1e-3
    assert max(drift) <= 2e-3
    assert np.abs(ours["lab_dice"] - stock["lab_dice"]).max() <= 2e-3
    assert np.abs(ours["unlab_dice"] - stock["unlab_dice"]).max() <= 2e-3
