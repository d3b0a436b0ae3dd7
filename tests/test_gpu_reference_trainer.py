"""The north star's contract, end to end on the GPU: the reference's OWN trainer loop with its call sites unchanged.

``CoTrainer._train_loop`` (generalframework/trainer/cotraining_totalloss.py:158-271) runs twice from the same seed on
synthetic loaders -- stock, and after ``dct_b200.install()`` -- through oracle/ref_trainer.py (which builds the trainer as
train_ACDC_cotraining.py:38-63 does and records every ``totalLoss`` and every Dice row without editing the reference).

  iteration 1 (identical weights): total loss within 1e-5 (measured: bit-identical), every Dice row bit-exact -- except rows
               whose input holds rounding collisions: pixels whose two largest ATen softmax values are within 4 ulp of each
               other.  There the arg-max depends on the exp implementation (ATen CPU / ATen CUDA / the pinned polynomial of
               DESIGN.md 3.5 all differ; an untrained UNet's near-uniform outputs produce a handful per batch), so such a row
               may move by as many pixels as it has collisions, which is what is asserted;
  later iterations: the two runs' weights drift apart chaotically (1e-7-level gradient differences through Adam's
               normalisation), so losses are compared at 2e-3 and Dice rows at 1e-2, and the drift is printed.

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
                                                (2, "unet", 4, 2, 256, 256, False)])
def test_reference_train_loop_stock_vs_dropins(K, arch, C, B, H, W, adv):
    dev = torch.device("cuda:0")
    iters = 3
    kw = dict(iters=iters, K=K, arch=arch, C=C, B=B, H=H, W=W, train_jsd=True, train_adv=adv)
    stock = ref_trainer.run_train_loop(dev, False, keep_inputs=2 * K, **kw)
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
    n_inexact = 0
    for i, (a, b) in enumerate(zip(ours["dice_rows"][:per_it], stock["dice_rows"][:per_it])):
        if np.array_equal(a, b):
            continue
        n_inexact += 1
        x, gt = stock["meter_inputs"][i]                 # same tensor in both runs (identical weights, same batch)
        p = torch.softmax(x, 1)                           # DiceMeter.add applies softmax to whatever it is handed
        top2 = torch.topk(p, 2, dim=1).values
        ulp = (top2[:, 0].contiguous().view(torch.int32) - top2[:, 1].contiguous().view(torch.int32)).abs()
        collide = (ulp <= 4).flatten(1).sum(1).cpu().numpy()            # rounding collisions per image
        print(f"  meter add {i}: rows differ by {float(np.abs(a - b).max()):.2e}; rounding-collision pixels per image {collide.tolist()}")
        # a collision pixel can move one count between two classes: |d dice_c| <= 2 * moved / (|gt_c| + |pred_c|)
        sizes = np.stack([((gt.squeeze(1) == c).flatten(1).sum(1) + (p.argmax(1) == c).flatten(1).sum(1)).cpu().numpy()
                          for c in range(C)], 1).astype(np.float64)
        bound = 2.0 * (collide[:, None] + 1e-9) / np.maximum(sizes - collide[:, None], 1.0)
        assert collide.sum() > 0, "rows differ although the input holds no rounding collision"
        assert (np.abs(a.astype(np.float64) - b) <= bound + 1e-7).all(), "Dice rows differ by more than the collisions explain"
    print(f"  iteration 1: {per_it - n_inexact} of {per_it} meter adds bit-exact")
    assert float(rel.max()) <= 2e-3
    assert max(drift) <= 1e-2
    assert np.abs(ours["lab_dice"] - stock["lab_dice"]).max() <= 1e-2
    assert np.abs(ours["unlab_dice"] - stock["unlab_dice"]).max() <= 1e-2
