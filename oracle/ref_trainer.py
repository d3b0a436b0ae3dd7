"""Drive the UNMODIFIED reference trainer on synthetic data (TEST / BENCH INFRASTRUCTURE, not product code).

``CoTrainer._train_loop`` (generalframework/trainer/cotraining_totalloss.py:158-271) is the north star's contract: its
loss / meter call sites must work unchanged once ``dct_b200.install()`` has rebound the names.  This module builds that
trainer exactly as ``train_ACDC_cotraining.py:38-63`` does -- ``Segmentator``s from the config dicts, criterions from the
``get_loss_fn`` registry, ``CoTrainer(...)`` -- over synthetic loaders of the dataset API the loop expects
(``([img, gt], meta, name)`` items, ``set_mode``, a writable ``training``), truncates the
hard-coded 300 iterations through the module-global ``tqdm_`` and records, without touching the reference:

  * every ``totalLoss`` (the tensor ``.backward()`` is called on at cotraining_totalloss.py:247),
  * every Dice row any ``DiceMeter.add`` appends (labeled and unlabeled meters),
  * the tensors ``_train_loop`` returns,
  * wall time per iteration (device-synchronised).

The reference is found through ``oracle/ref_shim.py``: /root/reference in the build container, ``baseline/_ref`` (staged by
``tools/stage_reference.sh``) on the GPU box.
"""
import os
import sys
import tempfile
import time
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)
import ref_shim  # noqa: E402


def available() -> bool:
    return ref_shim.reference_available()


class _Trunc:
    """Stand-in for ``tqdm_(range(300))``: the first ``n`` items, with the progress-bar methods the loop calls."""

    def __init__(self, it, n):
        self.it, self.n = it, n

    def __iter__(self):
        for i, x in enumerate(self.it):
            if i >= self.n:
                return
            yield x

    def set_postfix(self, *a, **k):
        pass

    def set_description(self, *a, **k):
        pass


class SynthSet(torch.utils.data.Dataset):
    """Seeded images in [0,1] and labels in [0,C) with the reference datasets' item structure and mode switches
    (dataset/medicalDataLoader.py:22-162)."""

    def __init__(self, n, C, H, W, seed, cin=1):
        g = torch.Generator().manual_seed(seed)
        self.img = torch.rand(n, cin, H, W, generator=g)
        # blocky labels (8 x 8 patches) so that the classes form regions, as segmentation masks do
        small = torch.randint(0, C, (n, 1, (H + 7) // 8, (W + 7) // 8), generator=g)
        self.gt = small.repeat_interleave(8, 2).repeat_interleave(8, 3)[:, :, :H, :W].contiguous()
        self.training = None

    def set_mode(self, mode):
        self.training = mode

    def __len__(self):
        return self.img.shape[0]

    def __getitem__(self, i):
        return [self.img[i], self.gt[i]], i, f"synth_{i:04d}"


def run_train_loop(device, use_dropins, iters=3, K=2, arch="enet", C=4, B=4, H=256, W=256, seed=1234, train_jsd=True,
                   train_adv=True, cot_weight=0.5, adv_weight=0.05, eps=0.03, deterministic=True, warmup_iters=0):
    """One truncated epoch of the reference's co-training loop, stock (``use_dropins=False``) or after
    ``dct_b200.install()``.  Returns a dict of recorded values (numpy) and timings."""
    ref_shim.install()
    warnings.filterwarnings("ignore")
    import generalframework.trainer.cotraining_totalloss as ct
    from generalframework import ModelMode
    from generalframework.loss import get_loss_fn
    from generalframework.models import Segmentator

    if use_dropins:
        import dct_b200
        dct_b200.install()
    if deterministic:
        torch.backends.cudnn.deterministic = True
        torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(seed)
    np.random.seed(seed)
    import random
    random.seed(seed)

    arch_dict = {"name": arch, "num_classes": C}
    optim_dict = {"name": "Adam", "lr": 0.001, "weight_decay": 0.0001}          # config/ACDC_config_cotraing.yaml:5-8
    sched_dict = {"name": "StepLR", "step_size": 90, "gamma": 0.1}
    segs = [Segmentator(arch_dict=dict(arch_dict), optim_dict=dict(optim_dict), scheduler_dict=dict(sched_dict))
            for _ in range(K)]
    mk = lambda s: torch.utils.data.DataLoader(SynthSet(B * max(iters + warmup_iters, 2), C, H, W, s), batch_size=B,  # noqa: E731
                                               shuffle=False, drop_last=True, num_workers=0)
    labeled = [mk(seed + 10 + k) for k in range(K)]
    unlabeled = mk(seed + 99)
    val = mk(seed + 77)
    criterions = {"sup": get_loss_fn("cross_entropy"), "jsd": get_loss_fn("jsd"), "adv": get_loss_fn("jsd")}
    save_dir = tempfile.mkdtemp(prefix="dct_ref_trainer_")
    const = lambda v: {"name": "ConstantScheduler", "begin_epoch": 0, "max_value": v}  # noqa: E731
    trainer = ct.CoTrainer(segmentators=segs, labeled_dataloaders=labeled, unlabeled_dataloader=unlabeled,
                           val_dataloader=val, criterions=criterions, max_epoch=1, save_dir=save_dir, device=str(device),
                           axises=list(range(1, C)), adv_scheduler_dict=const(adv_weight),
                           cot_scheduler_dict=const(cot_weight), adv_training_dict={"eplision": eps}, use_tqdm=True)

    # ---- recorders (monkey-patches of module globals / torch; restored below; the reference files are not touched)
    losses, rows, stamps = [], [], []
    orig_backward = torch.Tensor.backward
    orig_tqdm = ct.tqdm_
    Meter = ct.DiceMeter

    class RecMeter(Meter):
        def add(self, pred_logit, gt):
            super().add(pred_logit, gt)
            rows.append(self.diceLog[-1].detach().clone())

    def rec_backward(self, *a, **k):
        losses.append(self.detach().clone())
        return orig_backward(self, *a, **k)

    def run(n):
        ct.tqdm_ = lambda it, **k: _Trunc(it, n)
        return trainer._train_loop(labeled_dataloaders=labeled, unlabeled_dataloader=unlabeled, epoch=0,
                                   mode=ModelMode.TRAIN, save=False, train_jsd=train_jsd, train_adv=train_adv,
                                   augment_labeled_data=False, augment_unlabeled_data=False)

    try:
        if warmup_iters:
            run(warmup_iters)       # cuDNN autotuning, module load; re-seeds at its start like every epoch
        ct.DiceMeter = RecMeter
        torch.Tensor.backward = rec_backward
        if device.type == "cuda":
            torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        lab_dice, unlab_dice = run(iters)
        if device.type == "cuda":
            torch.cuda.synchronize(device)
        dt = time.perf_counter() - t0
    finally:
        torch.Tensor.backward = orig_backward
        ct.tqdm_ = orig_tqdm
        ct.DiceMeter = Meter
        if use_dropins:
            import dct_b200
            dct_b200.uninstall()
    per_iter = 2 if train_adv else 1      # FSGMGenerator back-propagates its own CE loss before the total loss
    total = [float(v) for v in losses[per_iter - 1::per_iter]]
    return {"total_loss": np.asarray(total, dtype=np.float64),
            "all_backward_losses": np.asarray([float(v) for v in losses], dtype=np.float64),
            "dice_rows": [r.float().cpu().numpy() for r in rows],
            "lab_dice": lab_dice.detach().cpu().numpy(), "unlab_dice": unlab_dice.detach().cpu().numpy(),
            "seconds": dt, "iters": iters, "it_per_s": iters / dt,
            "meter_class": f"{Meter.__module__}.{Meter.__name__}",
            "jsd_class": f"{type(criterions['jsd']).__module__}.{type(criterions['jsd']).__name__}"}
