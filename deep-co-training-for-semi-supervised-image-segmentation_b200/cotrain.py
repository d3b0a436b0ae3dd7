"""One co-training iteration with the fused hot path, data-parallel, with device-resident reporting.

``CoTrainer._train_loop`` (generalframework/trainer/cotraining_totalloss.py:158-271) does, per iteration:
  :205-216  for every segmentator k: labeled batch -> logits -> ``criterions['sup']`` -> ``diceMeters[k].add``
  :217-227  unlabeled batch -> K predictions -> ``unlabdiceMeters[k].add`` x K -> ``JSD_2D`` -> ``.mean()``
  :234-246  ``_FSGM_adv_training`` (:366-393): FGSM example of segmentator 1 on cat(labeled_2, unlabeled), KL of
            segmentator 0's prediction on it against segmentator 1's clean prediction
  :247-250  ``zero_grad`` -> ``totalLoss = sup + w_cot * jsd + w_adv * adv`` -> ``backward`` -> ``step``
  :251-264  progress bar: ``DiceMeter.value()`` (an O(iterations) ``torch.cat`` of the whole log) and ``.cpu()`` of
            K*(C+1) scalars, ``.item()`` of every loss -- a dozen host syncs per iteration
under ``nn.DataParallel`` (models/segmentators.py:34-36: scatter / gather through cuda:0 every forward).

Here the same iteration runs as one process per GPU (``torch.distributed`` + DDP around each network: gradient
all-reduce over NCCL is the only bandwidth-relevant collective), every loss / meter line goes through the fused
kernels (one launch per line: CE+Dice, JSD+K Dice, KL), and nothing is read back inside the loop: the loss sums
and the integer Dice counters accumulate in device buffers (``DeviceReport``) that are reduced across ranks and
copied to the host once per reporting interval.  The networks themselves are the caller's (stock cuDNN).
"""
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
from torch import Tensor, nn

from . import _runtime
from . import distributed as D
from .generators import fgsm_perturb
from .loss import jsd_consistency_from_logits, kl_consistency_from_logits, softmax_dim1, supervised_from_logits
from .metrics import dice_from_counts


class DeviceReport:
    """Loss sums and Dice counters of a reporting interval, kept on the device.

    ``counts['lab' | 'unlab']``: int64 ``[K, C, 3]`` (I, G, P) summed over every image seen -- the batch ('3d'-style)
    Dice the reference's progress bar approximates with a running mean of per-image rows; ``rows`` optionally
    keeps the per-image '2d' rows for the epoch-end ``DiceMeter`` statistics.  ``loss_sums``: float64 ``[K + 2]``
    (sup_0..sup_{K-1}, jsd, adv) and ``steps``.  ``reduce()`` is the only collective + host copy."""

    def __init__(self, K: int, C: int, device, keep_rows: bool = False, with_confusion: bool = False):
        self.K, self.C, self.device, self.keep_rows = K, C, device, keep_rows
        self.counts = {n: torch.zeros(K, C, 3, dtype=torch.int64, device=device) for n in ("lab", "unlab")}
        # Cityscapes trainers (cotraining_city.py:212): one IoU meter per segmentator on the labeled branch; the fused
        # loss kernel accumulates conf[gt][pred] straight into conf[k]
        self.conf = torch.zeros(K, C, C, dtype=torch.int64, device=device) if with_confusion else None
        self.loss_sums = torch.zeros(K + 2, dtype=torch.float64, device=device)
        self.steps = 0
        self.rows = {n: [[] for _ in range(K)] for n in ("lab", "unlab")}

    def reset(self):
        for c in self.counts.values():
            c.zero_()
        if self.conf is not None:
            self.conf.zero_()
        self.loss_sums.zero_()
        self.steps = 0
        self.rows = {n: [[] for _ in range(self.K)] for n in ("lab", "unlab")}

    def add_counts(self, which: str, k: int, per_image: Tensor):
        """per_image: int64 [B,C,3] written by a fused kernel for segmentator k."""
        self.counts[which][k] += per_image.sum(0)
        if self.keep_rows:
            self.rows[which][k].append(dice_from_counts(per_image))

    def add_losses(self, sup: Sequence[Tensor], jsd: Optional[Tensor], adv: Optional[Tensor]):
        vals = [s.detach().double() for s in sup]
        zero = torch.zeros((), dtype=torch.float64, device=self.device)
        vals.append(jsd.detach().double() if torch.is_tensor(jsd) else zero)
        vals.append(adv.detach().double() if torch.is_tensor(adv) else zero)
        self.loss_sums += torch.stack(vals)
        self.steps += 1

    def reduce(self, group=None) -> Dict[str, object]:
        """All-reduce the counters (int64 SUM) and loss sums over the data-parallel ranks, then ONE host copy.
        This is also where the kernels' contract flags (labels outside [0,C), non-simplex targets) are read back in
        'deferred' check mode: the reference's AssertionError surfaces here, once per reporting interval."""
        if self.device.type == "cuda":
            _runtime.raise_if_flagged()
        lab, unlab = self.counts["lab"].clone(), self.counts["unlab"].clone()
        sums = self.loss_sums.clone()
        world = 1
        if D.is_distributed():
            world = dist.get_world_size(group)
            D.all_reduce_counts(lab, group)
            D.all_reduce_counts(unlab, group)
            dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        out = {"steps": self.steps, "world": world}
        for name, c in (("lab", lab), ("unlab", unlab)):
            if c.is_cuda:
                dice = dice_from_counts(c)                                        # [K,C] float32, on the device
            else:  # host tensors (gloo tests of the reduction logic): the same float32 arithmetic
                i, s = c[..., 0].float(), (c[..., 1] + c[..., 2]).float()
                dice = (2 * i + 1e-8) / (s + 1e-8)
            out[name + "_dice"] = dice.cpu()
        if self.conf is not None:
            conf = self.conf.clone()
            if D.is_distributed():
                D.all_reduce_counts(conf, group)
            out["confusion"] = conf.cpu()                # [K,C,C] int64; IoU statistics: metrics.iou_from_confusion
        denom = max(self.steps, 1) * world
        out["losses"] = (sums / denom).cpu()
        return out


@dataclass
class CoTrainConfig:
    num_classes: int
    train_jsd: bool = True
    train_adv: bool = False
    cot_weight: float = 1.0            # cot_scheduler.value (cotraining_totalloss.py:248)
    adv_weight: float = 1.0            # adv_scheduler.value
    fgsm_eps: float = 0.05             # adv_training_dict['eplision'] (config/ACDC_config_cotraing.yaml)
    ignore_index: int = 255
    keep_rows: bool = False
    meter: str = "dice"                # "dice": DiceMeter on both branches (CoTrainer); "iou": IoU meter on the labeled
                                       # branch only (CoTrainer_City, trainer/cotraining_city.py:212,236-257)


class CoTrainStep:
    """The iteration above for K networks ``nets[k](img) -> logits [B,C,H,W]`` and their optimizers."""

    def __init__(self, nets: Sequence[nn.Module], optimizers: Sequence[torch.optim.Optimizer], cfg: CoTrainConfig,
                 device, ddp: bool = True, check_mode: Optional[str] = "deferred"):
        """``check_mode``: the assertion policy this loop runs under (``_runtime.set_check_mode``).  The default
        'deferred' is what "nothing is read back inside the loop" needs: in the drop-ins' default 'eager' mode every
        fused call that carries labels reads its 16-byte flag word back (one host sync per call, so that the
        reference's AssertionError fires at the reference's call site); here the flags stay on the device and are
        raised by ``report.reduce()``.  The mode is scoped to ``step()`` / ``evaluate()``; None leaves the process-wide
        mode in force."""
        assert len(nets) == len(optimizers) and len(nets) >= 1
        assert check_mode in (None, "eager", "deferred", "off")
        self.check_mode = check_mode
        self.device = torch.device(device)
        self.cfg = cfg
        self.K = len(nets)
        self.raw_nets = list(nets)
        if ddp and D.is_distributed():
            from torch.nn.parallel import DistributedDataParallel as DDP
            idx = self.device.index
            nets = [DDP(n, device_ids=[idx], gradient_as_bucket_view=True, broadcast_buffers=False) for n in nets]
        self.nets = list(nets)
        self.optimizers = list(optimizers)
        assert cfg.meter in ("dice", "iou")
        self.report = DeviceReport(self.K, cfg.num_classes, self.device, keep_rows=cfg.keep_rows,
                                   with_confusion=cfg.meter == "iou")

    # ---- the adversarial branch (cotraining_totalloss.py:366-393) with the fused KL
    def _fgsm_adv(self, img_2: Tensor, gt_2: Tensor, unl_img: Tensor) -> Tensor:
        # the FGSM pass only needs d loss / d image: it runs on the bare module (DDP's reducer is built around
        # .backward(), and the reference discards this pass's parameter gradients anyway, AEGenerator.py:29)
        net_src, net_dst = self.raw_nets[1], self.nets[0]
        img = torch.cat((img_2, unl_img), 0).detach().clone().requires_grad_(True)
        pred = net_src(img)
        gt = torch.cat((gt_2, pred.detach().max(1)[1][gt_2.shape[0]:].unsqueeze(1)), 0)   # AEGenerator.py:22-23
        loss = supervised_from_logits(pred, gt, ignore_index=self.cfg.ignore_index)
        (g_img,) = torch.autograd.grad(loss, img)
        img_adv, _ = fgsm_perturb(img.detach(), g_img, self.cfg.fgsm_eps)
        real = softmax_dim1(pred.detach())
        return kl_consistency_from_logits(net_dst(img_adv.detach()), real)

    def step(self, labeled: Sequence[Tuple[Tensor, Tensor]], unlabeled: Optional[Tuple[Tensor, Tensor]] = None) -> Tensor:
        """labeled: K pairs (img [B,Cin,H,W], gt [B,1,H,W] int64) already on the device (this rank's shard);
        unlabeled: (img, gt) -- gt feeds the unlabeled Dice meters only, as in the reference.  Returns the total
        loss tensor (not synchronised)."""
        with _runtime.check_mode(self.check_mode):
            return self._step(labeled, unlabeled)

    def _step(self, labeled, unlabeled):
        cfg, C, dev = self.cfg, self.cfg.num_classes, self.device
        sup_losses, total = [], 0
        for k, (img, gt) in enumerate(labeled):
            logits = self.nets[k](img)
            if cfg.meter == "iou":   # loss + gradient + confusion counts of the IoU meter from one launch
                sup = supervised_from_logits(logits, gt, ignore_index=cfg.ignore_index, confusion=self.report.conf[k])
            else:
                counts = torch.zeros(img.shape[0], C, 3, dtype=torch.int64, device=dev)
                sup = supervised_from_logits(logits, gt, ignore_index=cfg.ignore_index, dice_counts=counts)
                self.report.add_counts("lab", k, counts)
            sup_losses.append(sup)
            total = total + sup
        jsd = adv = None
        if cfg.train_jsd and unlabeled is not None:
            uimg, ugt = unlabeled
            ulogits = [net(uimg) for net in self.nets]
            if cfg.meter == "iou":   # no meters on the unlabeled branch (cotraining_city.py:250-257)
                jsd = jsd_consistency_from_logits(ulogits, weight=1.0)
            else:
                ucounts = torch.empty(self.K, uimg.shape[0], C, 3, dtype=torch.int64, device=dev)
                jsd = jsd_consistency_from_logits(ulogits, weight=1.0, labels=ugt, dice_counts=ucounts, accumulate=False)
                for k in range(self.K):
                    self.report.add_counts("unlab", k, ucounts[k])
            total = total + cfg.cot_weight * jsd
        if cfg.train_adv and unlabeled is not None and self.K >= 2:
            adv = self._fgsm_adv(labeled[1][0], labeled[1][1], unlabeled[0])
            total = total + cfg.adv_weight * adv
        for opt in self.optimizers:
            opt.zero_grad(set_to_none=True)
        total.backward()
        for opt in self.optimizers:
            opt.step()
        self.report.add_losses(sup_losses, jsd, adv)
        return total.detach()

    @torch.no_grad()
    def evaluate(self, img: Tensor, gt: Tensor) -> Tensor:
        """``_eval_loop`` body (:288-293) for one batch: K forward passes, CE + Dice counts per network.
        Returns int64 counts ``[K,B,C,3]``; '2d' rows / '3d' batch Dice follow from ``dice_from_counts``."""
        C = self.cfg.num_classes
        out = torch.zeros(self.K, img.shape[0], C, 3, dtype=torch.int64, device=self.device)
        with _runtime.check_mode(self.check_mode):
            for k, net in enumerate(self.raw_nets):
                supervised_from_logits(net(img), gt, ignore_index=self.cfg.ignore_index, dice_counts=out[k])
        return out


def init_distributed(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """torchrun launcher glue: (rank, world, local_rank) from the environment; initialises the process group
    (NCCL when CUDA is present, gloo otherwise) if WORLD_SIZE > 1.  Replaces ``nn.DataParallel``
    (models/segmentators.py:34-36) with one process per GPU."""
    import os
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend, **kw)
    return rank, world, local
