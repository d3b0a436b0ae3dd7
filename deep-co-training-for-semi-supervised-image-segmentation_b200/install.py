"""Install the B200 drop-ins into an importable reference tree WITHOUT editing it.

    import generalframework                      # the unmodified reference
    import dct_b200
    dct_b200.install()                           # rebinding only; configs and CoTrainer untouched

After ``install()`` the reference's trainers pick up the CUDA-backed classes at
their existing call sites:
  get_loss_fn('jsd' | 'cross_entropy')     -> loss.LOSS[...]                  (loss/__init__.py:6-16)
  KL_Divergence_2D(reduce=True)(adv, real) -> trainer-module globals           (cotraining_totalloss.py:13,392)
  VATGenerator / FSGMGenerator             -> trainer-module globals           (cotraining_totalloss.py:15)
  DiceMeter / IoU / KappaMetrics           -> metrics package + trainer globals (cotraining_totalloss.py:18)
  metrics2.DiceMeter / metrics2.KappaMetrics -> their own twins (metrics.DiceMeter2, ensemble.KappaMetrics2): the
                                              fork used by trainer/mean_teacher_trainer.py:18 reports differently
                                              (metrics2/dice_meter.py:43,82-84; metrics2/kappa.py:31-32)
  dice_coef / dice_batch / class2one_hot / probs2one_hot / pred2class / ... (the star-imported helpers of
  utils/utils.py:73-235)                   -> trainer-module globals ONLY      (trainer/trainer.py:171-175,222-227)
The helper functions are rebound only inside ``<package>.trainer.*``: the dataset code calls ``class2one_hot`` on
host tensors inside DataLoader workers (utils.py:188), which must keep the reference's CPU implementation.
``uninstall()`` restores the originals.
"""
import importlib
import sys

from . import ensemble, generators, loss, metrics, utils

_REBIND = {
    "JSD_2D": loss.JSD_2D, "JSD": loss.JSD, "Entropy_2D": loss.Entropy_2D, "Entropy": loss.Entropy,
    "KL_Divergence_2D": loss.KL_Divergence_2D, "KL_Divergence_2D_Logit": loss.KL_Divergence_2D_Logit,
    "KL_div": loss.KL_div, "CrossEntropyLoss2d": loss.CrossEntropyLoss2d,
    "FSGMGenerator": generators.FSGMGenerator, "VATGenerator": generators.VATGenerator,
    "DiceMeter": metrics.DiceMeter, "IoU": metrics.IoU, "ConfusionMatrix": metrics.ConfusionMatrix,
    "KappaMetrics": ensemble.KappaMetrics, "Kappa2Annotator": ensemble.Kappa2Annotator,
}
# classes DEFINED under <package>.metrics2 keep that fork's host-side semantics (same kernels underneath)
_REBIND_METRICS2 = {"DiceMeter": metrics.DiceMeter2, "KappaMetrics": ensemble.KappaMetrics2}
_REBIND_TRAINER_FUNCS = {name: getattr(utils, name) for name in (
    "simplex", "one_hot", "uniq", "sset", "intersection", "union", "pred2class", "probs2class", "class2one_hot",
    "probs2one_hot", "predlogit2one_hot", "meta_dice", "dice_coef", "dice_batch")}
_OURS = set(_REBIND.values()) | set(_REBIND_METRICS2.values())
_saved = []


def install(package: str = "generalframework") -> int:
    """Rebind the hot-path names in every already-imported ``<package>.*`` module. Returns #rebindings."""
    try:
        importlib.import_module(package + ".loss")
        importlib.import_module(package + ".metrics")
        importlib.import_module(package + ".metrics2")
    except Exception:
        pass
    n = 0
    for modname, mod in list(sys.modules.items()):
        if mod is None or not (modname == package or modname.startswith(package + ".")):
            continue
        for name, repl in _REBIND.items():
            cur = mod.__dict__.get(name)
            if not isinstance(cur, type) or cur in _OURS:   # absent, not a class, or already ours (install() twice)
                continue
            if ".metrics2." in (getattr(cur, "__module__", "") or "") + ".":
                repl = _REBIND_METRICS2.get(name, repl)   # dispatch on where the class was defined, not on its name
            if cur is not repl:
                _saved.append((mod, name, cur))
                setattr(mod, name, repl)
                n += 1
        if modname.startswith(package + ".trainer"):
            for name, repl in _REBIND_TRAINER_FUNCS.items():
                cur = mod.__dict__.get(name)
                if cur is not None and cur is not repl and callable(cur):
                    _saved.append((mod, name, cur))
                    setattr(mod, name, repl)
                    n += 1
        reg = mod.__dict__.get("LOSS")
        if isinstance(reg, dict) and "jsd" in reg:
            for key, repl in (("jsd", loss.JSD_2D), ("cross_entropy", loss.CrossEntropyLoss2d)):
                if key in reg and reg[key] is not repl:
                    _saved.append((reg, key, reg[key]))
                    reg[key] = repl
                    n += 1
    return n


def uninstall() -> None:
    while _saved:
        holder, name, orig = _saved.pop()
        if isinstance(holder, dict):
            holder[name] = orig
        else:
            setattr(holder, name, orig)
