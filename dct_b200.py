"""Importable alias of the package directory ``deep-co-training-for-semi-supervised-image-segmentation_b200``
(a hyphenated name cannot appear in an ``import`` statement).  ``import dct_b200`` yields that package;
its submodules are registered as ``dct_b200.<name>`` too, so ``from dct_b200.loss import JSD_2D`` works
and refers to the same module objects.
"""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_REAL = "deep-co-training-for-semi-supervised-image-segmentation_b200"
_pkg = importlib.import_module(_REAL)
for _name, _mod in list(sys.modules.items()):
    if _name.startswith(_REAL + "."):
        sys.modules[__name__ + _name[len(_REAL):]] = _mod
sys.modules[__name__] = _pkg
