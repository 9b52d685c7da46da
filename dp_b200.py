"""Importable alias for the package directory `disruption-prediciton-based-on-multimodal-deep-learning_b200/`
(hyphens cannot appear in an `import` statement).  `import dp_b200` and `from dp_b200.R2Plus1D import ...`
resolve to the one and only copy of that package."""
import importlib
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)
_NAME = "disruption-prediciton-based-on-multimodal-deep-learning_b200"
_pkg = importlib.import_module(_NAME)
for _k, _v in list(sys.modules.items()):
    if _k.startswith(_NAME + "."):
        sys.modules["dp_b200" + _k[len(_NAME):]] = _v
sys.modules["dp_b200"] = _pkg
