"""Importable alias of the ``pl-convlstm-gan_b200`` package (hyphens are not valid in `import`)."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("pl-convlstm-gan_b200")
sys.modules[__name__] = _pkg
